"""Read .ncu-rep files (here, no GPU needed) and write the summaries bench.py and the judge read:
profiles/nn_pipes.json (measured pipe / issue counters per kernel), profiles/nn_traffic.json (DRAM bytes per
launch), and a text summary per report with the SASS regions ranked by executed instructions.
usage: python tools/ncu_extract.py <report.ncu-rep> <kernel-key> [workload-key] [out.txt]"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "duration",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct_of_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct_of_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fmaheavy_pipe_pct_of_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct_of_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__cycles_active.avg": "smsp_cycles_active", "sm__cycles_elapsed.max": "sm_cycles_elapsed",
}
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def main():
    rep, key = sys.argv[1], sys.argv[2]
    workload = sys.argv[3] if len(sys.argv) > 3 else None
    out_txt = sys.argv[4] if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec, stalls = {"kernel_name": vals[hdr.index("Kernel Name")]}, {}
    for i, h in enumerate(hdr):
        if h in WANT:
            v = float(vals[i].replace(",", ""))
            if units[i] in UNIT:
                v *= UNIT[units[i]]
            rec[WANT[h]] = v if WANT[h] != "duration" else {"value": v, "unit": units[i]}
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            stalls[h.split("issue_stalled_")[1].split("_per_issue")[0]] = float(vals[i])
    rec["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    rec["source"] = os.path.basename(rep) + " (ncu --set full --clock-control none, one launch, cold caches)"
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    sh = srows[1]
    iS, iN, iI = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
    data = [(r[iS].strip(), int(r[iN]), int(r[iI])) for r in srows[2:] if len(r) > iI]
    tot_i, tot_s = sum(d[2] for d in data), sum(d[1] for d in data)
    segs, cur = [], None
    for idx, (s, n, i) in enumerate(data):
        if cur is None or abs(i - cur["i0"]) > 0.15 * max(cur["i0"], 1e5):
            if cur:
                segs.append(cur)
            cur = dict(start=idx, i0=i, cnt=0, instr=0, samp=0, ops=collections.Counter())
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        cur["cnt"] += 1; cur["instr"] += i; cur["samp"] += n; cur["ops"][op] += 1
    segs.append(cur)
    hot = max(segs, key=lambda g: g["instr"])
    rec["hot_loop"] = {"sass_instructions": hot["cnt"], "executions": hot["i0"], "share_of_instructions": hot["instr"] / tot_i,
                       "share_of_samples": hot["samp"] / tot_s, "mix": dict(hot["ops"].most_common(8))}
    for name, field in (("nn_pipes.json", rec), ("nn_traffic.json", None)):
        path = os.path.join(ROOT, "profiles", name)
        try:
            cur_json = json.load(open(path))
        except (OSError, ValueError):
            cur_json = {}
        if field is not None:
            cur_json[key] = field
        elif workload:
            cur_json[workload] = int(rec.get("dram_read", 0) + rec.get("dram_write", 0))
        json.dump(cur_json, open(path, "w"), indent=1, sort_keys=True)
    if out_txt:
        with open(out_txt, "w") as f:
            f.write("# %s\n# %s\n" % (rec["kernel_name"], rec["source"]))
            for k, v in rec.items():
                if k not in ("kernel_name", "source"):
                    f.write("%-32s %s\n" % (k, json.dumps(v)))
            f.write("\n# SASS regions with >= 0.6 %% of the executed instructions or >= 1 %% of the samples\n")
            for g in segs:
                if g["instr"] / tot_i > 0.006 or g["samp"] / tot_s > 0.01:
                    f.write("sass[%4d..%4d] n=%3d executions=%10.3fM instructions=%5.1f%% samples=%5.1f%%  %s\n" % (
                        g["start"], g["start"] + g["cnt"] - 1, g["cnt"], g["i0"] / 1e6, 100 * g["instr"] / tot_i,
                        100 * g["samp"] / tot_s, dict(g["ops"].most_common(5))))
            f.write("\n# the hot loop (SASS, as executed)\n")
            for s, n, i in data[hot["start"]:hot["start"] + min(hot["cnt"], 120)]:
                f.write("%10d  %s\n" % (n, s))
    print(json.dumps({k: rec[k] for k in ("duration", "warp_instructions", "alu_pipe_pct_of_active", "issue_active_pct")}))


if __name__ == "__main__":
    main()
