#!/usr/bin/env python
"""bench.py -- ICP iterations/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c3|c2|c4|c1]

A "step" is one ICP iteration (the loop body of OBJECT_OT_icp_align.execute, reference
operators/icp_align.py:96-151): transform the moving vertices, nearest base vertex for each,
threshold, fp64 covariance, SVD, pose update.  Default workload: C3 = 1M moving vs 1M base
synthetic "bunny-like" clouds (the configuration the north-star target is quoted on).

  value   iterations/s with vertices resident in HBM; device time by CUDA events on the engine's
          stream, summed over the K timed steps, MAX over ranks; L2 flushed between steps
  e2e     iterations/s through the public API with HOST buffers: every step uploads the moving
          vertices + matrices from pinned memory and reads the 4x4 result and stats back
  N > 1   the base vertex set is sharded across ranks (strong scaling on the same workload);
          per iteration one ncclAllReduce(MIN, uint64[M]) + one ncclAllReduce(SUM, double[20])
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_moving, n_base, seed_base, seed_moving, description)
    "c1": (1002, 1002, None, None, "C1: two 1002-vertex icospheres"),
    "c2": (100_000, 100_000, 100, 101, "C2: 100k moving vs 100k base bunny-like clouds"),
    "c3": (1_000_000, 1_000_000, 200, 201, "C3: 1M moving vs 1M base bunny-like clouds"),
    "c4": (1_000_000, 10_000_000, 300, 301, "C4: 1M moving vs 10M base, base sharded across ranks"),
    "c5": (5_000_000, 5_000_000, 500, 501, "C5: 5M vs 5M, 25% outliers, reject_fraction 0.25, icp_include mask"),
    "c2s": (100_000, 100_000, 100, 101, "C2 surface: 100k moving vs a 100k-vertex / 200k-triangle base MESH (closest point on surface)"),
    "c3s": (1_000_000, 1_000_000, 200, 201, "C3 surface: 1M moving vs a 1M-vertex / 2M-triangle base MESH (closest point on surface)"),
}


def make_workload(name):
    from object_alignment_b200 import meshgen

    if name == "c1":
        return meshgen.config_c1()
    if name == "c5":
        from object_alignment_b200 import build_vlist_from_arrays

        cfg = meshgen.config_c5(WORKLOADS[name][0], "include")
        cfg["vlist"] = build_vlist_from_arrays(cfg["moving"].shape[0], cfg["groups"])   # icp_align.py:56-80
        return cfg
    m, n, sb, sm, _ = WORKLOADS[name]
    cfg = meshgen.config_cloud(m, n, sb, sm)
    if name.endswith("s"):
        verts, tris = meshgen.bunny_like_mesh(n, sb)
        cfg["base"] = verts
        cfg["triangles"] = meshgen.triangle_corners(verts, tris)
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the observed power range
        thr = (min(pw) + max(pw)) / 2 if max(pw) > min(pw) else 0
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw)}


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step(cfg, sample, provider):
    """One bounded sample of the reference's per-iteration work on the host CPU, through the
    oracle port of make_pairs + affine_matrix_from_points (a single-threaded per-vertex Python
    loop, like the reference).  Returns seconds for `sample` vertices."""
    import contextlib
    import io

    from oracle import icp_oracle as orc

    align, base = provider["align"], provider["base"]
    vlist = list(range(sample))
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        A, B, ds = orc.make_pairs_port(align, base, provider["bvh"], vlist, cfg["thresh"], 1, True)
        orc.affine_matrix_from_points_port(A, B, scale=False)
    return time.perf_counter() - t0


def build_cpu_provider(cfg):
    from oracle import icp_oracle as orc
    from oracle import mathutils_shim as ms

    align = ms.Object("align", cfg["moving"])
    base = ms.Object("base", cfg["base"])
    return dict(align=align, base=base, bvh=orc.NearestVertexBVH(cfg["base"], mode="kdtree_fast"))


def cpu_baseline(cfg, sample, steps):
    """cpu_baseline object: iterations/s of the oracle port on 1 host core, extrapolated from a
    bounded sample of the moving vertices (per-vertex cost x M_eff)."""
    prov = build_cpu_provider(cfg)
    m = cfg["moving"].shape[0]
    sample = min(sample, m)
    cpu_reference_step(cfg, min(sample, 500), prov)                    # warm
    secs = [cpu_reference_step(cfg, sample, prov) for _ in range(steps)]
    per_vertex = min(secs) / sample
    return {
        "value": 1.0 / (per_vertex * m), "unit": "iterations/s", "cores": 1, "kind": "port",
        "sample": "%d of %d moving vertices vs the full %d-vertex base (cKDTree-backed find_nearest), "
                  "%d timed passes, per-vertex cost %.1f us extrapolated to M_eff"
                  % (sample, m, cfg["base"].shape[0], steps, per_vertex * 1e6),
        "us_per_vertex": per_vertex * 1e6, "host_cpus": os.cpu_count(),
    }


def cpu_fair_comparator(cfg):
    """A vectorised CPU ICP iteration on ALL host cores (scipy cKDTree + numpy Kabsch): not the
    reference, reported beside it so the GPU speed-up can be read against a strong CPU too."""
    from scipy.spatial import cKDTree

    from oracle import icp_oracle as orc

    t0 = time.perf_counter()
    tree = cKDTree(cfg["base"].astype(np.float64))
    build = time.perf_counter() - t0
    t0 = time.perf_counter()
    dd, ii = tree.query(cfg["moving"].astype(np.float64), k=1, workers=-1)
    keep = dd < cfg["thresh"]
    A = cfg["moving"][keep].astype(np.float64).T
    B = cfg["base"][ii[keep]].astype(np.float64).T
    orc.affine_matrix_from_points_port(A, B)
    it = time.perf_counter() - t0
    return {"value": 1.0 / it, "unit": "iterations/s", "cores": os.cpu_count(),
            "kind": "vectorised cKDTree.query(workers=-1) + numpy Kabsch (not the reference)",
            "tree_build_s": build}


def run_reference(args, rank):
    if rank != 0:
        return
    cfg = make_workload(args.workload)
    m, n = cfg["moving"].shape[0], cfg["base"].shape[0]
    prov = build_cpu_provider(cfg)
    sample = min(args.cpu_sample, m)
    for _ in range(args.warmup):
        cpu_reference_step(cfg, min(sample, 2000), prov)
    t = [cpu_reference_step(cfg, sample, prov) for _ in range(args.steps)]
    per_vertex = (sum(t) / len(t)) / sample
    ms_per_step = per_vertex * m * 1e3
    value = 1e3 / ms_per_step
    sample_txt = ("each step = the oracle port of make_pairs+affine_matrix_from_points over %d of %d "
                  "moving vertices vs the full base; ms_per_step is that time scaled to M_eff = %d"
                  % (sample, m, m))
    line = {
        "impl": "reference", "metric": "icp_iterations_per_sec", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][4], "n_moving": m, "n_base": n,
                   "thresh": cfg["thresh"], "sample_fraction": 1.0, "align_meth": "RIGID"},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": 1, "kind": "port",
                         "sample": sample_txt, "us_per_vertex": per_vertex * 1e6,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "correspondences_per_sec": value * m,
        "note": "the reference is a single-threaded per-vertex Python loop (functions/general.py:280-309); "
                "1 core is all it can use. /root/reference is absent on the GPU box, so this is the oracle port.",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    from object_alignment_b200 import IcpEngine, PinnedArray
    from object_alignment_b200 import dist as icpdist

    dist = icpdist.init_rendezvous() if world > 1 else None
    cfg = make_workload(args.workload)
    mov, base = cfg["moving"], cfg["base"]
    m, n = mov.shape[0], base.shape[0]
    ident = np.eye(4, dtype=np.float32)

    nn_mode = 1 if args.nn_mode == "bvh" else 0
    surface = "triangles" in cfg
    if surface:
        nn_mode = 2
    shard_q = args.shard == "queries" or (args.shard == "auto" and nn_mode != 0)
    vlist = cfg.get("vlist")
    m_eff = int(vlist.shape[0]) if vlist is not None else m
    eng = IcpEngine(thresh=cfg["thresh"], use_target=False, align_meth="0", stride=1,
                    device=local_rank, rank=rank, world_size=world, nn_variant=args.nn_variant,
                    nn_mode=nn_mode, reject_fraction=cfg.get("reject_fraction", 0.0),
                    shard_mode=1 if (shard_q and world > 1) else 0)
    t0 = time.perf_counter()
    if surface:
        eng.set_base_triangles(cfg["triangles"]) if shard_q or world == 1 else eng.set_base_triangles_sharded(cfg["triangles"])
    elif shard_q and world > 1:
        eng.set_base(base)
    else:
        eng.set_base_sharded(base)
    base_upload_s = time.perf_counter() - t0
    icpdist.connect_engine(eng, dist)
    if shard_q and world > 1:
        # this rank's contiguous slice of the (already sampled) query positions
        from object_alignment_b200 import shard_range
        pos = np.arange(m, dtype=np.int32) if vlist is None else vlist
        lo, hi = shard_range(pos.shape[0], rank, world)
        vlist = np.ascontiguousarray(pos[lo:hi])
    pinned = PinnedArray(mov.shape, np.float32)
    pinned.array[...] = mov
    eng.set_moving(pinned.array, vlist)
    eng.set_matrices(ident, ident)

    def barrier():
        eng.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- resident (kernel-side) measurement: K timed steps, L2 flushed between them ----------
    for _ in range(args.warmup):
        eng.run(1)
    eng.set_matrices(ident, ident)
    eng.reset_convergence()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    eng.set_profiling(True)
    launches0 = eng.launch_count()
    step_ms = []
    pairs = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        eng.flush_l2()
        done, conv, recs = eng.run(1)
        step_ms.append(eng.last_elapsed_ms())
        pairs.append(recs[0].pairs)
    barrier()
    wall_s = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() - launches0
    nn_ms, nn_launches = eng.kernel_time_ms(0)
    mom_ms, mom_launches = eng.kernel_time_ms(1)
    col_ms, col_launches = eng.kernel_time_ms(4)
    eng.set_profiling(False)
    total_ms = sum(step_ms)
    if dist is not None:
        total_ms = icpdist.allreduce_max_scalar(total_ms, dist)
    final_pose = eng.matrix_world()

    # ---- end-to-end through the public API with host buffers -------------------------------
    eng.set_matrices(ident, ident)
    eng.reset_convergence()
    pose = ident.copy()
    for _ in range(min(args.warmup, 2)):
        eng.set_moving(pinned.array, vlist); eng.set_matrices(pose, ident); eng.step()
    pose = ident.copy()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.set_moving(pinned.array, vlist)     # H2D of this step's inputs from pinned memory
        eng.set_matrices(pose, ident)
        M, st = eng.step()                      # D2H: the fp64 4x4 + the stats record
        pose = eng.matrix_world()               # D2H: the updated matrix_world
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        e2e_s = icpdist.allreduce_max_scalar(e2e_s, dist)
    h2d = mov.nbytes + 2 * 64 + (vlist.nbytes if vlist is not None else 0)
    d2h = 16 * 8 + 168 + 64

    if rank != 0:
        eng.close()
        return
    ms_per_step = total_ms / args.steps
    value = 1e3 / ms_per_step
    fp32_peak = eng.measure_fp32_peak()
    copy_gbs = eng.measure_copy_bandwidth()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    n_local = n if (shard_q and world > 1) else n // world
    nn_avg_ms = nn_ms / max(nn_launches, 1)
    m_all, m = m, m_eff                                        # work is counted on the sampled queries
    flops = 8.0 * m * n_local                                  # 8 FLOP per (query, base) pair
    achieved = flops / (nn_avg_ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "nn_traffic.json"))).get(args.workload)
    except (OSError, ValueError):
        pass
    if nn_mode != 0:
        line_roofline = bvh_roofline(m // (world if shard_q else 1), n_local, nn_avg_ms, hbm_peak_of(peaks), args.workload)
        if surface:
            line_roofline["kernel"] = "k_nn_surface_defer<fused>"
    sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
    pairs_per_s = m * n_local / (nn_avg_ms * 1e-3) if nn_avg_ms > 0 else 0.0
    cyc_per_pair = 148 * 4 * 32 * sm_hz * 1e6 / pairs_per_s if pairs_per_s else None
    roofline = line_roofline if nn_mode != 0 else {
        "kernel": bruteforce_kernel_name(args.nn_variant, n_local),
        "bound": "fp32",
        "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
        "traffic": traffic,
        "dram_gbs_from_traffic": (traffic / (nn_avg_ms * 1e-3) / 1e9) if (traffic and nn_avg_ms > 0) else None,
        "dram_frac_of_hbm_peak": (traffic / (nn_avg_ms * 1e-3) / 1e9 / hbm_peak_of(peaks)) if (traffic and nn_avg_ms > 0) else None,
        "peak_source": "FFMA/FFMA2 micro-benchmark measured in this run (icp_measure_fp32_peak); "
                       "MEASURED_PEAKS.json has no fp32 CUDA-core figure",
        "algorithmic": "8 FLOP x M_eff x N_local pair evaluations per launch (SURVEY 8d)",
        "pairs_per_s": pairs_per_s, "avg_launch_ms": nn_avg_ms,
        "share_of_step": nn_ms / sum(step_ms),
        # the lower-bound kernels execute fewer flops than the algorithmic 8 per pair, so `frac` can
        # exceed 1; what binds them is the issue rate: on sm_100a every 3-register fp32 warp
        # instruction (FFMA2, FMNMX3) holds its scheduler for 2 cycles (tools/ubench/pipes.cu)
        "issue_model": {"sm_cycles_per_pair_lane": cyc_per_pair,
                        "floor_cycles_per_pair": BRUTEFORCE_FLOOR.get(bruteforce_kernel_name(args.nn_variant, n_local).split("<")[0]),
                        "note": "floor = 2 cycles x (FFMA2 + FMNMX3 warp instructions per pair): direct 3.5, "
                                "expansion filter 2.0, cylinder filter 1.5, slab filter 1.0 instructions per pair"},
        "why_not_hbm": "brute-force NN is fp32-issue bound: %.0f MB compulsory traffic per iteration vs "
                       "%.1e pair evaluations (SURVEY D4)" % ((12 * n_local + 52 * m) / 1e6, float(m) * n_local),
    }
    if nn_mode == 0 and roofline["issue_model"]["floor_cycles_per_pair"] and cyc_per_pair:
        roofline["issue_model"]["frac_of_floor"] = roofline["issue_model"]["floor_cycles_per_pair"] / cyc_per_pair
    stream_bytes = 12.0 * n_local + 52.0 * m
    hbm_peak = hbm_peak_of(peaks)
    mom_avg_ms = mom_ms / max(mom_launches, 1)
    roofline_streaming = {
        "kernel": "k_filter_moments(+reduce)", "bound": "hbm",
        "achieved": (48.0 * m) / (mom_avg_ms * 1e-3) / 1e9 if mom_avg_ms > 0 else None,
        "peak": hbm_peak, "unit": "GB/s",
        "frac": ((48.0 * m) / (mom_avg_ms * 1e-3) / 1e9 / hbm_peak) if mom_avg_ms > 0 else None,
        "algorithmic": "48 B per query: key 8 + q 12 + base4 gather 16 + moving vertex 12",
        "avg_launch_ms": mom_avg_ms,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
        "copy_gbs_this_run": copy_gbs,
    }
    line = {
        "metric": "icp_iterations_per_sec", "value": value, "unit": "iterations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][4], "n_moving": m_all, "m_eff": m, "n_base": n,
                   "reject_fraction": cfg.get("reject_fraction", 0.0), "base_shards": 1 if (shard_q and world > 1) else world, "query_shards": world if (shard_q and world > 1) else 1,
                   "thresh": cfg["thresh"] if np.isfinite(cfg["thresh"]) else "inf", "sample_fraction": 1.0,
                   "align_meth": "RIGID", "nn_mode": "surface" if surface else args.nn_mode, "nn_variant": args.nn_variant,
                   "l2": "flushed between timed steps (256 MB memset)",
                   "collectives": ("none" if world == 1 else "ncclAllReduce SUM f64[20] per step" if shard_q
                                   else "ncclAllReduce MIN u64[M] + SUM f64[20] per step")},
        "correspondences_per_sec": value * m,
        "pair_evals_per_sec": value * m * n,
        "e2e": {"value": args.steps / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "base_upload_s_once": base_upload_s,
                "api": "IcpEngine.set_moving(pinned) + set_matrices + step() + matrix_world() per step"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_streaming": roofline_streaming,
        "clocks": clocks,
        "kernel_ms_total": {"nn": nn_ms, "filter_moments": mom_ms, "collectives": col_ms,
                            "steps_device": sum(step_ms), "wall_incl_l2_flush": wall_s * 1e3},
        "pairs_kept_last": int(pairs[-1]) if pairs else None,
        "final_pose_row0": [float(x) for x in final_pose[0]],
    }
    eng.close()
    if world == 1 and not args.no_secondary and vlist is None and not cfg.get("reject_fraction") and not surface:
        other = 0 if nn_mode == 1 else 1
        line["next_rows" if other == 1 else "north_star_kernel"] = {
            ("N2_bvh_search" if other == 1 else "bruteforce"): measure_secondary(
                cfg, local_rank, other, args.steps if other == 1 else min(args.steps, 5), 3)}
    if world == 1 and not args.no_cpu and vlist is None and not surface:
        line["cpu_baseline"] = cpu_baseline(cfg, args.cpu_sample, 2)
        line["cpu_fair"] = cpu_fair_comparator(cfg)
    print(json.dumps(line), flush=True)


BRUTEFORCE_FLOOR = {"k_nn_bruteforce": 7.0, "k_nn_bruteforce_filter": 4.0, "k_nn_bruteforce_cyl": 3.0,
                    "k_nn_bruteforce_slab": 2.0}


def bruteforce_kernel_name(variant, n_local):
    """The K2 kernel the library picks for the timed (seeded) passes (icp_b200.cu launch_nn)."""
    tiles = (n_local + 1023) // 1024
    if variant == 0:
        return "k_nn_bruteforce_slab<4,3,64>" if tiles >= 16 else "k_nn_bruteforce<4,4,1>"
    if 40 <= variant <= 43:
        return "k_nn_bruteforce_slab"
    if 30 <= variant <= 36:
        return "k_nn_bruteforce_cyl"
    if 20 <= variant <= 29:
        return "k_nn_bruteforce_filter"
    return "k_nn_bruteforce"


def hbm_peak_of(peaks):
    return peaks.get("hbm_gbs", 6650.0)


def bvh_roofline(m, n_local, avg_ms, hbm_peak, workload):
    """Roofline of k_nn_bvh: compulsory DRAM bytes per launch = queries in (12 B) + keys read and
    written (16 B) per query + one pass over the tree (64 B per 16-vertex run) and the sorted
    vertices (16 B each).  The tree walk itself is served by L1/L2 (the index of a 1M-vertex base
    is 24 MB against a 126 MB L2), so this fraction is small by construction; the walk's requested
    bytes are reported beside it."""
    compulsory = 28.0 * m + (16.0 + 4.0) * n_local
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "nn_traffic.json"))).get(workload + "_bvh")
    except (OSError, ValueError):
        pass
    ach = compulsory / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else None
    return {"kernel": "k_nn_bvh", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach / hbm_peak if ach else None, "traffic": traffic,
            "algorithmic": "28 B per query + 20 B per base vertex (index read once)",
            "avg_launch_ms": avg_ms, "queries_per_s": m / (avg_ms * 1e-3) if avg_ms > 0 else None,
            "dram_gbs_from_traffic": (traffic / (avg_ms * 1e-3) / 1e9) if (traffic and avg_ms > 0) else None,
            "note": "latency/issue-bound tree walk: per warp of 32 queries ~100 node visits x 64 B + ~35 leaf "
                    "visits x 256 B requested, served from L1/L2"}


def measure_secondary(cfg, local_rank, mode, steps, warmup):
    """Same resident measurement for the other NN mode on one GPU (reported beside the primary)."""
    from object_alignment_b200 import IcpEngine

    ident = np.eye(4, dtype=np.float32)
    with IcpEngine(thresh=cfg["thresh"], use_target=False, device=local_rank, nn_mode=mode) as e:
        e.set_base(cfg["base"])
        e.set_moving(cfg["moving"])
        e.set_matrices(ident, ident)
        for _ in range(warmup):
            e.run(1)
        e.set_matrices(ident, ident)
        e.reset_convergence()
        e.set_profiling(True)
        ms = []
        for _ in range(steps):
            e.flush_l2()
            e.run(1)
            ms.append(e.last_elapsed_ms())
        nn_ms, nn_n = e.kernel_time_ms(0)
        pose = e.matrix_world()
        # whole 50-iteration operator-style run, device resident, no flushes
        e.set_matrices(ident, ident)
        e.reset_convergence()
        e.set_profiling(False)
        e.run(50)
        run50 = e.last_elapsed_ms()
    return {"nn_mode": "bvh" if mode == 1 else "bruteforce", "value": 1e3 * steps / sum(ms),
            "unit": "iterations/s", "ms_per_step": sum(ms) / steps, "steps": steps,
            "nn_avg_launch_ms": nn_ms / max(nn_n, 1), "ms_per_step_first": ms[0], "ms_per_step_last": ms[-1],
            "run50_ms": run50, "run50_iterations_per_s": 50e3 / run50,
            "final_pose_row0": [float(x) for x in pose[0]]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--nn-variant", type=int, default=0)
    ap.add_argument("--nn-mode", default="bruteforce", choices=["bruteforce", "bvh"],
                    help="bruteforce = the north-star O(M N) kernel (default, primary line); "
                         "bvh = row N2, the exact spatially-indexed search")
    ap.add_argument("--cpu-sample", type=int, default=100_000)
    ap.add_argument("--shard", default="auto", choices=["auto", "base", "queries"],
                    help="multi-GPU split: base (north star: keys all-reduce MIN) or queries (replicated base, "
                         "only the 20-double moment sum); auto = base for bruteforce, queries for bvh/surface")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(json.dumps({"error": "launch with torchrun for --gpus > 1"}))
        sys.exit(2)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
