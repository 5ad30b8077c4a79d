#!/usr/bin/env python
"""bench.py -- ICP iterations/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c3|c2|c4|c5|c1]

A "step" is one ICP iteration (the loop body of OBJECT_OT_icp_align.execute, reference
operators/icp_align.py:96-151): transform the moving vertices, nearest base vertex for each,
threshold, fp64 covariance, SVD, pose update.  Default workload: C3 = 1M moving vs 1M base
synthetic "bunny-like" clouds (the configuration the north-star target is quoted on).

  value     iterations/s with vertices resident in HBM; device time by CUDA events on the engine's
            stream, summed over the K timed steps, MAX over ranks; L2 flushed between steps
  e2e       iterations/s through the public API with HOST buffers: every step uploads the moving
            vertices + matrices from pinned memory and reads the 4x4 result and stats back
  roofline  the dominant kernel (the nearest-vertex search) against what bounds it: warp-instruction
            issue.  `frac` = issue-slot utilisation of its hot loop (instructions per pair counted in the
            SASS, profiles/), pipe utilisations beside it, ncu's measured counters when profiles/ has them
  N > 1     the base vertex set is sharded across ranks (strong scaling on the same workload); per
            iteration one ncclAllReduce(MIN, uint64[M]) + one ncclAllReduce(SUM, double[20]).  The line
            carries `parity` (sharded result == the same run unsharded on rank 0) and a `c4` sub-object
            (1M moving vs 10M base, the north star's scaling configuration) measured the same way.
  --impl reference   the UNMODIFIED reference make_pairs + affine_matrix_from_points (staged under
            oracle/_ref by __graft_entry__.build()) on the host CPU, one bounded sample per step.
"""
from __future__ import annotations

import argparse
import hashlib
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
METRIC = "icp_iterations_per_sec"
UNIT = "iterations/s"


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


def make_config(args, cfg, world):
    """The `config` object: identical for both arms (the driver compares them)."""
    m, n = cfg["moving"].shape[0], cfg["base"].shape[0]
    vlist = cfg.get("vlist")
    return {
        "workload": WORKLOADS[args.workload][4], "n_moving": m,
        "m_eff": int(vlist.shape[0]) if vlist is not None else m, "n_base": n,
        "thresh": cfg["thresh"] if np.isfinite(cfg["thresh"]) else "inf",
        "reject_fraction": cfg.get("reject_fraction", 0.0), "sample_fraction": 1.0, "align_meth": "RIGID",
        "correspondence": "surface" if "triangles" in cfg else "vertex",
        "n_gpus": world,
        "l2": "flushed between timed steps (256 MB memset); CPU arm: n/a",
    }


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
class CpuReference:
    """The reference's own per-iteration work on the host CPU.

    kind "reference": the UNMODIFIED functions/general.py make_pairs (:257-329) and
    affine_matrix_from_points (:105-217), executed from /root/reference or its staged copy
    oracle/_ref through oracle/ref_loader.py, over mathutils-shim objects with a cKDTree-backed
    nearest-vertex `find_nearest`.  kind "port": oracle/icp_oracle.py's restatement of the
    same loop, only when the reference files are absent.  A step covers every `stride`-th vertex of
    the vertex list through make_pairs' own `sample` argument (general.py:274-275) against the FULL
    base; the reference is a single-threaded per-vertex Python loop, so 1 core is all it can use."""

    def __init__(self, cfg):
        from oracle import icp_oracle as orc
        from oracle import mathutils_shim as ms
        from oracle import ref_loader

        self.cfg = cfg
        self.align = ms.Object("align", cfg["moving"])
        self.base = ms.Object("base", cfg["base"])
        self.bvh = orc.NearestVertexBVH(cfg["base"], mode="kdtree_fast")
        vl = cfg.get("vlist")
        self.vlist = list(range(cfg["moving"].shape[0])) if vl is None else [int(v) for v in vl]
        self.thresh = cfg["thresh"] if np.isfinite(cfg["thresh"]) else 1e30
        if ref_loader.available():
            general, _ = ref_loader.load_reference()
            self.make_pairs, self.solve = general.make_pairs, general.affine_matrix_from_points
            self.kind, self.root = "reference", ref_loader.REFERENCE_ROOT
        else:
            self.make_pairs = orc.make_pairs_port
            self.solve = lambda A, B, shear, scale, usesvd: orc.affine_matrix_from_points_port(A, B, scale=scale)
            self.kind, self.root = "port", None

    def step(self, stride):
        """One sampled iteration body (icp_align.py:101-109); returns (seconds, queries, pairs kept)."""
        import contextlib

        t0 = time.perf_counter()
        with open(os.devnull, "w") as sink, contextlib.redirect_stdout(sink):      # general.py:309 prints per vertex
            A, B, ds = self.make_pairs(self.align, self.base, self.bvh, self.vlist, self.thresh, stride, True)
            self.solve(A, B, shear=False, scale=False, usesvd=True)
        return time.perf_counter() - t0, len(self.vlist[0::stride] if stride > 1 else self.vlist), A.shape[1]


def reference_stride(m_eff, budget_queries):
    return max(1, int(round(m_eff / float(budget_queries))))


def cpu_baseline(cfg, budget_queries, passes):
    ref = CpuReference(cfg)
    m_eff = len(ref.vlist)
    stride = reference_stride(m_eff, budget_queries)
    ref.step(max(stride, m_eff // 500))                                    # warm
    runs = [ref.step(stride) for _ in range(passes)]
    secs, q = min(r[0] for r in runs), runs[0][1]
    per_vertex = secs / q
    return {
        "value": 1.0 / (per_vertex * m_eff), "unit": UNIT, "cores": 1, "kind": ref.kind,
        "sample": "%d timed passes of the %s make_pairs + affine_matrix_from_points over every %d-th of the %d "
                  "queries (%d per pass, make_pairs' own `sample` stride) vs the full %d-vertex base with a "
                  "cKDTree-backed find_nearest; best pass %.2f s; value = 1 / (per-vertex cost x M_eff)"
                  % (passes, "unmodified reference's" if ref.kind == "reference" else "oracle port of",
                     stride, m_eff, q, cfg["base"].shape[0], secs),
        "us_per_vertex": per_vertex * 1e6, "sampled_pass_s": secs, "host_cpus": os.cpu_count(),
        "reference_root": ref.root,
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
    return {"value": 1.0 / it, "unit": UNIT, "cores": os.cpu_count(),
            "kind": "vectorised cKDTree.query(workers=-1) + numpy Kabsch (not the reference)",
            "tree_build_s": build}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg = make_workload(args.workload)
    ref = CpuReference(cfg)
    m_eff = len(ref.vlist)
    stride = reference_stride(m_eff, args.cpu_sample)
    for _ in range(args.warmup):
        ref.step(max(stride, m_eff // 2000))                               # short untimed passes
    runs = [ref.step(stride) for _ in range(args.steps)]
    secs = [r[0] for r in runs]
    q = runs[0][1]
    ms_sampled = 1e3 * sum(secs) / len(secs)
    per_vertex = (sum(secs) / len(secs)) / q
    full_ms = per_vertex * m_eff * 1e3
    value = 1e3 / full_ms
    sample_txt = ("each step = the %s make_pairs + affine_matrix_from_points over every %d-th query (%d of %d, "
                  "make_pairs' own `sample` stride, general.py:274-275) vs the full %d-vertex base"
                  % ("UNMODIFIED reference" if ref.kind == "reference" else "oracle port of the reference",
                     stride, q, m_eff, cfg["base"].shape[0]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_sampled,                       # the time actually measured per (sampled) step
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": make_config(args, cfg, args.gpus),
        "value_is": "full-workload iterations/s = 1 / (measured per-query cost x M_eff); the per-query cost of the "
                    "reference's loop does not depend on how many queries a pass covers" if stride > 1
                    else "measured on the full workload",
        "sampled_step": {"queries": q, "of": m_eff, "stride": stride, "ms": ms_sampled,
                         "pairs_kept": int(runs[-1][2])},
        "full_iteration_ms_extrapolated": full_ms,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": ref.kind,
                         "sample": sample_txt, "us_per_vertex": per_vertex * 1e6,
                         "host_cpus": os.cpu_count(), "reference_root": ref.root},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "correspondences_per_sec": q / (ms_sampled * 1e-3),
        "note": "the reference is a single-threaded per-vertex Python loop (functions/general.py:280-309): 1 core "
                "is all it can use; its BVHTree.find_nearest is stood in for by a cKDTree k=1 query per call",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def build_engine(cfg, args, rank, world, local_rank, nn_mode, shard_q, dist):
    """Engine with the base uploaded (sharded as asked) and the communicator built."""
    from object_alignment_b200 import IcpEngine
    from object_alignment_b200 import dist as icpdist

    surface = "triangles" in cfg
    eng = IcpEngine(thresh=cfg["thresh"], use_target=False, align_meth="0", stride=1,
                    device=local_rank, rank=rank, world_size=world, nn_variant=args.nn_variant,
                    nn_mode=nn_mode, reject_fraction=cfg.get("reject_fraction", 0.0),
                    shard_mode=1 if (shard_q and world > 1) else 0)
    t0 = time.perf_counter()
    if surface:
        eng.set_base_triangles(cfg["triangles"]) if shard_q or world == 1 else eng.set_base_triangles_sharded(cfg["triangles"])
    elif shard_q and world > 1:
        eng.set_base(cfg["base"])
    else:
        eng.set_base_sharded(cfg["base"])
    upload_s = time.perf_counter() - t0
    icpdist.connect_engine(eng, dist)
    return eng, upload_s


def timed_steps(eng, steps, warmup, barrier, sampler=None):
    """W untimed + K timed single-iteration runs from the identity pose, L2 flushed between steps."""
    ident = np.eye(4, dtype=np.float32)
    eng.set_matrices(ident, ident)
    eng.reset_convergence()
    for _ in range(warmup):
        eng.run(1)
    eng.set_matrices(ident, ident)
    eng.reset_convergence()
    barrier()
    if sampler is not None:
        sampler.start()
    eng.set_profiling(True)
    launches0 = eng.launch_count()
    step_ms, pairs = [], []
    wall0 = time.perf_counter()
    for _ in range(steps):
        eng.flush_l2()
        done, conv, recs = eng.run(1)
        step_ms.append(eng.last_elapsed_ms())
        pairs.append(int(recs[0].pairs))
    barrier()
    out = {"step_ms": step_ms, "pairs": pairs, "wall_s": time.perf_counter() - wall0,
           "clocks": sampler.stop() if sampler is not None else None,
           "launches": eng.launch_count() - launches0,
           "nn": eng.kernel_time_ms(0), "mom": eng.kernel_time_ms(1), "solve": eng.kernel_time_ms(2),
           "transform": eng.kernel_time_ms(3), "col": eng.kernel_time_ms(4), "pose": eng.matrix_world()}
    eng.set_profiling(False)
    return out


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def sharded_parity(eng, cfg, vlist, local_rank, rank, pairs_sharded, steps, args):
    """N > 1: the sharded engine's correspondences under the identity pose, and its kept-pair counts
    per timed step, against the SAME computation on one GPU (rank 0, full base).  Collective."""
    ident = np.eye(4, dtype=np.float32)
    eng.set_matrices(ident, ident)
    eng.reset_convergence()
    idx, d2 = eng.nn()                                   # collective: every rank calls it
    if rank != 0:
        return None
    from object_alignment_b200 import IcpEngine

    n_cmp = min(steps, 5)
    with IcpEngine(thresh=cfg["thresh"], use_target=False, device=local_rank, nn_variant=args.nn_variant,
                   reject_fraction=cfg.get("reject_fraction", 0.0)) as one:
        one.set_base(cfg["base"])
        one.set_moving(cfg["moving"], vlist)
        one.set_matrices(ident, ident)
        idx1, d21 = one.nn()
        for _ in range(max(args.warmup, 3)):
            one.run(1)
        one.set_matrices(ident, ident)
        one.reset_convergence()
        pairs_one = []
        for _ in range(n_cmp):
            pairs_one.append(int(one.run(1)[2][0].pairs))
        pose_one = one.matrix_world()
    same_idx = bool(np.array_equal(idx, idx1))
    same_d2 = bool(np.array_equal(d2.view(np.uint32), d21.view(np.uint32)))
    same_k = pairs_one == [int(p) for p in pairs_sharded[:n_cmp]]
    return {"ok": bool(same_idx and same_d2 and same_k),
            "what": "rank-0 idx/d2 of all %d queries under the identity pose and pairs kept in the first %d timed "
                    "steps, sharded over %d GPUs vs the same engine unsharded on rank 0" % (idx.shape[0], n_cmp, eng.params.world_size),
            "idx_equal": same_idx, "d2_bits_equal": same_d2, "pairs_per_step_equal": same_k,
            "sha_sharded": digest(idx, d2), "sha_unsharded": digest(idx1, d21),
            "pairs_sharded": [int(p) for p in pairs_sharded[:n_cmp]], "pairs_unsharded": pairs_one,
            "unsharded_pose_row0": [float(x) for x in pose_one[0]]}


# instructions per (query, base vertex) pair in the hot loop of each brute-force kernel, counted in the SASS
# (profiles/r2_ncu_k2_q16_summary.txt): (issue slots, alu-pipe cycles, fma-pipe cycles) per pair and lane
K2_MODELS = {
    "k_nn_bruteforce_q16<split>": dict(issue=0.625 + 1.0 / 32, alu=0.75, fma=0.5,
                                       mix="per 8 vertices and query: 2 VIADDMNMX.U16x2 + 1 VIMNMX3.U16x2 (alu pipe, 2 cycles each) "
                                           "+ 2 IMAD (fmaheavy pipe, 2 cycles each); 1 LDS.128 per 4 queries"),
    "k_nn_bruteforce_q16<alu>": dict(issue=0.5 + 1.0 / 32, alu=1.0, fma=0.0,
                                     mix="per 8 vertices: 1/R LDS.128 + 4 VIADDMNMX.U16x2 (alu pipe, 2 cycles each)"),
    "k_nn_bruteforce_slab": dict(issue=1.0 + 1.0 / 16, alu=1.0, fma=1.0, mix="per 4 vertices: 2 FFMA2 + 2 FMNMX3 + 2/R LDS.128"),
    "k_nn_bruteforce_cyl": dict(issue=1.5 + 3.0 / 32, alu=1.0, fma=2.0, mix="per 4 vertices: 4 FFMA2 + 2 FMNMX3"),
    "k_nn_bruteforce_filter": dict(issue=2.0 + 1.0 / 8, alu=1.0, fma=3.0, mix="per 4 vertices: 6 FFMA2 + 2 FMNMX3"),
    "k_nn_bruteforce": dict(issue=3.5 + 3.0 / 32, alu=1.0, fma=6.0, mix="per 4 vertices: 6 FADD2 + 2 FMUL2 + 4 FFMA2 + 2 FMNMX3"),
}


def bruteforce_kernel_name(variant, n_local):
    """The K2 kernel the library picks for the timed (seeded) passes (icp_b200.cu launch_nn)."""
    tiles = (n_local + 1023) // 1024
    if variant == 0:
        return "k_nn_bruteforce_q16<split>" if tiles >= 16 else "k_nn_bruteforce"
    if 60 <= variant <= 79 or variant in (58, 59):
        return "k_nn_bruteforce_q16<split>"
    if 50 <= variant <= 57:
        return "k_nn_bruteforce_q16<alu>"
    if 40 <= variant <= 45:
        return "k_nn_bruteforce_slab"
    if 30 <= variant <= 36:
        return "k_nn_bruteforce_cyl"
    if 20 <= variant <= 29:
        return "k_nn_bruteforce_filter"
    return "k_nn_bruteforce"


def load_profile_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except (OSError, ValueError):
        return {}


def hbm_peak_of(peaks):
    return peaks.get("hbm_gbs", 6650.0)


def bruteforce_roofline(kernel, m, n_local, nn_avg_ms, share, sm_mhz, sm_count, fp32_peak, peaks, workload):
    """Executed-work roofline of the O(M N) search.  The kernel is bound by warp-instruction issue, not by
    HBM and not by fp32 flops (its hot loop evaluates an integer lower bound): `frac` is the
    issue-slot utilisation = hot-loop warp instructions / (SM sub-partitions x cycles), the alu / fma pipe
    utilisations are beside it, and ncu's measured counters of the same kernel (profiles/nn_pipes.json)
    are attached when present.  The SURVEY 8d nominal (8 FLOP per pair) is kept as `algorithmic`."""
    mdl = K2_MODELS[kernel]
    pairs = float(m) * n_local
    secs = nn_avg_ms * 1e-3
    smsp = sm_count * 4
    cycles = secs * sm_mhz * 1e6                                    # per SM sub-partition
    warp_pairs = pairs / 32.0
    util = {k: warp_pairs * mdl[k] / (smsp * cycles) for k in ("issue", "alu", "fma")}
    achieved = warp_pairs * mdl["issue"] / secs / 1e9               # G warp-instructions / s
    peak = smsp * sm_mhz * 1e6 / 1e9                                # 1 warp instruction / cycle / sub-partition
    traffic = load_profile_json("nn_traffic.json").get(workload)
    measured = load_profile_json("nn_pipes.json").get(kernel)
    algo_tflops = 8.0 * pairs / secs / 1e12
    return {
        "kernel": kernel, "bound": "issue",
        "achieved": achieved, "peak": peak, "unit": "Gwarp-inst/s", "frac": util["issue"],
        "peak_source": "%d SM sub-partitions x %.0f MHz (SM clock sampled during the timed region) x 1 warp "
                       "instruction per cycle" % (smsp, sm_mhz),
        "model": {"instructions_per_pair": mdl["issue"], "alu_cycles_per_pair": mdl["alu"],
                  "fma_cycles_per_pair": mdl["fma"], "hot_loop": mdl["mix"],
                  "source": "SASS of the hot loop, profiles/r2_ncu_k2_q16_summary.txt (hot loop as executed, with per-instruction samples); slow paths and tile set-up "
                            "(measured 7 % of the executed instructions) are NOT counted as useful work"},
        "pipe_utilisation": {"issue": util["issue"], "alu": util["alu"], "fma": util["fma"]},
        "ncu_measured": measured,
        "traffic": traffic,
        "dram_gbs_from_traffic": (traffic / secs / 1e9) if traffic else None,
        "dram_frac_of_hbm_peak": (traffic / secs / 1e9 / hbm_peak_of(peaks)) if traffic else None,
        "pairs_per_s": pairs / secs, "avg_launch_ms": nn_avg_ms, "share_of_step": share,
        "sm_cycles_per_pair_lane": smsp * 32 * cycles / pairs,
        "algorithmic": {"flops_per_pair": 8, "tflops": algo_tflops, "fp32_peak_tflops_this_run": fp32_peak,
                        "algorithmic_speedup": algo_tflops / fp32_peak if fp32_peak else None,
                        "note": "SURVEY 8d nominal: 8 FLOP x M_eff x N_local per launch against the FFMA rate "
                                "measured in this run.  It exceeds 1 because the kernel does not execute those "
                                "flops: it proves most pairs irrelevant with a 16-bit interval test"},
        "why_not_hbm": "%.0f MB compulsory traffic per iteration vs %.1e pair evaluations (SURVEY D4)"
                       % ((12 * n_local + 52 * m) / 1e6, pairs),
    }


TREE_KERNELS = {0: "k_nn_tiletree_p<fused>", 13: "k_nn_tiletree<fused>", 12: "k_nn_bvh_packet<fused>"}


def bvh_roofline(m, n_local, avg_ms, hbm_peak, workload, kernel="k_nn_tiletree_p<fused>"):
    """Roofline of the tree search: compulsory DRAM bytes per launch with a cold L2.  Tile tree: per query the
    moving slot (16 B), the key read and written (16 B) and the winner's tile read and written (8 B); per
    base vertex its leaf tile (16 B), its share of the boxes (1 B) and the float4 copy the seed / winner
    gathers touch (16 B).  The walk itself is served by L1/L2 (the index of a 1M-vertex base is 17 MB
    against a 126 MB L2), so this fraction is small by construction: the kernel is bound by instruction
    issue and dependent-load latency, which `ncu_measured` shows."""
    compulsory = 40.0 * m + 33.0 * n_local
    traffic = load_profile_json("nn_traffic.json").get(workload + "_bvh")
    ach = compulsory / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else None
    return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach / hbm_peak if ach else None, "traffic": traffic,
            "algorithmic": "40 B per query + 33 B per base vertex (index and float4 copy read once)",
            "avg_launch_ms": avg_ms, "queries_per_s": m / (avg_ms * 1e-3) if avg_ms > 0 else None,
            "dram_gbs_from_traffic": (traffic / (avg_ms * 1e-3) / 1e9) if (traffic and avg_ms > 0) else None,
            "ncu_measured": load_profile_json("nn_pipes.json").get(kernel),
            "note": "issue- and latency-bound tree walk served from L1/L2; in tree mode the launch is the whole "
                    "iteration (walk + filter + moments + ticketed reduction + solve); see ncu_measured for the "
                    "executed-work view"}


def execute_total(cfg, local_rank, mode, iters=50):
    """What the reference's own timer spans (icp_align.py:50,152): index build + vlist + the whole loop.
    Here: engine creation, base upload + index build, moving upload, `iters` iterations, pose read-back."""
    from object_alignment_b200 import IcpEngine

    ident = np.eye(4, dtype=np.float32)
    t0 = time.perf_counter()
    with IcpEngine(thresh=cfg["thresh"], use_target=False, device=local_rank, nn_mode=mode) as e:
        e.set_base(cfg["base"])
        t1 = time.perf_counter()
        e.set_moving(cfg["moving"])
        e.set_matrices(ident, ident)
        e.run(1)                                   # builds the search index, unseeded first pass
        first_ms = e.last_elapsed_ms()
        t2 = time.perf_counter()
        e.run(iters - 1)
        rest_ms = e.last_elapsed_ms()
        pose = e.matrix_world()
        t3 = time.perf_counter()
    return {"nn_mode": "bvh" if mode == 1 else "bruteforce", "iterations": iters, "total_s": t3 - t0,
            "create_and_base_upload_s": t1 - t0, "moving_upload_index_build_first_iteration_s": t2 - t1,
            "first_iteration_device_ms": first_ms, "remaining_iterations_device_ms": rest_ms,
            "iterations_per_s_incl_everything": iters / (t3 - t0),
            "final_pose_row0": [float(x) for x in pose[0]]}


def measure_secondary(cfg, local_rank, mode, steps, warmup):
    """Same resident measurement for the other NN mode on one GPU (reported beside the primary)."""
    from object_alignment_b200 import IcpEngine

    with IcpEngine(thresh=cfg["thresh"], use_target=False, device=local_rank, nn_mode=mode) as e:
        e.set_base(cfg["base"])
        e.set_moving(cfg["moving"])
        r = timed_steps(e, steps, warmup, e.synchronize)
        ms = r["step_ms"]
        # whole 50-iteration operator-style run, device resident, no flushes
        ident = np.eye(4, dtype=np.float32)
        e.set_matrices(ident, ident)
        e.reset_convergence()
        e.run(50)
        run50 = e.last_elapsed_ms()
    return {"nn_mode": "bvh" if mode == 1 else "bruteforce", "value": 1e3 * steps / sum(ms),
            "unit": UNIT, "ms_per_step": sum(ms) / steps, "steps": steps,
            "nn_avg_launch_ms": r["nn"][0] / max(r["nn"][1], 1), "ms_per_step_first": ms[0], "ms_per_step_last": ms[-1],
            "run50_ms": run50, "run50_iterations_per_s": 50e3 / run50,
            "final_pose_row0": [float(x) for x in r["pose"][0]]}


def run_c4_sub(args, rank, world, local_rank, dist):
    """The north star's scaling configuration (1M moving vs 10M base, base sharded across the N ranks),
    measured like the primary workload with fewer steps; carried on every --gpus N line."""
    from object_alignment_b200 import PinnedArray
    from object_alignment_b200 import dist as icpdist

    sub = argparse.Namespace(**vars(args))
    sub.workload = "c4"
    cfg = make_workload("c4")
    eng, upload_s = build_engine(cfg, sub, rank, world, local_rank, 0, False, dist)
    eng.set_moving(cfg["moving"])

    def barrier():
        eng.synchronize()
        if dist is not None:
            dist.barrier()

    steps, warmup = min(args.steps, 6), 3
    r = timed_steps(eng, steps, warmup, barrier)
    total_ms = sum(r["step_ms"])
    if dist is not None:
        total_ms = icpdist.allreduce_max_scalar(total_ms, dist)
    parity = sharded_parity(eng, cfg, None, local_rank, rank, r["pairs"], steps, sub) if world > 1 else None
    eng.close()
    if rank != 0:
        return None
    m, n = cfg["moving"].shape[0], cfg["base"].shape[0]
    ms = total_ms / steps
    return {"workload": WORKLOADS["c4"][4], "n_moving": m, "n_base": n, "n_gpus": world, "base_shards": world,
            "steps": steps, "warmup": warmup, "metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms,
            "scaling": "strong", "nn_avg_launch_ms": r["nn"][0] / max(r["nn"][1], 1),
            "collectives_ms_per_step": r["col"][0] / steps, "base_upload_s_once": upload_s,
            "pair_evals_per_sec": 1e3 / ms * m * n, "pairs_kept_last": r["pairs"][-1],
            "final_pose_row0": [float(x) for x in r["pose"][0]], "parity": parity}


def run_ours(args, rank, world, local_rank):
    from object_alignment_b200 import PinnedArray
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
    vlist_full = cfg.get("vlist")
    vlist = vlist_full
    m_eff = int(vlist.shape[0]) if vlist is not None else m
    eng, base_upload_s = build_engine(cfg, args, rank, world, local_rank, nn_mode, shard_q, dist)
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
    r = timed_steps(eng, args.steps, args.warmup, barrier, ClockSampler(local_rank) if rank == 0 else None)
    step_ms, pairs, clocks = r["step_ms"], r["pairs"], r["clocks"]
    nn_ms, nn_launches = r["nn"]
    mom_ms, mom_launches = r["mom"]
    col_ms, _ = r["col"]
    total_ms = sum(step_ms)
    rank_ms = total_ms
    if dist is not None:
        total_ms = icpdist.allreduce_max_scalar(rank_ms, dist)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"rank": rank, "steps_ms": rank_ms / args.steps,
                                          "nn_ms": nn_ms / max(nn_launches, 1),
                                          "collectives_ms": col_ms / args.steps})
    else:
        per_rank = None
    final_pose = r["pose"]

    # ---- end-to-end through the public API with host buffers -------------------------------
    collective_upload = world > 1 and not shard_q
    upload = eng.set_moving_collective if collective_upload else eng.set_moving
    eng.set_matrices(ident, ident)
    eng.reset_convergence()
    pose = ident.copy()
    for _ in range(min(args.warmup, 2)):
        upload(pinned.array, vlist); eng.set_matrices(pose, ident); eng.step()
    pose = ident.copy()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        upload(pinned.array, vlist)             # H2D of this step's inputs from pinned memory
        eng.set_matrices(pose, ident)
        M, st = eng.step()                      # D2H: the fp64 4x4 + the stats record
        pose = eng.matrix_world()               # D2H: the updated matrix_world
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        e2e_s = icpdist.allreduce_max_scalar(e2e_s, dist)
    h2d = (mov.nbytes // world if collective_upload else mov.nbytes) + 2 * 64 + (vlist.nbytes if vlist is not None else 0)
    d2h = 16 * 8 + 168 + 64

    parity = None
    if world > 1 and not shard_q and not surface:
        parity = sharded_parity(eng, cfg, vlist_full, local_rank, rank, pairs, args.steps, args)

    if rank == 0:
        fp32_peak = eng.measure_fp32_peak()
        copy_gbs = eng.measure_copy_bandwidth()
    eng.close()
    pinned.close()

    c4 = None
    if args.workload == "c3" and nn_mode == 0 and not args.no_c4:
        c4 = run_c4_sub(args, rank, world, local_rank, dist)
    if rank != 0:
        return

    ms_per_step = total_ms / args.steps
    value = 1e3 / ms_per_step
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    n_local = n if (shard_q and world > 1) else n // world
    nn_avg_ms = nn_ms / max(nn_launches, 1)
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    if nn_mode != 0:
        roofline = bvh_roofline(m_eff // (world if shard_q else 1), n_local, nn_avg_ms, hbm_peak_of(peaks), args.workload,
                                "k_nn_surface_defer<fused>" if surface else TREE_KERNELS.get(args.nn_variant, "k_nn_bvh_packet<fused>"))
    else:
        roofline = bruteforce_roofline(bruteforce_kernel_name(args.nn_variant, n_local), m_eff, n_local, nn_avg_ms,
                                       nn_ms / sum(step_ms), sm_mhz, 148, fp32_peak, peaks, args.workload)
    hbm_peak = hbm_peak_of(peaks)
    mom_avg_ms = mom_ms / max(mom_launches, 1)
    roofline_streaming = {
        "kernel": "k_filter_moments(+reduce)", "bound": "hbm",
        "achieved": (48.0 * m_eff) / (mom_avg_ms * 1e-3) / 1e9 if mom_avg_ms > 0 else None,
        "peak": hbm_peak, "unit": "GB/s",
        "frac": ((48.0 * m_eff) / (mom_avg_ms * 1e-3) / 1e9 / hbm_peak) if mom_avg_ms > 0 else None,
        "algorithmic": "48 B per query: key 8 + q 12 + base4 gather 16 + moving vertex 12",
        "avg_launch_ms": mom_avg_ms,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
        "copy_gbs_this_run": copy_gbs,
    }
    config = make_config(args, cfg, world)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config,
        "engine": {"nn_mode": "surface" if surface else args.nn_mode, "nn_variant": args.nn_variant,
                   "base_shards": 1 if (shard_q and world > 1) else world,
                   "query_shards": world if (shard_q and world > 1) else 1,
                   "collectives": ("none" if world == 1 else "ncclAllReduce SUM f64[20] per step" if shard_q
                                   else "ncclAllReduce MIN u64[M] + SUM f64[20] per step")},
        "correspondences_per_sec": value * m_eff,
        "pair_evals_per_sec": value * m_eff * n,
        "e2e": {"value": args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "base_upload_s_once": base_upload_s,
                "api": ("IcpEngine.set_moving_collective(pinned: 1/N slice H2D + ncclAllGather)" if collective_upload
                        else "IcpEngine.set_moving(pinned)") + " + set_matrices + step() + matrix_world() per step"},
        "gpu_launches": int(r["launches"]),
        "roofline": roofline,
        "roofline_streaming": roofline_streaming,
        "clocks": clocks,
        "kernel_ms_total": {"nn": nn_ms, "filter_moments": mom_ms, "collectives": col_ms,
                            "solve": r["solve"][0], "transform": r["transform"][0],
                            "steps_device": sum(step_ms), "wall_incl_l2_flush": r["wall_s"] * 1e3},
        "per_rank": per_rank,
        "pairs_kept_last": int(pairs[-1]) if pairs else None,
        "final_pose_row0": [float(x) for x in final_pose[0]],
        "parity": parity,
        "c4": c4,
    }
    plain = world == 1 and vlist is None and not cfg.get("reject_fraction") and not surface
    if plain and not args.no_secondary:
        other = 0 if nn_mode == 1 else 1
        line["next_rows" if other == 1 else "north_star_kernel"] = {
            ("N2_bvh_search" if other == 1 else "bruteforce"): measure_secondary(
                cfg, local_rank, other, args.steps if other == 1 else min(args.steps, 5), 3)}
        line["execute_total"] = {"what": "engine creation + base upload + index build + moving upload + 50 iterations + "
                                         "pose read-back, wall clock (the reference's timer spans the same, icp_align.py:50,152)",
                                 "bruteforce": execute_total(cfg, local_rank, 0),
                                 "bvh": execute_total(cfg, local_rank, 1)}
    if plain and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(cfg, args.cpu_sample, 2)
        line["cpu_fair"] = cpu_fair_comparator(cfg)
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--cpu-sample", type=int, default=100_000,
                    help="queries per sampled pass of the CPU reference (stride = M_eff / this)")
    ap.add_argument("--shard", default="auto", choices=["auto", "base", "queries"],
                    help="multi-GPU split: base (north star: keys all-reduce MIN) or queries (replicated base, "
                         "only the 20-double moment sum); auto = base for bruteforce, queries for bvh/surface")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the 1M x 10M sub-measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(json.dumps({"error": "launch with torchrun for --gpus > 1"}))
        sys.exit(2)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
