"""Experiment: BVH NN time vs query order and vs alignment state (run under gpurun)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_alignment_b200 import IcpEngine, meshgen

def morton_order(p):
    lo, hi = p.min(0), p.max(0)
    g = np.clip(((p - lo) / (hi - lo) * 1023).astype(np.uint64), 0, 1023)
    def ex(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    return np.argsort(ex(g[:, 0]) | (ex(g[:, 1]) << 1) | (ex(g[:, 2]) << 2), kind="stable")

def time_nn(e, reps=5):
    e.set_profiling(True); e.nn(); e.set_profiling(True)
    for _ in range(reps):
        e._check(e._lib.icp_nn(e._h, None, None))
    ms, k = e.kernel_time_ms(0)
    return ms / k

m = int(os.environ.get("M", 1_000_000)); n = int(os.environ.get("N", 1_000_000))
cfg = meshgen.config_cloud(m, n, 200, 201)
aligned = meshgen.bunny_like(m, 201, noise=1e-3)       # same samples without the rigid offset
for variant in (10, 0):
    for name, mov in (("offset", cfg["moving"]), ("aligned", aligned)):
        with IcpEngine(nn_mode=1, nn_variant=variant) as e:
            e.set_base(cfg["base"]); e.set_moving(mov); e.set_matrices(np.eye(4), np.eye(4))
            print(json.dumps(dict(case=name, kernel="per-thread" if variant == 10 else "packet", m=m, n=n, warm_ms=time_nn(e))), flush=True)
import ctypes
for name, mov in (("offset", cfg["moving"]), ("aligned", aligned)):
  for cv in (99, 98):
    with IcpEngine(nn_mode=1, nn_variant=cv) as e:
        e.set_base(cfg["base"]); e.set_moving(mov); e.set_matrices(np.eye(4), np.eye(4))
        buf = (ctypes.c_uint64 * 4)()
        e._lib.icp_debug_counters(e._h, buf, 1)
        e.nn()
        e._lib.icp_debug_counters(e._h, buf, 1)
        print(json.dumps(dict(case=name, kernel="per-thread" if cv == 99 else "packet(per warp)", inner=buf[0] / buf[3], leaf=buf[1] / buf[3], push=buf[2] / buf[3])), flush=True)
for variant in (10, 0):
  with IcpEngine(nn_mode=1, use_target=False, nn_variant=variant) as e:
    e.set_base(cfg["base"]); e.set_moving(cfg["moving"]); e.set_matrices(np.eye(4), np.eye(4))
    e.run(2); e.set_matrices(np.eye(4), np.eye(4)); e.reset_convergence()
    per = []
    for i in range(30):
        e.run(1); per.append(round(e.last_elapsed_ms(), 3))
    print(json.dumps(dict(variant=variant, per_iter_ms=per)))
