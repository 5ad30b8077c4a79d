"""First (unseeded) iteration and a whole 50-iteration run at C3, brute force: exact tree-search seeds (default)
against greedy tile-tree seeds (nn_variant 79) and the unseeded expansion-filter pass (59)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_alignment_b200 import IcpEngine, meshgen  # noqa: E402

cfg = meshgen.config_cloud(1_000_000, 1_000_000, 200, 201)
ident = np.eye(4, dtype=np.float32)
for v in [int(x) for x in os.environ.get("VARIANTS", "0,79,59,0").split(",")]:
    with IcpEngine(thresh=0.5, use_target=False, nn_variant=v) as e:
        e.set_base(cfg["base"]); e.set_moving(cfg["moving"]); e.set_matrices(ident, ident)
        e.run(1)                      # index builds + first pass
        first_with_build = e.last_elapsed_ms()
        e.run(3)
        e.set_moving(cfg["moving"][::-1].copy())   # a different cloud of the same count would keep the seeds: force a cold pass instead
        e.set_base(cfg["base"])       # new base upload: every index and seed is dropped, the builds run again
        e.set_moving(cfg["moving"]); e.set_matrices(ident, ident); e.reset_convergence()
        e.set_profiling(True)
        t0 = time.perf_counter(); e.run(1); e.synchronize(); wall = time.perf_counter() - t0
        cold = e.last_elapsed_ms()
        prof = {k: e.kernel_time_ms(i) for i, k in enumerate(["nn", "moments", "solve", "transform_and_seed", "collectives"])}
        e.set_profiling(False)
        e.run(49)
        rest = e.last_elapsed_ms()
        print(json.dumps(dict(variant=v, first_iteration_with_cold_process_ms=first_with_build, first_iteration_ms=cold,
                              first_iteration_wall_ms=wall * 1e3, next_49_ms=rest, first_iteration_kernels=prof, pose_row0=[float(x) for x in e.matrix_world()[0]])), flush=True)
