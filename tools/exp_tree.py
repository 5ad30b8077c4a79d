"""Tree search variants (nn_mode BVH) at C3 / C4 sizes: ms per fused iteration, pose changing every step."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_alignment_b200 import IcpEngine, meshgen  # noqa: E402

m = int(os.environ.get("M", 1_000_000))
n = int(os.environ.get("N", 1_000_000))
steps = int(os.environ.get("STEPS", 10))
variants = [int(v) for v in os.environ.get("VARIANTS", "12,0").split(",")]
cfg = meshgen.config_cloud(m, n, 200, 201)
ident = np.eye(4, dtype=np.float32)
ref = None
for v in variants:
    with IcpEngine(thresh=0.5, use_target=False, nn_mode=1, nn_variant=v) as e:
        e.set_base(cfg["base"])
        e.set_moving(cfg["moving"])
        e.set_matrices(ident, ident)
        e.run(1)
        first = e.last_elapsed_ms()
        e.run(2)
        e.set_matrices(ident, ident)
        e.reset_convergence()
        e.set_profiling(True)
        ms, pairs = [], []
        for _ in range(steps):
            e.flush_l2()
            done, conv, recs = e.run(1)
            ms.append(e.last_elapsed_ms())
            pairs.append(recs[0].pairs)
        nn_ms, nn_n = e.kernel_time_ms(0)
        pose = e.matrix_world()
        e.set_profiling(False)
        e.set_matrices(ident, ident)
        e.reset_convergence()
        e.run(50)
        run50 = e.last_elapsed_ms()
        if ref is None:
            ref = (pose.copy(), list(pairs))
        same = bool(np.allclose(pose, ref[0], atol=2e-6) and pairs == ref[1])
        print(json.dumps(dict(variant=v, m=m, n=n, nn_ms=nn_ms / max(nn_n, 1), step_ms=sum(ms) / steps, cold_first_ms=first,
                              run50_ms=run50, same_as_first_variant=same, pairs_last=pairs[-1])), flush=True)
