"""Seeded brute-force K2 variants at C3 (or --size): ms per NN launch and per step, pose changing every step."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_alignment_b200 import IcpEngine, meshgen  # noqa: E402

m = int(os.environ.get("M", 1_000_000))
n = int(os.environ.get("N", 1_000_000))
steps = int(os.environ.get("STEPS", 6))
variants = [int(v) for v in os.environ.get("VARIANTS", "42,50,51,52,53,54,55,56,57").split(",")]
cfg = meshgen.config_cloud(m, n, 200, 201)
ident = np.eye(4, dtype=np.float32)
ref = None
for v in variants:
    with IcpEngine(thresh=0.5, use_target=False, nn_variant=v) as e:
        e.set_base(cfg["base"])
        e.set_moving(cfg["moving"])
        e.set_matrices(ident, ident)
        e.run(3)
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
        if ref is None:
            ref = (pose.copy(), list(pairs))
        same = bool(np.array_equal(pose, ref[0]) and pairs == ref[1])
        print(json.dumps(dict(variant=v, m=m, n=n, nn_ms=nn_ms / max(nn_n, 1), step_ms=sum(ms) / steps, first=ms[0],
                              last=ms[-1], same_as_first_variant=same, pairs_last=pairs[-1])), flush=True)
