"""Experiment: brute-force K2 per-iteration time, cold vs seeded (previous winners), by shard size
and variant (run under gpurun).  VARIANTS=0,20,2  SHARDS=1000000,500000,250000,125000"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_alignment_b200 import IcpEngine, meshgen

m = int(os.environ.get("M", 1_000_000)); n = int(os.environ.get("N", 1_000_000))
variants = [int(v) for v in os.environ.get("VARIANTS", "0,2").split(",")]
shards = [int(v) for v in os.environ.get("SHARDS", "1000000,500000,250000,125000").split(",")]
iters = int(os.environ.get("ITERS", 8))
cfg = meshgen.config_cloud(m, n, 200, 201)
for shard in shards:
    for v in variants:
        with IcpEngine(thresh=0.5, use_target=False, nn_variant=v) as e:
            e.set_base(cfg["base"][:shard]); e.set_moving(cfg["moving"]); e.set_matrices(np.eye(4), np.eye(4))
            e.set_profiling(True)
            per, nn = [], []
            for i in range(iters):
                e.run(1); per.append(round(e.last_elapsed_ms(), 3))
                t, k = e.kernel_time_ms(0); nn.append(t)
            nn = [round(b - a, 3) for a, b in zip([0.0] + nn[:-1], nn)]
            print(json.dumps(dict(shard=shard, variant=v, step_ms=per, nn_ms=nn,
                                  pairs_per_s_last=m * shard / (nn[-1] * 1e-3))), flush=True)
