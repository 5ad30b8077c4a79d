"""GPU tuning sweep for the brute-force NN kernel: pairs/s per variant and size (run under gpurun)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_alignment_b200 import IcpEngine, meshgen  # noqa: E402

out = {}
with IcpEngine() as e:
    out["fp32_peak_tflops"] = e.measure_fp32_peak()
    out["copy_gbs"] = e.measure_copy_bandwidth()
print(json.dumps(out), flush=True)
sizes = [(100_000, 100_000), (1_000_000, 1_000_000)]
variants = [int(v) for v in os.environ.get("VARIANTS", "1,2,3,4,5,6").split(",")]
modes = [int(v) for v in os.environ.get("MODES", "0").split(",")]
for (m, n) in sizes:
    cfg = meshgen.config_cloud(m, n, 200, 201)
    for mode, v in [(mo, v) for mo in modes for v in (variants if mo == 0 else [0])]:
        with IcpEngine(nn_variant=v, nn_mode=mode) as e:
            e.set_base(cfg["base"])
            e.set_moving(cfg["moving"])
            e.set_matrices(np.eye(4), np.eye(4))
            e.set_profiling(True)
            e.nn()      # warm
            e.set_profiling(True)
            reps = 3
            for _ in range(reps):
                e._check(e._lib.icp_nn(e._h, None, None))
            ms, k = e.kernel_time_ms(0)
            ms /= k
            rec = dict(m=m, n=n, mode=mode, variant=v, ms=ms, pairs_per_s=m * n / (ms * 1e-3))
            print(json.dumps(rec), flush=True)
