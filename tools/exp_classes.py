"""Population of the q16 query classes (exponent of the seed's squared distance relative to the median bin) at C5 / C3."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from object_alignment_b200 import IcpEngine  # noqa: E402

for name in sys.argv[1:] or ["c5", "c3"]:
    cfg = bench.make_workload(name)
    ident = np.eye(4, dtype=np.float32)
    with IcpEngine(thresh=0.5 if name != "c5" else 1e30, use_target=False, nn_mode=1 if name != "c5" else 0,
                   reject_fraction=cfg.get("reject_fraction", 0.0)) as e:
        e.set_base(cfg["base"]); e.set_moving(cfg["moving"], cfg.get("vlist")); e.set_matrices(ident, ident)
        e.run(8)
        idx, d2 = e.nn()
    ex = (d2.view(np.uint32) >> 23) & 0xFF
    hist = np.bincount(ex, minlength=256)
    m = int(np.searchsorted(np.cumsum(hist) * 2, d2.size))
    cls = np.where(ex <= m + 2, 0, np.minimum(15, (ex.astype(int) - m - 1) // 2))
    print(json.dumps({"workload": name, "queries": int(d2.size), "median_exponent": m,
                      "median_d2": float(np.median(d2)), "classes": np.bincount(cls, minlength=16).tolist()}))
