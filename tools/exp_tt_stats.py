"""Tile-tree walk statistics at C3 (library built with -DICP_TT_STATS, selected through ICP_B200_LIB):
how many lanes of a warp need each leaf tile that is scanned."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_alignment_b200 import IcpEngine, meshgen, _native  # noqa: E402

cfg = meshgen.config_cloud(1_000_000, 1_000_000, 200, 201)
ident = np.eye(4, dtype=np.float32)
with IcpEngine(thresh=0.5, use_target=False, nn_mode=1) as e:
    e.set_base(cfg["base"]); e.set_moving(cfg["moving"]); e.set_matrices(ident, ident)
    e.run(3)
    out = (C.c_uint64 * 4)()
    lib = e._lib
    lib.icp_debug_counters(e._h, out, 1)
    e.run(1)
    lib.icp_debug_counters(e._h, out, 1)
    print(json.dumps({"scans_by_lanes_needing": {"1-2": out[0], "3-4": out[1], "5-8": out[2], "9-32": out[3]}}))
