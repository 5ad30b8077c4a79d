"""Helpers to read tests/golden/*.npz (written by oracle/gen_golden.py from the live reference)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def scenario_names():
    """The ICP-run scenarios of oracle/gen_golden.py (not the solver KATs, not the row-N4 files)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not p.endswith("solver_kat.npz") and not os.path.basename(p).startswith("n4_"))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.z = z
        self.moving, self.base = z["moving"], z["base"]
        self.mw_align, self.mw_base = z["mw_align"], z["mw_base"]
        self.prefs = {k[5:]: z[k].item() for k in z.files if k.startswith("pref_")}
        self.groups = {}
        for g in z["group_order"]:
            g = str(g)
            self.groups[g] = (z["group_%s_mask" % g], z["group_%s_weight" % g])
        self.iterations = int(z["iterations"])
        self.converged = bool(z["converged"])
        off = z["idx_offsets"]
        self.idx = [z["idx_flat"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
        self.K, self.d_stats, self.M = z["K"], z["d_stats"], z["M"]
        self.matrix_world = z["matrix_world"]
        self.vlist = z["vlist"]
        self.base_faces = z["base_faces"] if "base_faces" in z.files else None
