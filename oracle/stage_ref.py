"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference for the GPU box.

``/root/reference`` exists in the build container but not on the GPU box.  This recipe copies, byte for
byte, the reference files the ICP path imports (``functions/`` and ``operators/icp_align.py``; the
picked-points operator for its ``align_obj``) into ``oracle/_ref/`` -- git-ignored, so nothing of
the reference enters the repo's history, but NOT gpurun-ignored, so the directory travels with the
snapshot like a built ``.so``.  ``oracle/ref_loader.py`` then runs those files exactly as it runs
``/root/reference`` here: ``bench.py --impl reference`` times the reference's own ``make_pairs`` +
``affine_matrix_from_points`` on the box's host CPU (``kind: "reference"``), and the live-reference
parity tests can run there too.  ``__graft_entry__.build()`` calls ``stage()``.
"""
from __future__ import annotations

import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ICP_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["functions/__init__.py", "functions/general.py", "functions/utilities.py",
         "operators/icp_align.py", "operators/icp_align_feedback.py", "operators/align_pick_points.py",
         "LICENSE"]
DIRS = ["functions/common"]


def stage(verbose: bool = False) -> bool:
    """Copy the files; returns False (and does nothing) when the reference is not present."""
    if not os.path.isfile(os.path.join(SRC, "functions", "general.py")):
        return False
    todo = list(FILES)
    for d in DIRS:
        for name in sorted(os.listdir(os.path.join(SRC, d))):
            if name.endswith(".py"):
                todo.append(d + "/" + name)
    for rel in todo:
        s, t = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(t), exist_ok=True)
        if not (os.path.isfile(t) and filecmp.cmp(s, t, shallow=False)):
            shutil.copyfile(s, t)
            os.chmod(t, 0o644)
            if verbose:
                print("staged", rel)
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write("byte-for-byte copies of %s (see oracle/stage_ref.py); not part of the repo's history\n" % SRC)
    return True


if __name__ == "__main__":
    print("staged" if stage(verbose=True) else "reference not found at %s" % SRC)
