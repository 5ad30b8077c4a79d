"""TEST INFRASTRUCTURE ONLY -- goldens for SURVEY section 8f row N4 from the UNMODIFIED reference.

Run here (the container that has /root/reference):  python -m oracle.gen_golden_n4

  n4_pick_*.npz       OBJECT_OT_align_pick_points.align_obj (operators/align_pick_points.py:415-465) driven
                      on hand-"picked" point lists: truncation to the shorter list, rigid / scale solve,
                      matrix_world @ new_mat, take_m_with propagation to scene objects named "m_*"
  n4_take_m_with.npz  OBJECT_OT_icp_align.execute with settings.take_m_with (icp_align.py:123-127)

The operator instances are created without their interactive invoke(); only the arithmetic methods run.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from object_alignment_b200 import meshgen  # noqa: E402
from oracle import icp_oracle, mathutils_shim as ms, ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def scene_objects(seed):
    """Two objects the transform must follow ("m_*") and one it must not touch."""
    mats = [meshgen.rigid_matrix(30.0 + 10 * i, (0.5 * i, -0.25, 0.125 * i), axis=(0.1 + i, 1.0, -0.4)) for i in range(3)]
    tri = meshgen.icosphere(1)
    return [ms.Object("m_marker", tri, mats[0]), ms.Object("marker", tri, mats[1]), ms.Object("m_2", tri, mats[2])]


def pick_scenarios():
    rng = np.random.default_rng(2024)
    R = meshgen.rotation_about_axis(25.0, (0.3, -1.0, 0.5))
    t = np.array([0.4, -0.3, 0.8])
    S = []
    a = (rng.random((3, 3)) - 0.5) * 4
    S.append(dict(name="n4_pick_rigid3", align=a, base=(R @ a.T).T + t, prefs=dict(align_meth="0")))
    a = (rng.random((6, 3)) - 0.5) * 4
    b = 1.3 * (R @ a.T).T + t + rng.normal(0, 1e-2, (6, 3))
    S.append(dict(name="n4_pick_scale6", align=a, base=b, prefs=dict(align_meth="1"),
                  mw_align=meshgen.rigid_matrix(-12.0, (0.2, 0.1, -0.3), axis=(1.0, 0.2, 0.1))))
    a = (rng.random((5, 3)) - 0.5) * 4
    b = np.concatenate([(R @ a.T).T + t + rng.normal(0, 5e-3, (5, 3)), rng.random((2, 3))])
    S.append(dict(name="n4_pick_truncate_takem", align=a, base=b, prefs=dict(align_meth="0", take_m_with=True),
                  mw_align=meshgen.rigid_matrix(8.0, (-0.1, 0.3, 0.2)), scene=True))
    S.append(dict(name="n4_pick_truncate_align_longer", align=np.concatenate([a, rng.random((3, 3))]), base=b[:5],
                  prefs=dict(align_meth="0")))
    return S


def run_pick(sc):
    mod = ref_loader.load_pick_points()
    import bpy as bpy_stub                      # the stub module ref_loader installed

    prefs = ms.Settings(**sc["prefs"])
    obj = ms.Object("align", meshgen.icosphere(1), sc.get("mw_align"))
    others = scene_objects(1) if sc.get("scene") else []
    bpy_stub.context.scene.objects = [obj] + others
    op = mod.OBJECT_OT_align_pick_points.__new__(mod.OBJECT_OT_align_pick_points)
    op.align_points = [ms.Vector(p) for p in np.asarray(sc["align"], np.float32)]
    op.base_points = [ms.Vector(p) for p in np.asarray(sc["base"], np.float32)]
    op.obj_align = obj
    ctx = ms.Context(obj, [obj])
    saved = mod.get_addon_preferences
    mod.get_addon_preferences = lambda: prefs
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            op.align_obj(ctx)
    finally:
        mod.get_addon_preferences = saved
        bpy_stub.context.scene.objects = []
    return obj, others, ctx, prefs


def run_take_m_with():
    general, icp_align = ref_loader.load_reference()
    import bpy as bpy_stub                      # the stub module ref_loader installed

    c1 = meshgen.config_c1()
    align = ms.Object("align", c1["moving"])
    base = ms.Object("base", c1["base"])
    others = scene_objects(2)
    prefs = ms.Settings(sample_fraction=1.0, icp_iterations=6, use_target=False, take_m_with=True)
    from mathutils import bvhtree as _bvhtree_stub
    _bvhtree_stub.BVHTree = icp_oracle.NearestVertexBVH
    icp_align.BVHTree = icp_oracle.NearestVertexBVH
    bpy_stub.context.scene.objects = [align, base] + others
    saved = icp_align.get_addon_preferences
    icp_align.get_addon_preferences = lambda: prefs
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            assert icp_align.OBJECT_OT_icp_align().execute(ms.Context(align, [align, base])) == {"FINISHED"}
    finally:
        icp_align.get_addon_preferences = saved
        bpy_stub.context.scene.objects = []
    return c1, align, others, prefs


def main():
    os.makedirs(OUT, exist_ok=True)
    for sc in pick_scenarios():
        start = [o.matrix_world.a.copy() for o in (scene_objects(1) if sc.get("scene") else [])]
        obj, others, ctx, prefs = run_pick(sc)
        out = dict(align_points=np.asarray(sc["align"], np.float32), base_points=np.asarray(sc["base"], np.float32),
                   mw_align=np.asarray(sc.get("mw_align", np.identity(4)), np.float32),
                   matrix_world=obj.matrix_world.a.copy(), tag_count=np.int32(obj.tag_count),
                   view_layer_updates=np.int32(ctx.view_layer.updates),
                   pref_align_meth=np.array(prefs.align_meth), pref_take_m_with=np.array(prefs.take_m_with),
                   scene_names=np.array([o.name for o in others], dtype="U32"),
                   scene_start=np.array(start, np.float32).reshape(-1, 4, 4),
                   scene_final=np.array([o.matrix_world.a for o in others], np.float32).reshape(-1, 4, 4))
        np.savez_compressed(os.path.join(OUT, sc["name"] + ".npz"), **out)
        print("%-32s K=%d" % (sc["name"], min(len(sc["align"]), len(sc["base"]))))
    c1, align, others, prefs = run_take_m_with()
    start = [o.matrix_world.a.copy() for o in scene_objects(2)]
    np.savez_compressed(os.path.join(OUT, "n4_take_m_with.npz"),
                        moving=c1["moving"], base=c1["base"], matrix_world=align.matrix_world.a.copy(),
                        scene_names=np.array([o.name for o in others], dtype="U32"),
                        scene_start=np.array(start, np.float32), scene_final=np.array([o.matrix_world.a for o in others], np.float32),
                        scene_tags=np.array([o.tag_count for o in others], np.int32),
                        **{"pref_" + k: np.array(v) for k, v in vars(prefs).items()})
    print("n4_take_m_with written")


if __name__ == "__main__":
    main()
