"""SURVEY section 8f row N4 on the GPU: picked-points initial alignment (the reference's
OBJECT_OT_align_pick_points.align_obj, operators/align_pick_points.py:415-465, over icp_solve_pairs),
take_m_with propagation (operators/icp_align.py:123-127) and the pinned staging upload -- against
goldens produced by the UNMODIFIED reference (oracle/gen_golden_n4.py)."""
import os
import types

import numpy as np
import pytest

import object_alignment_b200 as oab
from object_alignment_b200 import meshgen
from object_alignment_b200.engine import IcpEngine
from oracle import mathutils_shim as ms
from tests.golden_util import GOLDEN

pytestmark = pytest.mark.gpu


def _scene(z):
    tri = meshgen.icosphere(1)
    return [ms.Object(str(n), tri, z["scene_start"][i]) for i, n in enumerate(z["scene_names"])]


@pytest.mark.parametrize("name", ["n4_pick_rigid3", "n4_pick_scale6", "n4_pick_truncate_takem",
                                  "n4_pick_truncate_align_longer"])
def test_pick_points_align_obj_matches_reference_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    obj = ms.Object("align", meshgen.icosphere(1), z["mw_align"])
    others = _scene(z)
    oab.set_addon_preferences(oab.IcpPreferences(align_meth=str(z["pref_align_meth"]),
                                                 take_m_with=bool(z["pref_take_m_with"])))
    ctx = ms.Context(obj, [obj])
    ctx.scene = types.SimpleNamespace(objects=[obj] + others)
    op = oab.OBJECT_OT_align_pick_points(obj, [ms.Vector(p) for p in z["align_points"]],
                                         [ms.Vector(p) for p in z["base_points"]])
    op.align_obj(ctx)
    k = min(len(z["align_points"]), len(z["base_points"]))
    assert len(op.align_points) == len(op.base_points) == k
    np.testing.assert_allclose(obj.matrix_world.a, z["matrix_world"], rtol=0, atol=1e-5)
    assert obj.tag_count == int(z["tag_count"]) and ctx.view_layer.updates == int(z["view_layer_updates"])
    for o, want, start in zip(others, z["scene_final"], z["scene_start"]):
        np.testing.assert_allclose(o.matrix_world.a, want, rtol=0, atol=1e-5)
        if not o.name.startswith("m_"):
            assert np.array_equal(o.matrix_world.a, start) and o.tag_count == 0
        elif bool(z["pref_take_m_with"]):
            assert o.tag_count == 1 and not np.array_equal(o.matrix_world.a, start)


def test_pick_points_needs_three_pairs():
    obj = ms.Object("align", meshgen.icosphere(1))
    oab.set_addon_preferences(oab.IcpPreferences())
    op = oab.OBJECT_OT_align_pick_points(obj, [ms.Vector((0, 0, 0)), ms.Vector((1, 0, 0))],
                                         [ms.Vector((0, 1, 0)), ms.Vector((1, 1, 0))])
    with pytest.raises(ValueError, match="wrong shape or type"):
        op.align_obj(ms.Context(obj, [obj]))


@pytest.mark.parametrize("modal", [False, True])
def test_take_m_with_follows_every_iteration(modal):
    z = np.load(os.path.join(GOLDEN, "n4_take_m_with.npz"))
    align = ms.Object("align", z["moving"])
    base = ms.Object("base", z["base"])
    others = _scene(z)
    oab.set_addon_preferences(oab.IcpPreferences(
        sample_fraction=float(z["pref_sample_fraction"]), icp_iterations=int(z["pref_icp_iterations"]),
        use_target=bool(z["pref_use_target"]), take_m_with=True, redraw_frequency=4))
    ctx = ms.Context(align, [align, base])
    ctx.scene = types.SimpleNamespace(objects=[align, base] + others)
    if modal:
        op = oab.OBJECT_OT_icp_align_feedback()
        assert op.invoke(ctx, ms.Event("NONE")) == {"RUNNING_MODAL"}
        while op.modal(ctx, ms.Event("TIMER")) == {"RUNNING_MODAL"}:
            pass
    else:
        assert oab.OBJECT_OT_icp_align().execute(ctx) == {"FINISHED"}
    np.testing.assert_allclose(align.matrix_world.a, z["matrix_world"], rtol=0, atol=1e-5)
    for o, want, start in zip(others, z["scene_final"], z["scene_start"]):
        np.testing.assert_allclose(o.matrix_world.a, want, rtol=0, atol=2e-5)
        if o.name.startswith("m_"):
            assert o.tag_count >= 1 and not np.array_equal(o.matrix_world.a, start)
        else:
            assert np.array_equal(o.matrix_world.a, start) and o.tag_count == 0
    # the m_* objects moved by exactly the align object's own total update: M_total = mw_start^-1 mw_final
    total = np.linalg.inv(np.eye(4)) @ align.matrix_world.a.astype(np.float64)
    for o, start in zip(others, z["scene_start"]):
        if o.name.startswith("m_"):
            np.testing.assert_allclose(o.matrix_world.a, start.astype(np.float64) @ total, atol=2e-5)


def test_operator_uploads_through_pinned_staging(monkeypatch):
    """execute() stages the coordinates in page-locked memory (icp_alloc_pinned) and frees it again."""
    from object_alignment_b200 import engine as eng_mod, operators
    made = []
    real = eng_mod.PinnedArray

    class Spy(real):
        def __init__(self, shape, dtype):
            super().__init__(shape, dtype)
            made.append(self)

    monkeypatch.setattr(operators, "PinnedArray", Spy)
    cfg = meshgen.config_c1()
    align, base = ms.Object("align", cfg["moving"]), ms.Object("base", cfg["base"])
    oab.set_addon_preferences(oab.IcpPreferences(icp_iterations=3, sample_fraction=1.0, use_target=False))
    assert oab.OBJECT_OT_icp_align().execute(ms.Context(align, [align, base])) == {"FINISHED"}
    assert len(made) == 1 and made[0].array is None          # allocated once, released after the upload
