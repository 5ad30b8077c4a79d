"""CPU tests of the host-side logic that mirrors the reference (no GPU needed)."""
import numpy as np
import pytest

import object_alignment_b200 as oab
from object_alignment_b200 import meshgen
from oracle import icp_oracle as orc
from oracle import mathutils_shim as ms
from tests.golden_util import Golden


@pytest.mark.parametrize("name", ["c1_include", "c1_exclude", "c1_include_wins", "c1_fixed20"])
def test_build_vlist_matches_reference_rule(name):
    g = Golden(name)
    align = ms.Object("a", g.moving, None, g.groups)
    v = oab.build_vlist(align)
    assert v.dtype == np.int32 and np.array_equal(v, g.vlist)             # icp_align.py:56-80
    assert np.array_equal(oab.build_vlist_from_arrays(g.moving.shape[0], g.groups), g.vlist)
    assert np.array_equal(orc.build_vlist_port(g.moving.shape[0], g.groups), g.vlist)


def test_stride_factor():
    assert [oab.stride_from_sample_fraction(s) for s in (0.5, 0.4, 0.67, 1.0, 0.3)] == [2, 2, 1, 1, 3]


def test_shard_ranges_cover_base_exactly():
    for n in (0, 1, 7, 1000, 10_000_000):
        for w in (1, 2, 3, 4, 8):
            r = [oab.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def test_operator_surface():
    op = oab.OBJECT_OT_icp_align
    assert op.bl_idname == "object.align_icp" and op.bl_label == "ICP Align"
    assert op.bl_options == {'REGISTER', 'UNDO'}
    a = ms.Object("a", np.zeros((3, 3), np.float32))
    b = ms.Object("b", np.zeros((3, 3), np.float32))
    assert op.poll(ms.Context(a, [a, b])) is True
    assert op.poll(ms.Context(a, [a])) is False
    prefs = oab.IcpPreferences()
    assert (prefs.icp_iterations, prefs.sample_fraction, prefs.min_start, prefs.target_d,
            prefs.use_target, prefs.take_m_with, prefs.align_meth) == (50, 0.5, 0.5, 0.01, True, False, '0')


def test_meshgen_configs():
    c1 = meshgen.config_c1()
    assert c1["base"].shape == (1002, 3) and c1["moving"].dtype == np.float32
    np.testing.assert_allclose(np.linalg.norm(c1["base"], axis=1), 1.0, atol=1e-6)
    assert len(np.unique(c1["base"], axis=0)) == 1002
    c = meshgen.config_cloud(1000, 2000, 1, 2)
    assert c["base"].shape == (2000, 3) and c["moving"].shape == (1000, 3)
    assert np.array_equal(c["base"], meshgen.config_cloud(1000, 2000, 1, 2)["base"])


def test_hilbert_curve_code_restated_is_a_bijection_with_unit_steps():
    """`curve_code21` (csrc/nn_bvh.cuh: Skilling's transpose algorithm, the order of base vertices, triangles and
    query slots), restated with 4 bits per axis: every cell gets its own position and consecutive positions are
    adjacent cells -- the property the warp patches and the leaf tiles rely on for tight boxes (any order would
    be CORRECT; this one has no jumps)."""
    bits = 4

    def code(gx, gy, gz):
        x = [gx, gy, gz]
        q = 1 << (bits - 1)
        while q > 1:
            p = q - 1
            for i in range(3):
                if x[i] & q:
                    x[0] ^= p
                else:
                    t = (x[0] ^ x[i]) & p
                    x[0] ^= t
                    x[i] ^= t
            q >>= 1
        x[1] ^= x[0]
        x[2] ^= x[1]
        t, q = 0, 1 << (bits - 1)
        while q > 1:
            if x[2] & q:
                t ^= q - 1
            q >>= 1
        x = [v ^ t for v in x]
        c = 0
        for b in range(bits - 1, -1, -1):
            for i in range(3):
                c = (c << 1) | ((x[i] >> b) & 1)
        return c

    n = 1 << bits
    cells = np.array([(a, b, c) for a in range(n) for b in range(n) for c in range(n)])
    codes = np.array([code(*map(int, cell)) for cell in cells])
    assert np.array_equal(np.sort(codes), np.arange(n ** 3))
    walk = cells[np.argsort(codes)]
    assert (np.abs(np.diff(walk, axis=0)).sum(axis=1) == 1).all()
