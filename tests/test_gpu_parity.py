"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) and the reference-shaped
Python surface, against the CPU oracle and the golden files of the unmodified reference.

Bar: indices / kept sets / A / B bit-exact; fp64 solver outputs within 1e-9; the composed
pose (rotation and translation after all iterations) within 1e-5 of the reference
(BASELINE.json north_star).
"""
import numpy as np
import pytest

import object_alignment_b200 as oab
from object_alignment_b200 import meshgen
from object_alignment_b200.engine import IcpEngine, IcpError
from oracle import icp_oracle as orc
from oracle import mathutils_shim as ms
from tests.golden_util import GOLDEN, Golden, scenario_names

pytestmark = pytest.mark.gpu

NN_VARIANTS = [0, 1, 2, 3, 4, 5, 6, 20, 21, 22, 23, 24, 30, 31, 32, 33, 34, 35, 36]


def _rand_cloud(n, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal((n, 3)) * scale).astype(np.float32)


# ------------------------------------------------------------------ K2: nearest vertex
@pytest.mark.parametrize("variant", NN_VARIANTS)
def test_nn_bit_exact_small(variant):
    base = _rand_cloud(5000, 1)
    base[1234] = base[77]                      # exact tie -> lowest index must win
    base[4999] = base[0]
    mov = np.concatenate([base[[77, 0]], _rand_cloud(3001, 2)])
    with IcpEngine(nn_variant=variant) as e:
        e.set_base(base)
        e.set_moving(mov)
        e.set_matrices(np.eye(4), np.eye(4))
        idx, d2 = e.nn()
    oi, od = orc.nn_bruteforce(mov, base)
    assert idx[0] == 77 and idx[1] == 0 and d2[0] == 0.0
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


@pytest.mark.parametrize("n_base,n_mov", [(1, 1), (1, 300), (15, 7), (16, 257), (1023, 1),
                                          (1024, 1024), (1025, 1025), (4097, 2049),
                                          (9000, 12345)])
def test_nn_ragged_sizes(n_base, n_mov):
    base = _rand_cloud(n_base, 10 + n_base)
    mov = _rand_cloud(n_mov, 20 + n_mov)
    with IcpEngine() as e:
        e.set_base(base)
        e.set_moving(mov)
        e.set_matrices(np.eye(4), np.eye(4))
        idx, d2 = e.nn()
    oi, od = orc.nn_bruteforce(mov, base)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


def test_nn_empty_inputs():
    with IcpEngine() as e:
        e.set_base(np.zeros((0, 3), np.float32))
        e.set_moving(_rand_cloud(10, 1))
        e.set_matrices(np.eye(4), np.eye(4))
        idx, d2 = e.nn()
        assert (idx == -1).all()
        e.set_base(_rand_cloud(10, 2))
        e.set_moving(np.zeros((0, 3), np.float32))
        idx, d2 = e.nn()
        assert idx.shape == (0,)
        with pytest.raises(IcpError) as ei:
            e.step()                                  # no pairs -> the solver's ValueError
        assert ei.value.code == 2


def test_nn_vlist_stride_and_matrices():
    base = _rand_cloud(3000, 5)
    mov = _rand_cloud(2500, 6)
    mw_a = meshgen.rigid_matrix(13.0, (0.2, -0.1, 0.05))
    mw_b = meshgen.rigid_matrix(-21.0, (0.0, 0.3, -0.2), axis=(0.3, 0.1, -1.0)).astype(np.float64)
    mw_b[:3, :3] = mw_b[:3, :3] @ np.diag([1.2, 0.9, 1.05])
    mw_b = mw_b.astype(np.float32)
    vlist = np.sort(np.random.default_rng(7).choice(2500, 1700, replace=False)).astype(np.int32)
    for stride in (0, 1, 2, 3, 7):
        with IcpEngine(stride=stride) as e:
            e.set_base(base)
            e.set_moving(mov, vlist)
            e.set_matrices(mw_a, mw_b)
            idx, d2 = e.nn()
        res = orc.make_pairs_vec(mov, vlist, mw_a, mw_b, base, 1e9, stride, return_all=True)
        assert idx.shape == res["idx"].shape
        assert np.array_equal(idx, res["idx"])
        assert np.array_equal(d2.view(np.uint32), res["d2"].view(np.uint32))


@pytest.mark.parametrize("variant", [20, 21, 30, 33, 35, 36, 40, 43, 44, 45, 50, 51, 54, 57])
@pytest.mark.parametrize("case", ["surface", "far_outliers", "duplicates", "collinear", "volume", "lattice",
                                  "far_from_origin", "sorted_base"])
def test_filter_kernel_matches_oracle(case, variant):
    """The expansion-form filter variant of K2 must return the same keys as the exact metric:
    near-ties, exact ties, coherent base order (many running-min improvements) and data far from
    the origin (where the expansion cancels worst) included."""
    if case == "far_from_origin":
        base = (_rand_cloud(30_000, 81) * 0.5 + np.array([800.0, -500.0, 300.0], np.float32)).astype(np.float32)
        mov = (_rand_cloud(9_000, 82) * 0.5 + np.array([800.0, -500.0, 300.0], np.float32)).astype(np.float32)
    elif case == "sorted_base":
        cfg = meshgen.config_cloud(10_000, 40_000, 83, 84)
        base = cfg["base"][np.lexsort(cfg["base"].T)]
        mov = cfg["moving"]
    else:
        base, mov = _clouds_for_bvh(case)
    with IcpEngine(nn_variant=variant) as e:
        e.set_base(base)
        e.set_moving(mov)
        e.set_matrices(np.eye(4), np.eye(4))
        idx, d2 = e.nn()
        idx2, d22 = e.nn()          # seeded with the exact answer: every tie of the winner must still be visited
    oi, od = orc.nn_bruteforce(mov, base)
    assert np.array_equal(idx, oi) and np.array_equal(idx2, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))
    assert np.array_equal(d22.view(np.uint32), od.view(np.uint32))


def test_filter_kernel_equals_direct_kernel_at_1m():
    cfg = meshgen.config_cloud(1_000_000, 1_000_000, 200, 201)
    res = {}
    for variant in (2, 20, 35, 40, 0):
        with IcpEngine(nn_variant=variant) as e:
            e.set_base(cfg["base"])
            e.set_moving(cfg["moving"])
            e.set_matrices(np.eye(4), np.eye(4))
            res[variant] = e.nn()
            if variant == 0:
                res["seeded"] = e.nn()           # second pass: seeded cylinder kernel
    for other in (20, 35, 40, 0, "seeded"):
        assert np.array_equal(res[2][0], res[other][0]) and np.array_equal(res[2][1], res[other][1])


def test_nn_100k_vs_100k_c2():
    cfg = meshgen.config_cloud(100_000, 100_000, 100, 101)
    with IcpEngine() as e:
        e.set_base(cfg["base"])
        e.set_moving(cfg["moving"])
        e.set_matrices(np.eye(4), np.eye(4))
        idx, d2 = e.nn()
    oi, od = orc.nn_bruteforce(cfg["moving"], cfg["base"])
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


def test_nn_full_size_properties_1m():
    """C3 size (1M x 1M): size-independent properties + an oracle spot check."""
    cfg = meshgen.config_cloud(1_000_000, 1_000_000, 200, 201)
    base, mov = cfg["base"], cfg["moving"]
    with IcpEngine() as e:
        e.set_base(base)
        e.set_moving(mov)
        e.set_matrices(np.eye(4), np.eye(4))
        idx, d2 = e.nn()
        idx_again, d2_again = e.nn()                          # idempotent / deterministic
        assert np.array_equal(idx, idx_again) and np.array_equal(d2, d2_again)
        # a base vertex used as a query finds itself (or an identical earlier vertex) at d2 == 0
        e.set_moving(base[::97])
        sidx, sd2 = e.nn()
        assert (sd2 == 0).all()
        assert np.array_equal(base[sidx], base[::97]) and (sidx <= np.arange(0, 1_000_000, 97)).all()
    assert idx.min() >= 0 and idx.max() < 1_000_000
    # reported distance is the S4 metric of the reported pair
    assert np.array_equal(orc._d2_spec_numpy_exact(mov, base[idx]).view(np.uint32), d2.view(np.uint32))
    # spot check 3000 queries against the CPU brute force
    sel = np.random.default_rng(0).choice(1_000_000, 3000, replace=False)
    oi, od = orc.nn_bruteforce(mov[sel], base)
    assert np.array_equal(idx[sel], oi) and np.array_equal(d2[sel], od)
    # sharding property: min over per-shard keys == unsharded keys
    keys = []
    with IcpEngine() as e:
        e.set_moving(mov[:50_000])
        e.set_matrices(np.eye(4), np.eye(4))
        for lo, hi in ((0, 300_000), (300_000, 650_001), (650_001, 1_000_000)):
            e.set_base(base[lo:hi], lo, 1_000_000)
            i, d = e.nn()
            keys.append(orc.pack_key(d, i))
    assert np.array_equal(np.minimum.reduce(keys), orc.pack_key(d2[:50_000], idx[:50_000]))


# ------------------------------------------------------------------ N2: BVH search == brute force
def _clouds_for_bvh(case):
    rng = np.random.default_rng(77)
    if case == "surface":
        cfg = meshgen.config_cloud(40_000, 60_000, 61, 62)
        return cfg["base"], cfg["moving"]
    if case == "far_outliers":            # queries far outside the base's bounding box
        cfg = meshgen.config_cloud(20_000, 50_000, 63, 64)
        mov = cfg["moving"].copy()
        mov[::3] = ((rng.random((len(mov[::3]), 3)) - 0.5) * 40).astype(np.float32)
        return cfg["base"], mov
    if case == "duplicates":              # many exact ties: lowest index must win
        base = _rand_cloud(3000, 65)
        base = np.concatenate([base, base[::-1], base[:1500]])
        return base, np.concatenate([base[::7], _rand_cloud(5000, 66)])
    if case == "collinear":               # degenerate extent (all on a line / a plane)
        t = rng.random(20_000).astype(np.float32)
        base = np.stack([t, np.zeros_like(t), np.zeros_like(t)], axis=1)
        return base, _rand_cloud(4000, 67, 0.5)
    if case == "volume":
        return _rand_cloud(70_001, 68), _rand_cloud(30_000, 69, 1.5)
    if case == "lattice":                 # regular grid: massive numbers of equidistant ties
        g = np.arange(24, dtype=np.float32)
        base = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
        mov = (base[::5] + np.float32(0.5)).astype(np.float32)
        return base, np.concatenate([mov, base[::11]])
    raise KeyError(case)


TREE_VARIANTS = [0, 13, 12, 10]      # tile-tree walk (default: persistent warps; 13: first generation), LBVH packet walk, LBVH per-thread walk


@pytest.mark.parametrize("variant", TREE_VARIANTS)
@pytest.mark.parametrize("case", ["surface", "far_outliers", "duplicates", "collinear", "volume", "lattice"])
def test_bvh_matches_bruteforce_oracle(case, variant):
    base, mov = _clouds_for_bvh(case)
    with IcpEngine(nn_mode=1, nn_variant=variant) as e:
        e.set_base(base)
        e.set_moving(mov)
        e.set_matrices(np.eye(4), np.eye(4))
        idx, d2 = e.nn()
        idx2, d22 = e.nn()                   # warm-started with the exact answer: ties must survive
    oi, od = orc.nn_bruteforce(mov, base)
    assert np.array_equal(idx, oi) and np.array_equal(idx2, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))
    assert np.array_equal(d22.view(np.uint32), od.view(np.uint32))


@pytest.mark.parametrize("n_base,n_mov", [(1, 1), (1, 300), (7, 9), (8, 8), (9, 100), (1023, 1),
                                          (1025, 1025), (4097, 2049), (33, 70), (1056, 31), (40000, 5)])
@pytest.mark.parametrize("variant", [0, 13, 12])
def test_bvh_ragged_sizes(n_base, n_mov, variant):
    base = _rand_cloud(n_base, 10 + n_base)
    mov = _rand_cloud(n_mov, 20 + n_mov)
    with IcpEngine(nn_mode=1, nn_variant=variant) as e:
        e.set_base(base)
        e.set_moving(mov)
        e.set_matrices(np.eye(4), np.eye(4))
        idx, d2 = e.nn()
        e.set_base(np.zeros((0, 3), np.float32))
        assert (e.nn()[0] == -1).all()
    oi, od = orc.nn_bruteforce(mov, base)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)


def test_bvh_equals_bruteforce_kernel_at_1m():
    """Full C3 size: every one of the 1M keys from the BVH search equals the brute-force kernel's,
    before alignment (offset clouds) -- and the sharded form composes by MIN as well."""
    cfg = meshgen.config_cloud(1_000_000, 1_000_000, 200, 201)
    res = {}
    for mode in (0, 1):
        with IcpEngine(nn_mode=mode) as e:
            e.set_base(cfg["base"])
            e.set_moving(cfg["moving"])
            e.set_matrices(np.eye(4), np.eye(4))
            res[mode] = e.nn()
            if mode == 1:
                keys = []
                for lo, hi in ((0, 400_000), (400_000, 1_000_000)):
                    e.set_base(cfg["base"][lo:hi], lo, 1_000_000)
                    i, d = e.nn()
                    keys.append(orc.pack_key(d, i))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(np.minimum.reduce(keys), orc.pack_key(res[0][1], res[0][0]))


@pytest.mark.parametrize("name", ["c1_fixed20", "c1_worldmats", "cloud_2k_3k", "c1_defaults"])
def test_bvh_trajectory_matches_bruteforce(name):
    g = Golden(name)
    p = g.prefs
    vlist = orc.build_vlist_port(g.moving.shape[0], g.groups)
    out = {}
    for mode in (0, 1):
        with IcpEngine(thresh=p["min_start"], target_d=p["target_d"], use_target=p["use_target"],
                       align_meth=p["align_meth"], stride=orc.stride_factor_port(p["sample_fraction"]),
                       nn_mode=mode) as e:
            e.set_base(g.base)
            e.set_moving(g.moving, vlist)
            e.set_matrices(g.mw_align, g.mw_base)
            done, conv, recs = e.run(p["icp_iterations"])
            out[mode] = (done, conv, e.matrix_world(), [r.M for r in recs], [r.pairs for r in recs])
    assert out[0][0] == out[1][0] == g.iterations and out[0][1] == out[1][1] == g.converged
    # same keys every iteration; the fp64 moments are summed in slot (Morton) order in BVH mode,
    # so M may differ in the last bits and the composed fp32 pose by an ulp
    np.testing.assert_allclose(out[0][2], out[1][2], rtol=0, atol=1e-6)
    assert out[0][4] == out[1][4]
    for a, b in zip(out[0][3], out[1][3]):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
    np.testing.assert_allclose(out[1][2], g.matrix_world, atol=1e-5)


# ------------------------------------------------------------------ make_pairs / solver API
def _objects(g):
    align = ms.Object("align", g.moving, g.mw_align, g.groups)
    base = ms.Object("base", g.base, g.mw_base, faces=g.base_faces)      # faces -> surface mode (N1)
    return align, base


@pytest.mark.parametrize("name", ["c1_worldmats", "c1_tight_thresh", "cloud_2k_3k", "c1_exclude"])
def test_make_pairs_bit_exact(name):
    g = Golden(name)
    align, base = _objects(g)
    vlist = oab.build_vlist(align)
    assert np.array_equal(vlist, g.vlist)
    nn = oab.VertexNN.FromObject(base, None)
    try:
        for thresh, sample, stats in ((g.prefs["min_start"], 1, True), (0.08, 2, True), (0.3, 3, False)):
            A, B, ds = oab.make_pairs(align, base, nn, vlist, thresh, sample, calc_stats=stats)
            oA, oB, ods = orc.make_pairs_vec(g.moving, vlist, g.mw_align, g.mw_base, g.base, thresh,
                                             sample, calc_stats=stats)
            assert A.shape == oA.shape and A.dtype == np.float64
            assert np.array_equal(A, oA) and np.array_equal(B, oB)
            if stats:
                np.testing.assert_allclose(ds, ods, rtol=1e-12)
            else:
                assert ds is None
        assert oab.make_pairs(align, base, nn, vlist, 0.0) is None           # general.py:277
        assert oab.make_pairs(align, base, nn, vlist, -1.0) is None
    finally:
        nn.close()


def test_affine_matrix_from_points_kat_and_errors():
    z = np.load(GOLDEN + "/solver_kat.npz")
    for tag, scale in (("rigid", False), ("sim", True), ("refl", False)):
        M = oab.affine_matrix_from_points(z[tag + "_v0"], z[tag + "_v1"], shear=False, scale=scale,
                                          usesvd=True)
        np.testing.assert_allclose(M, z[tag + "_M"], rtol=0, atol=1e-9)
    # far-from-origin data: the shifted moments must not cancel
    rng = np.random.default_rng(5)
    v0 = rng.standard_normal((3, 5000)) + np.array([[1e4], [-2e4], [3e4]])
    R = meshgen.rotation_about_axis(17.0)
    v1 = R @ v0 + np.array([[5.0], [6.0], [7.0]])
    M = oab.affine_matrix_from_points(v0, v1, shear=False, scale=False, usesvd=True)
    np.testing.assert_allclose(M, orc.affine_matrix_from_points_port(v0, v1), rtol=0, atol=1e-6)
    np.testing.assert_allclose(M[:3, :3], R, atol=1e-9)
    # planar (rank-2) and exactly 3 points
    v0 = rng.standard_normal((3, 3))
    v1 = R @ v0 + 1.0
    M = oab.affine_matrix_from_points(v0, v1, shear=False, scale=False, usesvd=True)
    np.testing.assert_allclose(M[:3, :3] @ v0 + M[:3, 3:4], v1, atol=1e-9)
    with pytest.raises(ValueError, match="wrong shape or type"):
        oab.affine_matrix_from_points(np.zeros((3, 2)), np.zeros((3, 2)), shear=False)
    with pytest.raises(ValueError):
        oab.affine_matrix_from_points(np.zeros((3, 5)), np.zeros((3, 4)), shear=False)
    with pytest.raises(NotImplementedError):
        oab.affine_matrix_from_points(np.zeros((3, 5)), np.zeros((3, 5)))    # shear=True default


# ------------------------------------------------------------------ the operator, end to end
@pytest.mark.parametrize("nn_mode", [0, 1])
@pytest.mark.parametrize("name", scenario_names())
def test_operator_execute_matches_reference_golden(name, nn_mode, monkeypatch):
    from object_alignment_b200 import functions as oab_functions
    monkeypatch.setattr(oab_functions, "DEFAULT_NN_MODE", nn_mode)
    g = Golden(name)
    align, base = _objects(g)
    oab.set_addon_preferences(oab.IcpPreferences(**g.prefs))
    ctx = ms.Context(align, [align, base])
    op = oab.OBJECT_OT_icp_align()
    assert op.poll(ctx)
    assert op.execute(ctx) == {"FINISHED"}
    assert op.iterations == g.iterations
    assert op.converged == g.converged
    assert align.rotation_mode == "QUATERNION"
    assert [r.pairs for r in op.records] == g.K.tolist()
    for it, r in enumerate(op.records):
        np.testing.assert_allclose(r.M, g.M[it], rtol=0, atol=1e-5)
        if g.prefs["use_target"]:
            np.testing.assert_allclose([r.mean_dist, r.std_dist], g.d_stats[it], rtol=1e-4)
    np.testing.assert_allclose(align.matrix_world.a, g.matrix_world, rtol=0, atol=1e-5)


@pytest.mark.parametrize("name", ["c1_fixed20", "c1_worldmats", "c1_scale", "cloud_2k_3k"])
def test_stepwise_parity_against_oracle(name):
    """Every iteration is checked against the oracle started from the engine's OWN current pose,
    so the comparison stays exact (indices, K) however many iterations are composed."""
    g = Golden(name)
    p = g.prefs
    vlist = orc.build_vlist_port(g.moving.shape[0], g.groups)
    stride = orc.stride_factor_port(p["sample_fraction"])
    with IcpEngine(thresh=p["min_start"], target_d=p["target_d"], use_target=p["use_target"],
                   align_meth=p["align_meth"], stride=stride) as e:
        e.set_base(g.base)
        e.set_moving(g.moving, vlist)
        e.set_matrices(g.mw_align, g.mw_base)
        for it in range(min(8, g.iterations)):
            mw = e.matrix_world()
            res = orc.make_pairs_vec(g.moving, vlist, mw, g.mw_base, g.base, p["min_start"], stride,
                                     calc_stats=True, return_all=True)
            idx, d2 = e.nn()
            assert np.array_equal(idx, res["idx"])
            assert np.array_equal(d2.view(np.uint32), res["d2"].view(np.uint32))
            M, st = e.step()
            assert st.pairs == int(res["keep"].sum())
            Mo = orc.affine_matrix_from_points_port(res["A"], res["B"], scale=(p["align_meth"] == "1"))
            np.testing.assert_allclose(M, Mo, rtol=0, atol=1e-10)
            np.testing.assert_allclose([st.mean_dist, st.std_dist], res["d_stats"], rtol=1e-9)
            np.testing.assert_allclose(st.trans, ms.length3(Mo.astype(np.float32)[:3, 3]), rtol=1e-6)
            np.testing.assert_allclose(st.angle, orc.rotation_angle(Mo.astype(np.float32)), atol=1e-6)
            # pose update: matrix_world @ fp32(M) with rule S3, bit-exact given the engine's M
            np.testing.assert_array_equal(e.matrix_world(), ms.matmul44(mw, M.astype(np.float32)))


@pytest.mark.parametrize("variant", [0, 13, 12])
def test_bvh_warm_start_stays_exact_over_iterations(variant):
    """BVH mode re-uses last iteration's winner as the initial bound; every iteration's keys must
    still equal the brute-force oracle evaluated at the engine's own current pose."""
    cfg = meshgen.config_cloud(50_000, 80_000, 71, 72)
    with IcpEngine(thresh=0.5, use_target=False, nn_mode=1, nn_variant=variant) as e:
        e.set_base(cfg["base"])
        e.set_moving(cfg["moving"])
        e.set_matrices(np.eye(4), np.eye(4))
        for it in range(6):
            mw = e.matrix_world()
            res = orc.make_pairs_vec(cfg["moving"], np.arange(50_000), mw, np.eye(4), cfg["base"], 0.5, 1,
                                     return_all=True, nn="kdtree")
            idx, d2 = e.nn()                     # warm-started from the previous pass
            assert np.array_equal(idx, res["idx"]) and np.array_equal(d2, res["d2"])
            M, st = e.step()
            assert st.pairs == int(res["keep"].sum())
            np.testing.assert_allclose(M, orc.affine_matrix_from_points_port(res["A"], res["B"]), atol=1e-10)


@pytest.mark.parametrize("variant", [0, 13, 12])
def test_tree_warm_start_on_a_lattice_resolves_ties_to_the_lowest_index(variant):
    """Shift the pose by exactly one lattice cell: the previous winner (the seed) is now a higher-index TIE
    of the new winner, usually in ANOTHER leaf tile.  The tile-tree walk defers the index look-up and must
    resolve such ties between tiles on the spot; the fused step() passes in between keep the seeds warm."""
    g = np.arange(28, dtype=np.float32)
    base = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)     # 21952 vertices
    mov = (base[::3] + np.float32(0.5)).astype(np.float32)
    mov = mov[(mov < 26).all(axis=1)]
    with IcpEngine(nn_mode=1, nn_variant=variant, thresh=1e9, use_target=False) as e:
        e.set_base(base)
        e.set_moving(mov)
        for k, shift in enumerate(((0, 0, 0), (-1, 0, 0), (0, 1, 0), (0, 0, -1), (1, -1, 1), (0, 0, 0))):
            mw = np.eye(4, dtype=np.float32)
            mw[:3, 3] = shift
            e.set_matrices(mw, np.eye(4))
            if k % 2 == 1:
                e.step()                              # fused kernel (seeds + winner tiles refreshed), pose moves
                e.set_matrices(mw, np.eye(4))
            idx, d2 = e.nn()
            oi, od = orc.nn_bruteforce((mov + np.array(shift, np.float32)).astype(np.float32), base)
            assert np.array_equal(idx, oi) and np.array_equal(d2, od)


@pytest.mark.parametrize("mode", ["bvh", "surface"])
def test_tree_modes_keep_slot_order_and_seeds_across_same_shape_uploads(mode):
    """icp_set_moving with the same vertex count keeps the Morton slot order and the previous winners
    (only the slot-ordered coordinates are refreshed).  A stale order and stale seeds may only cost
    speed: the same mesh re-uploaded, an unrelated cloud of the same count, and a step() in between
    must all still give the oracle's answer."""
    if mode == "bvh":
        cfg = meshgen.config_cloud(20_000, 40_000, 73, 74)
        base, mov = cfg["base"], cfg["moving"]
        set_base = lambda e: e.set_base(base)
        oracle = lambda q: orc.nn_bruteforce(q, base)[:2]
        kw = dict(nn_mode=1)
    else:
        v, t = meshgen.bunny_like_mesh(8_000, 75)
        tri = meshgen.triangle_corners(v, t)
        mov = meshgen.apply_rigid(meshgen.bunny_like(6_000, 76, noise=2e-3), 3.0, (0.02, -0.01, 0.03))
        set_base = lambda e: e.set_base_triangles(tri)
        oracle = lambda q: orc.nn_surface_bruteforce(q, tri)[:2]
        kw = dict(nn_mode=2)
    other = _rand_cloud(mov.shape[0], 77, 0.8)
    with IcpEngine(thresh=1e9, use_target=False, **kw) as e:
        set_base(e)
        e.set_moving(mov)
        e.set_matrices(np.eye(4), np.eye(4))
        for q in (mov, mov, other, mov):
            e.set_moving(q)                          # same count: order and seeds are kept
            e.set_matrices(np.eye(4), np.eye(4))
            idx, d2 = e.nn()
            oi, od = oracle(q)
            assert np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od.view(np.uint32))
            e.step()                                 # fused iteration on the refreshed slots
        e.set_moving(other)
        e.set_matrices(np.eye(4), np.eye(4))
        M, st = e.step()                             # fused kernel straight after a stale-order upload
        assert st.pairs == other.shape[0]
        e.set_matrices(np.eye(4), np.eye(4))
        idx, d2 = e.nn()
        oi, od = oracle(other)
        assert np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od.view(np.uint32))


@pytest.mark.parametrize("hard_class", ["1", "3", "16"])
def test_q16_class_order_and_gated_filter_on_a_small_cloud_with_outliers(monkeypatch, hard_class):
    """The q16 kernel walks the slots in class order (similar pruning radius together) and leaves the hard classes
    to the gated expansion-filter launch; by default only from 2^18 slots on.  Forced here on a 40k cloud with 25 %
    far outliers (ICP_Q16_CLASS_MIN=0) for three settings of the hard class: every pass must still return the
    oracle's keys, cold and seeded, and the percentile cut must keep the same set."""
    monkeypatch.setenv("ICP_Q16_CLASS_MIN", "0")
    monkeypatch.setenv("ICP_Q16_HARD_CLASS", hard_class)
    cfg = meshgen.config_cloud(40_000, 60_000, 81, 82)
    mov = cfg["moving"].copy()
    rng = np.random.default_rng(83)
    sel = rng.choice(40_000, 10_000, replace=False)
    mov[sel] = ((rng.random((10_000, 3)) - 0.5) * 4).astype(np.float32)
    with IcpEngine(thresh=1e9, use_target=False, reject_fraction=0.25) as e:
        e.set_base(cfg["base"])
        e.set_moving(mov)
        e.set_matrices(np.eye(4), np.eye(4))
        for it in range(4):
            mw = e.matrix_world()
            res = orc.make_pairs_vec(mov, np.arange(40_000), mw, np.eye(4), cfg["base"], 1e9, 1, reject_fraction=0.25,
                                     return_all=True, nn="kdtree")
            idx, d2 = e.nn()
            assert np.array_equal(idx, res["idx"]) and np.array_equal(d2.view(np.uint32), res["d2"].view(np.uint32))
            M, st = e.step()
            assert st.pairs == int(res["keep"].sum())


@pytest.mark.parametrize("variant", [0, 20, 2, 30, 33, 35, 36, 40, 41, 42, 44, 45, 50, 51, 52, 53, 54, 55, 56, 57,
                                     58, 59, 60, 61, 62, 63, 64, 65, 66, 67, 68, 69, 70, 71, 72, 73, 77, 78, 79])
def test_bruteforce_seeded_passes_stay_exact(variant):
    """From the second pass on the brute-force filter kernel starts its running minimum from the
    previous winner re-evaluated under the new pose.  Every pass must still return the oracle's
    keys: over converging iterations, after a wild pose jump (useless seeds), after the moving
    vertices are replaced by unrelated ones of the same count (stale seeds), and on a lattice
    where the seed becomes a higher-index TIE of the new winner."""
    cfg = meshgen.config_cloud(30_000, 80_000, 71, 72)
    n_mov = cfg["moving"].shape[0]
    with IcpEngine(thresh=0.5, use_target=False, nn_variant=variant) as e:
        e.set_base(cfg["base"])
        e.set_moving(cfg["moving"])
        e.set_matrices(np.eye(4), np.eye(4))
        for it in range(5):
            mw = e.matrix_world()
            res = orc.make_pairs_vec(cfg["moving"], np.arange(n_mov), mw, np.eye(4), cfg["base"], 0.5, 1,
                                     return_all=True, nn="kdtree")
            idx, d2 = e.nn()
            assert np.array_equal(idx, res["idx"]) and np.array_equal(d2, res["d2"])
            M, st = e.step()                                        # seeded as well
            assert st.pairs == int(res["keep"].sum())
        jump = meshgen.rigid_matrix(171.0, (0.4, -0.3, 0.25), axis=(-1.0, 0.2, 0.4))
        e.set_matrices(jump, np.eye(4))
        res = orc.make_pairs_vec(cfg["moving"], np.arange(n_mov), jump, np.eye(4), cfg["base"], 1e9, 1,
                                 return_all=True, nn="kdtree")
        idx, d2 = e.nn()
        assert np.array_equal(idx, res["idx"]) and np.array_equal(d2, res["d2"])
        other = _rand_cloud(n_mov, 99, 0.7)
        e.set_moving(other)
        e.set_matrices(np.eye(4), np.eye(4))
        idx, d2 = e.nn()
        oi, od = orc.nn_bruteforce(other, cfg["base"])
        assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    # lattice: shift the pose by exactly one cell, the old winner is now a tie with a higher index
    g = np.arange(28, dtype=np.float32)
    base = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)     # 21952 vertices
    mov = (base[::3] + np.float32(0.5)).astype(np.float32)
    mov = mov[(mov < 26).all(axis=1)]
    with IcpEngine(nn_variant=variant) as e:
        e.set_base(base)
        e.set_moving(mov)
        for shift in ((0, 0, 0), (-1, 0, 0), (0, 1, 0), (0, 0, -1), (1, -1, 1)):
            mw = np.eye(4, dtype=np.float32)
            mw[:3, 3] = shift
            e.set_matrices(mw, np.eye(4))
            idx, d2 = e.nn()
            oi, od = orc.nn_bruteforce((mov + np.array(shift, np.float32)).astype(np.float32), base)
            assert np.array_equal(idx, oi) and np.array_equal(d2, od)


def test_reference_shaped_loop_with_icp_step():
    """The reference's own loop shape (icp_align.py:96-151), one fused call per iteration."""
    g = Golden("c1_fixed20")
    align, base = _objects(g)
    vlist = oab.build_vlist(align)
    nn = oab.VertexNN.FromObject(base, None)
    try:
        idx, dist = oab.calc_target(align, base, nn, vlist, 1)
        assert np.array_equal(idx, g.idx[0])
        for n in range(g.iterations):
            M, st = oab.icp_step(align, base, nn, vlist, 0.5, 1, calc_stats=False)
            np.testing.assert_allclose(M, g.M[n], atol=1e-5)
        np.testing.assert_allclose(align.matrix_world.a, g.matrix_world, atol=1e-5)
        with pytest.raises(TypeError):
            oab.icp_step(align, base, nn, vlist, 0.0)
        with pytest.raises(ValueError, match="wrong shape or type"):
            oab.icp_step(align, base, nn, vlist, 1e-9)          # nothing survives the threshold
        # the handle stays usable after the error
        M, st = oab.icp_step(align, base, nn, vlist, 0.5, 1)
        assert st.pairs == 1002
    finally:
        nn.close()


@pytest.mark.parametrize("name", ["c1_fixed20", "c1_defaults", "c1_include", "surf_c1_defaults"])
def test_modal_operator_matches_reference_golden(name):
    """Row N3: the timer-driven operator reaches the same pose in the same number of iterations."""
    g = Golden(name)
    align, base = _objects(g)
    prefs = dict(g.prefs, redraw_frequency=3)
    oab.set_addon_preferences(oab.IcpPreferences(**prefs))
    ctx = ms.Context(align, [align, base])
    op = oab.OBJECT_OT_icp_align_feedback()
    assert op.poll(ctx)
    assert op.invoke(ctx, ms.Event("NONE")) == {"RUNNING_MODAL"}
    assert op.modal(ctx, ms.Event("MOUSEMOVE")) == {"PASS_THROUGH"}
    ticks = 0
    while True:
        ticks += 1
        r = op.modal(ctx, ms.Event("TIMER"))
        if r != {"RUNNING_MODAL"}:
            break
        assert ticks < 100
    assert r == {"FINISHED"}
    assert op.total_iters == g.iterations and op.converged == g.converged
    assert ctx.area.redraws == ticks and ctx.window_manager.timers == []
    np.testing.assert_allclose(align.matrix_world.a, g.matrix_world, rtol=0, atol=1e-5)
    # ESC cancels and releases the engine
    align2, base2 = _objects(g)
    ctx2 = ms.Context(align2, [align2, base2])
    op2 = oab.OBJECT_OT_icp_align_feedback()
    op2.invoke(ctx2, ms.Event("NONE"))
    op2.modal(ctx2, ms.Event("TIMER"))
    assert op2.modal(ctx2, ms.Event("ESC")) == {"CANCELLED"} and op2.base_nn is None


# ------------------------------------------------------------------ N1: closest point on the surface
def _surface_case(case):
    rng = np.random.default_rng(91)
    if case == "icosphere":
        v, t = meshgen.icosphere(12, return_faces=True)
        q = (v[rng.integers(0, len(v), 4000)] * (1 + 0.05 * rng.standard_normal((4000, 1)))).astype(np.float32)
    elif case == "bunny":
        v, t = meshgen.bunny_like_mesh(20_000, 92)
        q = meshgen.apply_rigid(meshgen.bunny_like(15_000, 93, noise=2e-3), 3.0, (0.02, -0.01, 0.03))
    elif case == "far":
        v, t = meshgen.bunny_like_mesh(6000, 94)
        q = ((rng.random((3000, 3)) - 0.5) * 12).astype(np.float32)
    elif case == "on_vertices":        # queries exactly on mesh vertices: d2 == 0, many tied triangles
        v, t = meshgen.icosphere(8, return_faces=True)
        q = v.copy()
    elif case == "single_triangle":
        v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
        t = np.array([[0, 1, 2]], np.int32)
        q = ((rng.random((500, 3)) - 0.3) * 2).astype(np.float32)
    return v, t, q


@pytest.mark.parametrize("case", ["icosphere", "bunny", "far", "on_vertices", "single_triangle"])
def test_surface_search_matches_oracle(case):
    v, t, q = _surface_case(case)
    tri = meshgen.triangle_corners(v, t)
    with IcpEngine(nn_mode=2, thresh=1e9) as e:
        e.set_base_triangles(tri)
        e.set_moving(q)
        e.set_matrices(np.eye(4), np.eye(4))
        idx, d2 = e.nn()
        idx2, d22 = e.nn()                       # warm-started second pass
        A, B, ds = e.make_pairs(calc_stats=True)
    oi, od, oc = orc.nn_surface_bruteforce(q, tri)
    assert np.array_equal(idx, oi) and np.array_equal(idx2, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32)) and np.array_equal(d22, od)
    assert np.array_equal(B, oc.astype(np.float64).T)            # targets = closest points (identity poses)
    assert np.array_equal(A, q.astype(np.float64).T)


@pytest.mark.parametrize("name", ["surf_c1_fixed12", "surf_c1_worldmats", "surf_bunny_1k"])
def test_surface_stepwise_parity(name):
    g = Golden(name)
    p = g.prefs
    tri = g.base[g.base_faces].reshape(-1, 9)
    vlist = orc.build_vlist_port(g.moving.shape[0], g.groups)
    stride = orc.stride_factor_port(p["sample_fraction"])
    with IcpEngine(thresh=p["min_start"], target_d=p["target_d"], use_target=p["use_target"],
                   align_meth=p["align_meth"], stride=stride, nn_mode=2) as e:
        e.set_base_triangles(tri)
        e.set_moving(g.moving, vlist)
        e.set_matrices(g.mw_align, g.mw_base)
        for it in range(min(6, g.iterations)):
            mw = e.matrix_world()
            res = orc.make_pairs_vec(g.moving, vlist, mw, g.mw_base, g.base, p["min_start"], stride,
                                     calc_stats=True, return_all=True, tri_xyz=tri)
            idx, d2 = e.nn()
            assert np.array_equal(idx, res["idx"]) and np.array_equal(d2, res["d2"])
            A, B, ds = e.make_pairs(calc_stats=True)
            assert np.array_equal(A, res["A"]) and np.array_equal(B, res["B"])
            M, st = e.step()
            assert st.pairs == int(res["keep"].sum())
            np.testing.assert_allclose(M, orc.affine_matrix_from_points_port(res["A"], res["B"]), atol=1e-10)


def test_surface_packet_walk_equals_per_thread_walk_and_fused_iteration():
    """200k queries vs a 100k-vertex / 200k-triangle mesh: the deferred warp-packet walk (default), the
    eager packet walk (nn_variant 11), the per-thread walk (nn_variant 10) and the fused kernels agree -- same keys bit for bit,
    same kept-pair counts every iteration, poses equal up to the fp64 summation order."""
    v, t = meshgen.bunny_like_mesh(100_000, 95)
    tri = meshgen.triangle_corners(v, t)
    q = meshgen.apply_rigid(meshgen.bunny_like(200_000, 96, noise=2e-3), 4.0, (0.02, -0.01, 0.03))
    rng = np.random.default_rng(97)
    q[rng.choice(200_000, 2000, replace=False)] = ((rng.random((2000, 3)) - 0.5) * 5).astype(np.float32)  # outliers
    out = {}
    for variant in (0, 14, 11, 10):
        with IcpEngine(nn_mode=2, thresh=0.5, use_target=False, nn_variant=variant) as e:
            e.set_base_triangles(tri)
            e.set_moving(q)
            e.set_matrices(np.eye(4), np.eye(4))
            idx, d2 = e.nn()
            done, conv, recs = e.run(6)              # 0: deferred walk, one launch per iteration (ticketed tail); 14: + separate reduce / solve; 11: fused eager walk; 10: K1 + per-thread walk + K3/K4
            idx_after, d2_after = e.nn()             # warm-started search under the final pose
            out[variant] = (idx, d2, [r.pairs for r in recs], e.matrix_world(), idx_after, d2_after)
    a, b = out[0], out[10]
    for other in (out[14], out[11], out[10]):
        assert np.array_equal(a[0], other[0]) and np.array_equal(a[1].view(np.uint32), other[1].view(np.uint32))
        assert a[2] == other[2]
        np.testing.assert_allclose(a[3], other[3], atol=2e-6)
    # spot-check the packet walk against the CPU oracle
    sel = rng.choice(200_000, 1500, replace=False)
    oi, od, _ = orc.nn_surface_bruteforce(q[sel], tri)
    assert np.array_equal(a[0][sel], oi) and np.array_equal(a[1][sel].view(np.uint32), od.view(np.uint32))
    if np.array_equal(a[3], b[3]):                   # identical poses -> identical final searches
        assert np.array_equal(a[4], b[4]) and np.array_equal(a[5].view(np.uint32), b[5].view(np.uint32))


def test_surface_mode_needs_triangles():
    with IcpEngine(nn_mode=2) as e:
        e.set_base(_rand_cloud(100, 1))
        e.set_moving(_rand_cloud(10, 2))
        e.set_matrices(np.eye(4), np.eye(4))
        with pytest.raises(IcpError) as ei:
            e.nn()
        assert ei.value.code == 6
    base_cloud = ms.Object("b", _rand_cloud(50, 3))                   # no faces -> vertex index
    nn = oab.base_index_from_object(base_cloud)
    assert type(nn) is oab.VertexNN
    nn.close()
    v, t = meshgen.icosphere(2, return_faces=True)
    nn = oab.base_index_from_object(ms.Object("m", v, faces=t))
    assert type(nn) is oab.SurfaceNN
    nn.close()


# ------------------------------------------------------------------ percentile rejection (new)
@pytest.mark.parametrize("rf", [0.25, 0.5, 0.013])
def test_reject_fraction_matches_definition(rf):
    cfg = meshgen.config_cloud(20_000, 30_000, 11, 12)
    mov = cfg["moving"].copy()
    rng = np.random.default_rng(13)
    out = rng.choice(20_000, 5000, replace=False)
    mov[out] = (rng.random((5000, 3)).astype(np.float32) - 0.5) * 6
    mov[7] = mov[3]                                   # equal d2 at different positions
    with IcpEngine(thresh=1e9, reject_fraction=rf, use_target=True) as e:
        e.set_base(cfg["base"])
        e.set_moving(mov)
        e.set_matrices(np.eye(4), np.eye(4))
        A, B, ds = e.make_pairs(calc_stats=True)
        M, st = e.step()
    res = orc.make_pairs_vec(mov, np.arange(20_000), np.eye(4), np.eye(4), cfg["base"], 1e9, 1,
                             calc_stats=True, reject_fraction=rf, return_all=True)
    assert A.shape[1] == 20_000 - int(np.floor(rf * 20_000)) == int(res["keep"].sum())
    assert np.array_equal(A, res["A"]) and np.array_equal(B, res["B"])
    np.testing.assert_allclose(ds, res["d_stats"], rtol=1e-12)
    assert st.pairs == A.shape[1]
    np.testing.assert_allclose(M, orc.affine_matrix_from_points_port(res["A"], res["B"]), atol=1e-10)


# ------------------------------------------------------------------ errors and misc
def test_error_behaviour():
    with IcpEngine() as e:
        with pytest.raises(IcpError) as ei:
            e.nn()
        assert ei.value.code == 6                                 # ICP_ERR_STATE
        e.set_base(_rand_cloud(100, 1))
        e.set_moving(_rand_cloud(50, 2))
        sing = np.eye(4, dtype=np.float32)
        sing[2, 2] = 0.0
        e.set_matrices(sing, np.eye(4))
        with pytest.raises(IcpError) as ei:
            e.step()
        assert ei.value.code == 7                                 # Matrix.inverted() raises
        e.set_matrices(np.eye(4), np.eye(4))
        e.update_params(thresh=0.0)
        with pytest.raises(IcpError) as ei:
            e.step()
        assert ei.value.code == 1
        e.update_params(thresh=10.0)
        M, st = e.step()
        assert st.pairs == 50
        with pytest.raises(IcpError):
            e.set_moving(_rand_cloud(5, 1), np.array([0, 7], np.int32))   # vlist out of range
    with pytest.raises(ZeroDivisionError):
        oab.stride_from_sample_fraction(0)


def test_run_is_deterministic_and_device_resident():
    cfg = meshgen.config_cloud(30_000, 40_000, 31, 32)
    outs = []
    for _ in range(2):
        with IcpEngine(thresh=0.5, use_target=False) as e:
            e.set_base(cfg["base"])
            e.set_moving(cfg["moving"])
            e.set_matrices(np.eye(4), np.eye(4))
            done, conv, recs = e.run(6)
            assert done == 6 and not conv
            outs.append((e.matrix_world(), [r.M for r in recs]))
    assert np.array_equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert np.array_equal(a, b)
    # the pose converges towards the inverse of the 5 degree / (0.02,-0.01,0.03) offset
    res = orc.icp_execute_port(cfg["moving"], np.arange(30_000), np.eye(4), np.eye(4), cfg["base"],
                               thresh=0.5, sample_fraction=1.0, iters=6, use_target=False, nn="kdtree")
    np.testing.assert_allclose(outs[0][0], res["matrix_world"], atol=1e-5)
