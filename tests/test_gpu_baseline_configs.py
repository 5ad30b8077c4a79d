"""GPU parity at the BASELINE.json configurations themselves (C3, C4 shape, C5 shape), through the C ABI.

The kernel bench.py times -- the seeded brute-force pass with the pose changing every step -- is
checked on ALL queries against the CPU oracle (``make_pairs_vec(nn="kdtree")``: scipy cKDTree
candidates re-ranked with the exact S4 metric, lowest index on ties) at the engine's own pose, so the
comparison is exact however many iterations are composed.  Reference lines restated by the oracle:
functions/general.py:274-306 (pairing, threshold), operators/icp_align.py:56-80 (mask rule).
"""
import numpy as np
import pytest

import object_alignment_b200 as oab
from object_alignment_b200 import meshgen
from object_alignment_b200.engine import IcpEngine
from oracle import icp_oracle as orc

pytestmark = pytest.mark.gpu

IDENT = np.eye(4, dtype=np.float32)


def _oracle_at(cfg, vlist, mw, stride=1, thresh=0.5, reject_fraction=0.0):
    return orc.make_pairs_vec(cfg["moving"], vlist, mw, IDENT, cfg["base"], thresh, stride, calc_stats=True,
                              reject_fraction=reject_fraction, return_all=True, nn="kdtree")


def test_c3_seeded_passes_equal_oracle_on_all_queries_under_pose_changes():
    """C3 (1M x 1M): cold pass, then three seeded passes each under a new pose (the ICP's own updates),
    every one of the 1M keys against the oracle and against the BVH search; K, d_stats and M of every
    step against make_pairs_vec + the solver port."""
    cfg = meshgen.config_cloud(1_000_000, 1_000_000, 200, 201)
    m = cfg["moving"].shape[0]
    vl = np.arange(m)
    with IcpEngine(thresh=0.5, use_target=True, target_d=0.0) as e, \
            IcpEngine(thresh=0.5, use_target=False, nn_mode=1) as tree:
        e.set_base(cfg["base"]); e.set_moving(cfg["moving"]); e.set_matrices(IDENT, IDENT)
        tree.set_base(cfg["base"]); tree.set_moving(cfg["moving"])
        for it in range(4):                      # it == 0: cold kernel; 1..3: seeded kernel, pose changed by step()
            mw = e.matrix_world()
            res = _oracle_at(cfg, vl, mw)
            idx, d2 = e.nn()
            assert np.array_equal(idx, res["idx"]), "pass %d: %d indices differ" % (it, int((idx != res["idx"]).sum()))
            assert np.array_equal(d2.view(np.uint32), res["d2"].view(np.uint32))
            tree.set_matrices(mw, IDENT)
            tidx, td2 = tree.nn()
            assert np.array_equal(tidx, idx) and np.array_equal(td2.view(np.uint32), d2.view(np.uint32))
            M, st = e.step()                     # seeded again (same pose), then the pose moves
            assert st.pairs == int(res["keep"].sum())
            np.testing.assert_allclose([st.mean_dist, st.std_dist], res["d_stats"], rtol=1e-9)
            np.testing.assert_allclose(M, orc.affine_matrix_from_points_port(res["A"], res["B"]), rtol=0, atol=1e-9)


def test_c4_shape_seeded_bruteforce_equals_bvh_and_oracle_spot_check():
    """C4 shape (1M moving vs 10M base): the seeded brute-force keys after a pose change equal the BVH
    search's on all 1M queries; 3000 of them against the CPU brute force over the full 10M base."""
    cfg = meshgen.config_cloud(1_000_000, 10_000_000, 300, 301)
    with IcpEngine(thresh=0.5, use_target=False) as e:
        e.set_base(cfg["base"]); e.set_moving(cfg["moving"]); e.set_matrices(IDENT, IDENT)
        e.step()                                 # cold pass + pose update
        e.step()                                 # seeded pass + pose update
        mw = e.matrix_world()
        idx, d2 = e.nn()                         # seeded pass under the new pose
    with IcpEngine(thresh=0.5, use_target=False, nn_mode=1) as tree:
        tree.set_base(cfg["base"]); tree.set_moving(cfg["moving"]); tree.set_matrices(mw, IDENT)
        tidx, td2 = tree.nn()
    assert np.array_equal(idx, tidx) and np.array_equal(d2.view(np.uint32), td2.view(np.uint32))
    sel = np.sort(np.random.default_rng(4).choice(1_000_000, 3000, replace=False))
    res = orc.make_pairs_vec(cfg["moving"], sel, mw, IDENT, cfg["base"], 0.5, 1, return_all=True, nn="brute")
    assert np.array_equal(idx[sel], res["idx"]) and np.array_equal(d2[sel].view(np.uint32), res["d2"].view(np.uint32))


@pytest.mark.parametrize("nn_mode", [0, 1])
def test_c5_shape_outliers_mask_reject_stride(nn_mode):
    """C5 shape at 600k x 600k: 25 % outliers, the icp_include vertex-group rule (icp_align.py:56-80),
    reject_fraction 0.25 and stride 2 together -- kept set, A, B, K, d_stats and M against the oracle,
    on the cold pass and on a seeded pass under the next pose."""
    cfg = meshgen.config_c5(600_000, "include")
    vlist = oab.build_vlist_from_arrays(cfg["moving"].shape[0], cfg["groups"])
    assert np.array_equal(vlist, orc.build_vlist_port(cfg["moving"].shape[0], cfg["groups"]))
    assert 0.55 < vlist.shape[0] / 600_000 < 0.65
    thresh = 1e9                                  # C5 rejects by percentile only
    with IcpEngine(thresh=thresh, use_target=True, target_d=0.0, stride=2, reject_fraction=0.25, nn_mode=nn_mode) as e:
        e.set_base(cfg["base"]); e.set_moving(cfg["moving"], vlist); e.set_matrices(IDENT, IDENT)
        for it in range(3):
            mw = e.matrix_world()
            res = _oracle_at(cfg, vlist, mw, stride=2, thresh=thresh, reject_fraction=0.25)
            L = res["idx"].shape[0]
            assert L == (vlist.shape[0] + 1) // 2
            idx, d2 = e.nn()
            assert np.array_equal(idx, res["idx"]) and np.array_equal(d2.view(np.uint32), res["d2"].view(np.uint32))
            A, B, ds = e.make_pairs(calc_stats=True)
            assert A.shape[1] == L - int(np.floor(0.25 * L)) == int(res["keep"].sum())
            assert np.array_equal(A, res["A"]) and np.array_equal(B, res["B"])
            np.testing.assert_allclose(ds, res["d_stats"], rtol=1e-12)
            M, st = e.step()
            assert st.pairs == A.shape[1]
            np.testing.assert_allclose([st.mean_dist, st.std_dist], res["d_stats"], rtol=1e-9)
            np.testing.assert_allclose(M, orc.affine_matrix_from_points_port(res["A"], res["B"]), rtol=0, atol=1e-9)


@pytest.mark.parametrize("nn_mode", [0, 1])
def test_nn_and_make_pairs_after_convergence_use_the_new_pose(nn_mode):
    """A run that converged latches the device-side stop flag; stand-alone queries afterwards
    (icp_nn, icp_make_pairs -- the function surface's calc_target / make_pairs) must still search under
    whatever pose and vertex list the caller set since."""
    cfg = meshgen.config_cloud(20_000, 30_000, 41, 42)
    m = cfg["moving"].shape[0]
    with IcpEngine(thresh=0.5, use_target=True, target_d=0.05, nn_mode=nn_mode) as e:
        e.set_base(cfg["base"]); e.set_moving(cfg["moving"]); e.set_matrices(IDENT, IDENT)
        done, conv, recs = e.run(50)
        assert conv and done < 50
        jump = meshgen.rigid_matrix(25.0, (0.1, -0.2, 0.05), axis=(0.2, 1.0, -0.3))
        e.set_matrices(jump, IDENT)
        res = orc.make_pairs_vec(cfg["moving"], np.arange(m), jump, IDENT, cfg["base"], 0.5, 1, calc_stats=True,
                                 return_all=True)
        idx, d2 = e.nn()
        assert np.array_equal(idx, res["idx"]) and np.array_equal(d2.view(np.uint32), res["d2"].view(np.uint32))
        A, B, ds = e.make_pairs(calc_stats=True)
        assert np.array_equal(A, res["A"]) and np.array_equal(B, res["B"])
        # fewer queries than before (new stride): the buffers are re-sized, nothing stale is read
        vl = np.arange(0, m, 3, dtype=np.int32)
        e.set_moving(cfg["moving"], vl)
        res = orc.make_pairs_vec(cfg["moving"], vl, jump, IDENT, cfg["base"], 0.5, 1, return_all=True)
        idx, d2 = e.nn()
        assert np.array_equal(idx, res["idx"]) and np.array_equal(d2.view(np.uint32), res["d2"].view(np.uint32))
        # the run itself is still latched: icp_run does nothing more until the caller resets it
        done2, conv2, _ = e.run(3)
        assert done2 == 0 and conv2
        e.reset_convergence()
        done3, _, _ = e.run(2)
        assert done3 == 2


def test_set_moving_rejects_bad_vlist_without_touching_the_handle():
    base = meshgen.bunny_like(5000, 1)
    mov = meshgen.bunny_like(3000, 2)
    with IcpEngine(thresh=0.5) as e:
        e.set_base(base); e.set_moving(mov, np.arange(0, 3000, 2, dtype=np.int32)); e.set_matrices(IDENT, IDENT)
        idx0, d0 = e.nn()
        with pytest.raises(oab.IcpError):
            e.set_moving(mov[:100], np.array([5, 250], np.int32))       # 250 is out of range for 100 vertices
        assert e.query_count() == 1500                                   # previous upload still in force
        idx1, d1 = e.nn()
        assert np.array_equal(idx0, idx1) and np.array_equal(d0, d1)


def test_function_surface_rereads_coordinates_every_call():
    """The reference reads vert.co on every make_pairs call (general.py:284): editing the align mesh in
    place between two calls must change the answer without any invalidate()."""
    from oracle import mathutils_shim as ms
    base = meshgen.bunny_like(8000, 3)
    mov = meshgen.bunny_like(2000, 4, noise=1e-3)
    align = ms.Object("align", mov.copy())
    bobj = ms.Object("base", base)
    vlist = list(range(2000))
    nn = oab.VertexNN.FromObject(bobj, None)
    try:
        i0, _ = oab.calc_target(align, bobj, nn, vlist, 1)
        moved = (mov[::-1] * np.float32(0.9)).astype(np.float32)
        align2 = ms.Object("align", moved)                             # same vertex count, new coordinates
        i1, dist1 = oab.calc_target(align2, bobj, nn, vlist, 1)
        oi, od = orc.nn_bruteforce(moved, base)
        assert np.array_equal(i1, oi) and not np.array_equal(i0, i1)
    finally:
        nn.close()


def test_tree_iteration_with_a_four_level_ticket_tree():
    """1.2M query slots = 37 500 batches of 32: the ticketed reduction of the one-launch tree iteration needs FOUR
    levels here (32^3 = 32 768 batches fill three).  K, d_stats and M of two steps against the oracle at the
    engine's own pose, and the keys against the brute-force engine."""
    cfg = meshgen.config_cloud(1_200_000, 200_000, 410, 411)
    m = cfg["moving"].shape[0]
    with IcpEngine(thresh=0.5, use_target=True, target_d=0.0, nn_mode=1) as e:
        e.set_base(cfg["base"]); e.set_moving(cfg["moving"]); e.set_matrices(IDENT, IDENT)
        for it in range(2):
            mw = e.matrix_world()
            res = _oracle_at(cfg, np.arange(m), mw)
            idx, d2 = e.nn()
            assert np.array_equal(idx, res["idx"]) and np.array_equal(d2.view(np.uint32), res["d2"].view(np.uint32))
            M, st = e.step()
            assert st.pairs == int(res["keep"].sum())
            np.testing.assert_allclose([st.mean_dist, st.std_dist], res["d_stats"], rtol=1e-9)
            np.testing.assert_allclose(M, orc.affine_matrix_from_points_port(res["A"], res["B"]), rtol=0, atol=1e-9)
