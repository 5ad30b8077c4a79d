"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run here (the container that has /root/reference):  python -m oracle.gen_golden

For every scenario the reference's own ``OBJECT_OT_icp_align.execute``
(/root/reference/operators/icp_align.py:47-161) is driven over duck-typed objects
(oracle/mathutils_shim.py) with ``BVHTree`` replaced by the nearest-vertex provider
(SURVEY.md D2).  ``make_pairs`` / ``affine_matrix_from_points`` are observed through
recording wrappers placed in the operator module's namespace -- the reference sources are
not modified or copied.  What is stored per iteration: the winning base index of every
query, K, d_stats, the 4x4 M; plus the final matrix_world, the iteration count, and the
inputs themselves so the fixture is self-contained on the GPU box.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from object_alignment_b200 import meshgen  # noqa: E402
from oracle import icp_oracle, mathutils_shim as ms, ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _groups_arrays(n, seed, kind):
    rng = np.random.default_rng(seed)
    member = rng.random(n) < 0.6
    w = np.where(rng.random(n) < 0.7, 1.0, rng.random(n)).astype(np.float32)
    if kind == "exclude":
        w = np.where(rng.random(n) < 0.4, rng.random(n) * 0.2, w).astype(np.float32)
    return member, w


def scenarios():
    c1 = meshgen.config_c1()
    base_w = meshgen.rigid_matrix(20.0, (0.3, -0.2, 0.1), axis=(0.2, -1.0, 0.4)).astype(np.float64)
    base_w[:3, :3] = base_w[:3, :3] @ np.diag([1.3, 0.8, 1.1])        # non-uniform scale
    base_w = base_w.astype(np.float32)
    # moving vertices placed so that world positions still overlap the transformed base
    mv_world = (c1["moving"].astype(np.float64) @ base_w[:3, :3].astype(np.float64).T
                + base_w[:3, 3].astype(np.float64))
    align_w = meshgen.rigid_matrix(-7.0, (0.1, 0.05, -0.02), axis=(1.0, 0.3, 0.2))
    inv_aw = np.linalg.inv(align_w.astype(np.float64))
    mv_local = (mv_world @ inv_aw[:3, :3].T + inv_aw[:3, 3]).astype(np.float32)
    cloud = meshgen.config_cloud(2000, 3000, 100, 101)
    n = c1["moving"].shape[0]
    inc = _groups_arrays(n, 7, "include")
    exc = _groups_arrays(n, 8, "exclude")
    S = []
    S.append(dict(name="c1_fixed20", moving=c1["moving"], base=c1["base"],
                  prefs=dict(sample_fraction=1.0, icp_iterations=20, use_target=False)))
    S.append(dict(name="c1_defaults", moving=c1["moving"], base=c1["base"], prefs=dict()))
    S.append(dict(name="c1_scale", moving=(c1["moving"] * np.float32(1.06)), base=c1["base"],
                  prefs=dict(sample_fraction=1.0, icp_iterations=12, use_target=True,
                             target_d=0.001, align_meth="1")))
    S.append(dict(name="c1_include", moving=c1["moving"], base=c1["base"],
                  groups={"icp_include": inc},
                  prefs=dict(sample_fraction=1.0, icp_iterations=8, use_target=False)))
    S.append(dict(name="c1_exclude", moving=c1["moving"], base=c1["base"],
                  groups={"icp_exclude": exc},
                  prefs=dict(sample_fraction=0.4, icp_iterations=8, use_target=False)))
    S.append(dict(name="c1_include_wins", moving=c1["moving"], base=c1["base"],
                  groups={"icp_exclude": exc, "icp_include": inc},
                  prefs=dict(sample_fraction=1.0, icp_iterations=3, use_target=False)))
    S.append(dict(name="c1_worldmats", moving=mv_local, base=c1["base"],
                  mw_align=align_w, mw_base=base_w,
                  prefs=dict(sample_fraction=1.0, icp_iterations=10, use_target=True,
                             min_start=0.35, target_d=0.002)))
    S.append(dict(name="c1_tight_thresh", moving=c1["moving"], base=c1["base"],
                  prefs=dict(sample_fraction=1.0, icp_iterations=10, use_target=True,
                             min_start=0.06, target_d=0.0)))
    # ---- surface scenarios (row N1): the base HAS faces, the provider answers with the closest
    #      point on the triangulated surface -- the reference's real semantics
    sv, st = meshgen.icosphere(10, return_faces=True)
    S.append(dict(name="surf_c1_fixed12", moving=c1["moving"], base=sv, base_faces=st,
                  prefs=dict(sample_fraction=1.0, icp_iterations=12, use_target=False)))
    S.append(dict(name="surf_c1_defaults", moving=c1["moving"], base=sv, base_faces=st, prefs=dict()))
    S.append(dict(name="surf_c1_worldmats", moving=mv_local, base=sv, base_faces=st,
                  mw_align=align_w, mw_base=base_w,
                  prefs=dict(sample_fraction=1.0, icp_iterations=8, use_target=True,
                             min_start=0.35, target_d=0.002)))
    bv, bt = meshgen.bunny_like_mesh(1500, 110)
    bm = meshgen.apply_rigid(meshgen.bunny_like(1200, 111, noise=2e-3), 4.0, (0.02, -0.01, 0.03))
    S.append(dict(name="surf_bunny_1k", moving=bm, base=bv, base_faces=bt,
                  prefs=dict(sample_fraction=1.0, icp_iterations=10, use_target=False, min_start=0.5)))
    S.append(dict(name="cloud_2k_3k", moving=cloud["moving"], base=cloud["base"],
                  prefs=dict(sample_fraction=1.0, icp_iterations=15, use_target=False,
                             min_start=0.5)))
    return S


def run_reference(sc):
    general, icp_align = ref_loader.load_reference()
    align = ms.Object("align", sc["moving"], sc.get("mw_align"), sc.get("groups"))
    base = ms.Object("base", sc["base"], sc.get("mw_base"), faces=sc.get("base_faces"))
    prefs = ms.Settings(**sc["prefs"])
    from mathutils import bvhtree as _bvhtree_stub      # the stub module the reference imported from
    _bvhtree_stub.BVHTree = (icp_oracle.NearestSurfaceBVH if sc.get("base_faces") is not None
                             else icp_oracle.NearestVertexBVH)
    icp_align_mod = ref_loader.load_reference()[1]
    icp_align_mod.BVHTree = _bvhtree_stub.BVHTree
    rec = dict(idx=[], K=[], d_stats=[], M=[], vl=None)
    bvh_box = {}

    def make_pairs_rec(align_obj, base_obj, base_bvh, vlist, thresh, sample=0, calc_stats=False):
        bvh_box["bvh"] = base_bvh
        base_bvh.log = []
        out = general.make_pairs(align_obj, base_obj, base_bvh, vlist, thresh, sample, calc_stats)
        rec["idx"].append(np.array(base_bvh.log, dtype=np.int32))
        rec["vl_full"] = np.array(vlist, dtype=np.int32)
        if out is not None:
            rec["K"].append(out[0].shape[1])
            rec["d_stats"].append(out[2] if out[2] is not None else [np.nan, np.nan])
            rec["A_last"], rec["B_last"] = out[0], out[1]
        return out

    def affine_rec(v0, v1, shear=True, scale=True, usesvd=True):
        assert shear is False and usesvd is True
        M = general.affine_matrix_from_points(v0, v1, shear=shear, scale=scale, usesvd=usesvd)
        rec["M"].append(M)
        return M

    saved = (icp_align.make_pairs, icp_align.affine_matrix_from_points,
             icp_align.get_addon_preferences)
    icp_align.make_pairs = make_pairs_rec
    icp_align.affine_matrix_from_points = affine_rec
    icp_align.get_addon_preferences = lambda: prefs
    try:
        ctx = ms.Context(align, [align, base])
        op = icp_align.OBJECT_OT_icp_align()
        assert op.poll(ctx)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ret = op.execute(ctx)
        assert ret == {"FINISHED"}
    finally:
        (icp_align.make_pairs, icp_align.affine_matrix_from_points,
         icp_align.get_addon_preferences) = saved
    text = buf.getvalue()
    converged = "Converged in" in text
    return align, rec, converged, prefs


def main():
    os.makedirs(OUT, exist_ok=True)
    for sc in scenarios():
        align, rec, converged, prefs = run_reference(sc)
        iters = len(rec["M"])
        out = dict(
            moving=np.asarray(sc["moving"], np.float32), base=np.asarray(sc["base"], np.float32),
            mw_align=np.asarray(sc.get("mw_align", np.identity(4)), np.float32),
            mw_base=np.asarray(sc.get("mw_base", np.identity(4)), np.float32),
            vlist=rec["vl_full"],
            iterations=np.int32(iters), converged=np.bool_(converged),
            K=np.array(rec["K"], np.int64), d_stats=np.array(rec["d_stats"], np.float64),
            M=np.array(rec["M"], np.float64), matrix_world=align.matrix_world.a.copy(),
            A_last=rec["A_last"], B_last=rec["B_last"],
        )
        # ragged per-iteration index logs (stride can shorten them) -> one flat array + offsets
        out["idx_flat"] = np.concatenate(rec["idx"]) if rec["idx"] else np.zeros(0, np.int32)
        out["idx_offsets"] = np.cumsum([0] + [len(x) for x in rec["idx"]]).astype(np.int64)
        for k, v in vars(prefs).items():
            out["pref_" + k] = np.array(v)
        for gname, (mask, w) in (sc.get("groups") or {}).items():
            out["group_%s_mask" % gname] = np.asarray(mask, bool)
            out["group_%s_weight" % gname] = np.asarray(w, np.float32)
        if sc.get("base_faces") is not None:
            out["base_faces"] = np.asarray(sc["base_faces"], np.int32)
        out["group_order"] = np.array(list((sc.get("groups") or {}).keys()), dtype="U32")
        path = os.path.join(OUT, sc["name"] + ".npz")
        np.savez_compressed(path, **out)
        print("%-18s iters=%2d converged=%-5s K[-1]=%d  %d bytes" % (
            sc["name"], iters, converged, rec["K"][-1], os.path.getsize(path)))

    # known-answer vectors from the reference docstrings (general.py:41-49,125-141) and
    # the error behaviour of the solver (general.py:149-157) / make_pairs (general.py:277)
    general, _ = ref_loader.load_reference()
    rng = np.random.default_rng(42)
    kat = {}
    v0 = (rng.random((3, 100)) - 0.5) * 20
    R = meshgen.rotation_about_axis(33.0, (0.3, -0.5, 0.8))
    for tag, s in (("rigid", 1.0), ("sim", 1.7)):
        v1 = s * (R @ v0) + np.array([[0.3], [-1.2], [2.5]])
        v1n = v1 + rng.normal(0, 1e-3, v1.shape)
        kat["%s_v0" % tag], kat["%s_v1" % tag] = v0, v1n
        kat["%s_M" % tag] = general.affine_matrix_from_points(v0, v1n, shear=False,
                                                              scale=(tag == "sim"), usesvd=True)
    # reflection case: det(U Vh) < 0 (general.py:184-187)
    v1 = np.diag([1.0, 1.0, -1.0]) @ v0 + rng.normal(0, 1e-2, v0.shape)
    kat["refl_v0"], kat["refl_v1"] = v0, v1
    kat["refl_M"] = general.affine_matrix_from_points(v0, v1, shear=False, scale=False, usesvd=True)
    kat["quat_q"] = np.array([0.99810947, 0.06146124, 0, 0])
    kat["quat_M"] = general.quaternion_matrix(kat["quat_q"])
    np.savez_compressed(os.path.join(OUT, "solver_kat.npz"), **kat)
    print("solver_kat written")


if __name__ == "__main__":
    main()
