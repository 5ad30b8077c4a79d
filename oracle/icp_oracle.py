"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference ICP hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; it is the checker, never the thing
shipped or measured as the product.

Parity status: **pinned against the reference run here** -- ``oracle/gen_golden.py``
imports the UNMODIFIED ``/root/reference/functions/general.py`` and
``/root/reference/operators/icp_align.py`` under stub ``bpy``/``mathutils`` modules
(``oracle/ref_loader.py``) and stores their outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those files.
The reference has no tests or golden vectors of its own for this path,
and Blender's C ``mathutils`` cannot be run here, so the fp32 rounding rules S1..S5 in
``oracle/mathutils_shim.py`` remain documented assumptions.

Functions and the reference lines they follow
  build_vlist_port            /root/reference/operators/icp_align.py:56-80
  stride_factor_port          /root/reference/operators/icp_align.py:89
  make_pairs_port             /root/reference/functions/general.py:257-329  (per-vertex loop, faithful)
  make_pairs_vec              same arithmetic, numpy-vectorised (used at 100k..10M)
  affine_matrix_from_points_port   /root/reference/functions/general.py:146-167,179-190,208-217
  icp_execute_port            /root/reference/operators/icp_align.py:82-151
  percentile_keep_port        NEW semantics, defined by this repo
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

from . import mathutils_shim as ms

f32 = np.float32
f64 = np.float64

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    """Load (building on first use) the C nearest-vertex oracle."""
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle_nn.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "liboracle_nn.so"])
        lib = ctypes.CDLL(so)
        lib.oracle_nn_bruteforce.restype = ctypes.c_int
        lib.oracle_nn_bruteforce.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
        ]
        lib.oracle_d2.restype = ctypes.c_float
        lib.oracle_d2.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_d2_candidates.restype = None
        lib.oracle_d2_candidates.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
        ]
        lib.oracle_num_threads.restype = ctypes.c_int
        lib.oracle_nn_surface_bruteforce.restype = ctypes.c_int
        lib.oracle_nn_surface_bruteforce.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
        ]
        _LIB = lib
    return _LIB


def num_threads() -> int:
    return int(_lib().oracle_num_threads())


# ------------------------------------------------------------------ nearest vertex (spec S4)
def nn_bruteforce(q: np.ndarray, base: np.ndarray, index_offset: int = 0, threads: int = 0):
    """Exact nearest base VERTEX for each query: (idx int32[M], d2 float32[M]).

    Metric S4 (fp32, two fmaf), ties -> lowest index.  Restates the query at
    general.py:297 in the nearest-vertex variant.
    """
    q = np.ascontiguousarray(q, dtype=f32).reshape(-1, 3)
    base = np.ascontiguousarray(base, dtype=f32).reshape(-1, 3)
    m = q.shape[0]
    idx = np.empty(m, dtype=np.int32)
    d2 = np.empty(m, dtype=f32)
    rc = _lib().oracle_nn_bruteforce(
        q.ctypes.data, m, base.ctypes.data, base.shape[0], int(index_offset),
        idx.ctypes.data, d2.ctypes.data, int(threads))
    if rc != 0:
        raise MemoryError("oracle_nn_bruteforce")
    return idx, d2


def nn_surface_bruteforce(q: np.ndarray, tri_xyz: np.ndarray, index_offset: int = 0, threads: int = 0):
    """Closest point on the triangulated surface for each query (rules S9 / S10; restates the
    reference's real query, icp_align.py:53 + general.py:297):
    (triangle idx int32[M], d2 float32[M], closest point float32[M,3]); ties -> lowest triangle."""
    q = np.ascontiguousarray(q, dtype=f32).reshape(-1, 3)
    tri = np.ascontiguousarray(tri_xyz, dtype=f32).reshape(-1, 9)
    m = q.shape[0]
    idx = np.empty(m, dtype=np.int32)
    d2 = np.empty(m, dtype=f32)
    co = np.empty((m, 3), dtype=f32)
    _lib().oracle_nn_surface_bruteforce(q.ctypes.data, m, tri.ctypes.data, tri.shape[0], int(index_offset),
                                        idx.ctypes.data, d2.ctypes.data, co.ctypes.data, int(threads))
    return idx, d2, co


class NearestSurfaceBVH:
    """Drop-in for ``BVHTree.FromObject(base_obj, depsgraph)`` with the reference's REAL semantics:
    ``find_nearest(co)`` -> (closest point on the triangulated surface, None, triangle index, dist).
    Brute force over the triangles (exact under S9/S10); for golden generation at small sizes."""

    def __init__(self, coords, tris):
        coords = np.ascontiguousarray(coords, dtype=f32).reshape(-1, 3)
        self.tri_xyz = np.ascontiguousarray(coords[np.asarray(tris)].reshape(-1, 9))
        self.log = []

    @classmethod
    def FromObject(cls, obj, depsgraph=None):
        n = len(obj.data.vertices)
        buf = np.empty(n * 3, dtype=f32)
        obj.data.vertices.foreach_get("co", buf)
        lt = obj.data.loop_triangles
        tb = np.empty(len(lt) * 3, dtype=np.int32)
        lt.foreach_get("vertices", tb)
        return cls(buf.reshape(-1, 3), tb.reshape(-1, 3))

    def find_nearest(self, co):
        if self.tri_xyz.shape[0] == 0:
            return (None, None, None, None)
        q = np.array([co[0], co[1], co[2]], dtype=f32).reshape(1, 3)
        idx, d2, pt = nn_surface_bruteforce(q, self.tri_xyz, threads=1)
        self.log.append(int(idx[0]))
        return (ms.Vector(pt[0]), None, int(idx[0]), math.sqrt(float(d2[0])))


def pack_key(d2: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """(float_bits(d2) << 32) | index -- the uint64 key whose MIN resolves shards."""
    bits = np.ascontiguousarray(d2, dtype=f32).view(np.uint32).astype(np.uint64)
    return (bits << np.uint64(32)) | np.ascontiguousarray(idx).astype(np.uint32).astype(np.uint64)


def unpack_key(key: np.ndarray):
    key = np.ascontiguousarray(key, dtype=np.uint64)
    idx = (key & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32)
    d2 = (key >> np.uint64(32)).astype(np.uint32).view(f32)
    return idx, d2


def nn_kdtree_exact(q: np.ndarray, base: np.ndarray, tree=None, k: int = 8, workers: int = -1):
    """Same answer as nn_bruteforce, found through scipy's cKDTree + an exact S4 re-check.

    Candidates are the k nearest in fp64; the S4 metric is evaluated on them and the
    lowest-index minimum wins.  A query is handed to the brute force whenever the k-th
    candidate is not clearly (1e-4 relative in d2) farther than the best, so that no
    rounding of S4 can promote a point outside the candidate set.
    """
    from scipy.spatial import cKDTree

    q = np.ascontiguousarray(q, dtype=f32).reshape(-1, 3)
    base = np.ascontiguousarray(base, dtype=f32).reshape(-1, 3)
    n = base.shape[0]
    k = min(k, n)
    if tree is None:
        tree = cKDTree(base.astype(f64))
    dd, ii = tree.query(q.astype(f64), k=k, workers=workers)
    if k == 1:
        dd, ii = dd[:, None], ii[:, None]
    cand = base[ii]  # (M,k,3) fp32
    d2 = _d2_spec_numpy_exact(q[:, None, :], cand)
    dmin = d2.min(axis=1)
    big = np.where(d2 == dmin[:, None], ii, np.iinfo(np.int64).max)
    idx = big.min(axis=1).astype(np.int32)
    if k < n:
        unsafe = ~(dd[:, -1] ** 2 > (dd[:, 0] ** 2) * (1 + 1e-4) + 1e-30)
        if unsafe.any():
            bi, bd = nn_bruteforce(q[unsafe], base)
            idx[unsafe] = bi
            dmin[unsafe] = bd
    return idx, dmin.astype(f32)


def _fma_f32_exact(a, b, c):
    """Correctly-rounded fp32 fma(a, b, c) in numpy via an error-free transformation.

    a*b is exact in fp64 (24+24 bits).  s = fl64(p + c) with error e (TwoSum); rounding s
    to fp32 is only wrong when s sits exactly on an fp32 tie and e != 0 -- fix by nudging.
    """
    p = a.astype(f64) * b.astype(f64)
    c = np.broadcast_to(c.astype(f64), p.shape)
    s = p + c
    bb = s - p
    e = (p - (s - bb)) + (c - bb)
    r = s.astype(f32)
    # detect the double-rounding hazard: s exactly halfway between two floats and e != 0
    lo = np.nextafter(r, f32(-np.inf)).astype(f64)
    hi = np.nextafter(r, f32(np.inf)).astype(f64)
    rr = r.astype(f64)
    tie_lo = (s == (lo + rr) / 2) & (e != 0)
    tie_hi = (s == (hi + rr) / 2) & (e != 0)
    r = np.where(tie_lo & (e < 0), lo, r.astype(f64))
    r = np.where(tie_hi & (e > 0), hi, r)
    # ties that were rounded the right way stay; convert back
    return r.astype(f32)


def _d2_spec_numpy_exact(q, b):
    """S4 on numpy arrays (broadcasting), with correctly-rounded fp32 fused steps."""
    q = np.asarray(q, dtype=f32)
    b = np.asarray(b, dtype=f32)
    dx = (q[..., 0] - b[..., 0]).astype(f32)
    dy = (q[..., 1] - b[..., 1]).astype(f32)
    dz = (q[..., 2] - b[..., 2]).astype(f32)
    t = (dx * dx).astype(f32)
    t = _fma_f32_exact(dy, dy, t)
    return _fma_f32_exact(dz, dz, t)


class NearestVertexBVH:
    """Drop-in for ``BVHTree.FromObject(base_obj, depsgraph)`` (icp_align.py:53) that answers
    ``find_nearest(co)`` (general.py:297) with the nearest base VERTEX.

    Returns ``(co1 Vector, None, index, distance)`` like mathutils' BVHTree.  Every query's
    winning index is appended to ``self.log`` so golden files can record correspondences.
    """

    def __init__(self, base_coords, mode="auto"):
        self.base = np.ascontiguousarray(base_coords, dtype=f32).reshape(-1, 3)
        self.log = []
        self.mode = mode
        self._tree = None
        if mode in ("kdtree", "kdtree_fast") or (mode == "auto" and self.base.shape[0] > 4096):
            from scipy.spatial import cKDTree

            self._tree = cKDTree(self.base.astype(f64))

    @classmethod
    def FromObject(cls, obj, depsgraph=None, mode="auto"):
        n = len(obj.data.vertices)
        buf = np.empty(n * 3, dtype=f32)
        obj.data.vertices.foreach_get("co", buf)
        return cls(buf.reshape(-1, 3), mode=mode)

    def find_nearest(self, co):
        if self.mode == "kdtree_fast":
            # timing-only provider (bench.py cpu_baseline): one fp64 k=1 lookup, the cheapest
            # stand-in for Blender's C BVH; no S4 re-check, so near-ties may differ from S4
            d, j = self._tree.query(co.v)
            return (ms.Vector(self.base[j]), None, j, d)
        q = np.array([co[0], co[1], co[2]], dtype=f32).reshape(1, 3)
        if self.base.shape[0] == 0:
            return (None, None, None, None)
        if self._tree is None:
            idx, d2 = nn_bruteforce(q, self.base, threads=1)
        else:
            idx, d2 = nn_kdtree_exact(q, self.base, tree=self._tree, workers=1)
        j = int(idx[0])
        self.log.append(j)
        return (ms.Vector(self.base[j]), None, j, math.sqrt(float(d2[0])))


# ------------------------------------------------------------------ mask / parameters
def build_vlist_port(n_verts: int, groups: dict | None) -> np.ndarray:
    """Vertex-group mask rule of icp_align.py:56-80 -> ascending int32 vertex indices.

    ``groups``: name -> (member_mask bool[M], weight float32[M]).  'icp_include' wins over
    'icp_exclude' (the ``elif`` at :67); include keeps weight > 0.9 (:64); exclude keeps
    non-members (:71) and members with weight < 0.1 (:75).
    """
    groups = groups or {}
    if "icp_include" in groups:
        mask, w = groups["icp_include"]
        keep = np.asarray(mask, bool) & (np.asarray(w, f32).astype(f64) > 0.9)
    elif "icp_exclude" in groups:
        mask, w = groups["icp_exclude"]
        mask = np.asarray(mask, bool)
        keep = (~mask) | (mask & (np.asarray(w, f32).astype(f64) < 0.1))
    else:
        keep = np.ones(n_verts, dtype=bool)
    return np.nonzero(keep)[0].astype(np.int32)


def stride_factor_port(sample_fraction: float) -> int:
    """icp_align.py:89 -- Python's round() (banker's rounding); ZeroDivisionError at 0."""
    return round(1 / sample_fraction)


# ------------------------------------------------------------------ vectorised arithmetic (S1, S5)
def matvec43_vec(m: np.ndarray, v: np.ndarray) -> np.ndarray:
    """S1 for an (M,3) float32 array: float products, double accumulation, float store."""
    m = np.asarray(m, dtype=f32)
    v = np.ascontiguousarray(v, dtype=f32)
    out = np.empty_like(v)
    for r in range(3):
        acc = (m[r, 0] * v[:, 0]).astype(f32).astype(f64)
        acc = acc + (m[r, 1] * v[:, 1]).astype(f32).astype(f64)
        acc = acc + (m[r, 2] * v[:, 2]).astype(f32).astype(f64)
        acc = acc + f64(f32(m[r, 3] * f32(1.0)))
        out[:, r] = acc.astype(f32)
    return out


def length3_vec(d: np.ndarray) -> np.ndarray:
    """S5 for an (M,3) float32 array -> float64 lengths."""
    d = np.asarray(d, dtype=f32)
    acc = (d[:, 2] * d[:, 2]).astype(f32).astype(f64)
    acc = acc + (d[:, 1] * d[:, 1]).astype(f32).astype(f64)
    acc = acc + (d[:, 0] * d[:, 0]).astype(f32).astype(f64)
    return np.sqrt(acc)


def percentile_keep_port(dist: np.ndarray, reject_fraction: float) -> np.ndarray:
    """NEW semantics: keep the K = L - floor(reject_fraction * L) queries with the
    smallest (d, query_position) lexicographic key, where d is the fp32 base-local squared
    distance S4 carried in the NN key (it is the one quantity every rank holds for every query
    once the shards are resolved).  Returns a bool mask over positions."""
    n = dist.shape[0]
    k = n - int(math.floor(float(reject_fraction) * n))
    order = np.lexsort((np.arange(n), dist))
    keep = np.zeros(n, dtype=bool)
    keep[order[:k]] = True
    return keep


def make_pairs_vec(coords, vlist, mw_align, mw_base, base, thresh, sample=0, calc_stats=False,
                   reject_fraction=0.0, nn="brute", index_offset=0, return_all=False, tri_xyz=None):
    """numpy-vectorised restatement of make_pairs (general.py:257-329), nearest-vertex variant.

    coords: (M,3) fp32 align-local vertices; vlist: int indices (already masked);
    mw_align/mw_base: (4,4) fp32 world matrices; base: (N,3) fp32 base-local vertices.
    Returns (A, B, d_stats) like the reference -- A, B are 3xK float64 -- or, with
    return_all, a dict that also carries idx / d2 / dist / keep per query position.
    """
    if not thresh > 0:          # general.py:277 -- the reference falls through and returns None
        return None
    coords = np.ascontiguousarray(coords, dtype=f32).reshape(-1, 3)
    base = np.ascontiguousarray(base, dtype=f32).reshape(-1, 3)
    mx1 = np.asarray(mw_align, dtype=f32).reshape(4, 4)
    mx2 = np.asarray(mw_base, dtype=f32).reshape(4, 4)
    imx1 = ms.invert44(mx1)     # general.py:265
    imx2 = ms.invert44(mx2)     # general.py:266
    vl = np.asarray(vlist, dtype=np.int64)
    if sample > 1:              # general.py:274-275
        vl = vl[0::sample]
    a = coords[vl]
    q = matvec43_vec(imx2, matvec43_vec(mx1, a))                      # general.py:287
    if tri_xyz is not None:                                           # surface variant (N1)
        idx, d2, co1 = nn_surface_bruteforce(q, tri_xyz, index_offset=index_offset)
    elif nn == "brute":
        idx, d2 = nn_bruteforce(q, base, index_offset=index_offset)   # general.py:297 (D2)
        co1 = base[idx - index_offset]
    else:
        idx, d2 = nn_kdtree_exact(q, base)
        idx = idx + np.int32(index_offset)
        co1 = base[idx - index_offset]
    wq = matvec43_vec(mx2, q)
    w1 = matvec43_vec(mx2, co1)
    dist = length3_vec((wq - w1).astype(f32))                         # general.py:299
    keep = dist < float(thresh)                                       # general.py:302 (strict)
    if reject_fraction and reject_fraction > 0:
        keep &= percentile_keep_port(d2, reject_fraction)
    b = matvec43_vec(imx1, w1)                                        # general.py:304
    A = np.ascontiguousarray(a[keep].astype(f64).T)                   # general.py:313-321
    B = np.ascontiguousarray(b[keep].astype(f64).T)
    if calc_stats:                                                    # general.py:323-328
        kd = dist[keep]
        d_stats = [float(np.mean(kd)), float(np.std(kd))] if kd.size else [float("nan")] * 2
    else:
        d_stats = None
    if return_all:
        return dict(A=A, B=B, d_stats=d_stats, idx=idx, d2=d2, dist=dist, keep=keep, q=q,
                    b=b, vl=vl)
    return A, B, d_stats


# ------------------------------------------------------------------ faithful per-vertex port
def make_pairs_port(align_obj, base_obj, base_bvh, vlist, thresh, sample=0, calc_stats=False):
    """Per-vertex Python loop with the same cost structure as general.py:257-329: one vertex
    lookup, five matrix-vector products, one find_nearest and one length per vertex, then an
    item-wise fill of the 3xK arrays.  This is what ``cpu_baseline`` times (kind "port")."""
    world_a = align_obj.matrix_world
    world_b = base_obj.matrix_world
    inv_a = world_a.inverted()
    inv_b = world_b.inverted()
    src, dst, dists = [], [], []
    if sample > 1:
        vlist = vlist[0::sample]
    if not thresh > 0:
        return None
    verts = align_obj.data.vertices
    for vi in vlist:
        p = verts[vi].co
        probe = inv_b @ (world_a @ p)
        hit, _normal, face, _d = base_bvh.find_nearest(probe)
        gap = (world_b @ probe - world_b @ hit).length
        if face != -1 and gap < thresh:
            src.append(p)
            dst.append(inv_a @ (world_b @ hit))
            if calc_stats:
                dists.append(gap)
    k = len(src)
    A = np.zeros(shape=[3, k])
    B = np.zeros(shape=[3, k])
    for i in range(k):
        s, d = src[i], dst[i]
        A[0][i], A[1][i], A[2][i] = s[0], s[1], s[2]
        B[0][i], B[1][i], B[2][i] = d[0], d[1], d[2]
    d_stats = [np.mean(dists), np.std(dists)] if calc_stats else None
    return A, B, d_stats


# ------------------------------------------------------------------ rigid / similarity solver
def affine_matrix_from_points_port(v0, v1, scale=False):
    """The only branch ICP reaches (shear=False, usesvd=True; general.py:146-167,179-190,208-217).

    fp64 copies; ValueError when K < ndims or shapes differ (:149-157); centroids (:160-167);
    SVD of the 3x3 cross-covariance (:181); R = U Vh with the reflection fix on the last
    singular direction (:183-187); optional uniform scale sqrt(sum v1^2 / sum v0^2) on the
    centred data (:208-212); centroids moved back (:215-217).
    """
    v0 = np.array(v0, dtype=f64, copy=True)
    v1 = np.array(v1, dtype=f64, copy=True)
    nd = v0.shape[0]
    if nd < 2 or v0.shape[1] < nd or v0.shape != v1.shape:
        raise ValueError("input arrays are of wrong shape or type")
    c0 = v0.mean(axis=1)
    c1 = v1.mean(axis=1)
    x0 = v0 - c0[:, None]
    x1 = v1 - c1[:, None]
    u, s, vh = np.linalg.svd(x1 @ x0.T)
    rot = u @ vh
    if np.linalg.det(rot) < 0.0:
        rot = rot - np.outer(u[:, nd - 1], vh[nd - 1, :] * 2.0)
    if scale:
        rot = rot * math.sqrt((x1 * x1).sum() / (x0 * x0).sum())
    out = np.identity(nd + 1)
    out[:nd, :nd] = rot
    out[:nd, nd] = c1 - rot @ c0
    return out


def rotation_angle(m44: np.ndarray) -> float:
    """|angle| of the rotation part, for the report at icp_align.py:139,149 (reporting only)."""
    r = np.asarray(m44, dtype=f64)[:3, :3]
    n = np.linalg.norm(r, axis=0)
    n[n == 0] = 1.0
    r = r / n
    c = (r[0, 0] + r[1, 1] + r[2, 2] - 1.0) / 2.0
    return math.acos(max(-1.0, min(1.0, c)))


def icp_execute_port(coords, vlist, mw_align, mw_base, base, *, thresh=0.5, sample_fraction=0.5,
                     iters=50, target_d=0.01, use_target=True, align_meth="0",
                     reject_fraction=0.0, nn="brute", record=False, tri_xyz=None):
    """The iteration loop of OBJECT_OT_icp_align.execute (icp_align.py:82-151), vectorised.

    Per iteration: make_pairs (:101) -> solver (:106-109) -> fp32 copy of M (:116-119) ->
    matrix_world = matrix_world @ new_mat in fp32 (:121, rule S3) -> translation-ring
    convergence test (:136-142; ring starts at 2*target_d (:93), slot n mod 5, converged when
    all five are < target_d; only evaluated when use_target made d_stats truthy).
    Returns dict(matrix_world fp32 (4,4), iterations, converged, per-iteration records).
    """
    mw = np.array(mw_align, dtype=f32).reshape(4, 4).copy()
    mw_base = np.asarray(mw_base, dtype=f32).reshape(4, 4)
    factor = stride_factor_port(sample_fraction)
    ring = [target_d * 2] * 5
    n, converged = 0, False
    recs = []
    while n < iters and not converged:
        res = make_pairs_vec(coords, vlist, mw, mw_base, base, thresh, factor,
                             calc_stats=use_target, reject_fraction=reject_fraction, nn=nn,
                             return_all=True, tri_xyz=tri_xyz)
        if res is None:
            raise TypeError("cannot unpack non-iterable NoneType object")  # icp_align.py:101
        M = affine_matrix_from_points_port(res["A"], res["B"], scale=(align_meth == "1"))
        new_mat = M.astype(f32)
        mw = ms.matmul44(mw, new_mat)
        tlen = ms.length3(new_mat[:3, 3])
        ang = rotation_angle(new_mat)
        if res["d_stats"]:
            ring[n % 5] = tlen
            if all(d < target_d for d in ring):
                converged = True
        rec = dict(K=int(res["keep"].sum()), M=M, trans=tlen, angle=ang, d_stats=res["d_stats"])
        if record:
            rec.update(idx=res["idx"], keep=res["keep"], d2=res["d2"], dist=res["dist"])
        recs.append(rec)
        n += 1
    return dict(matrix_world=mw, iterations=n, converged=converged, records=recs)
