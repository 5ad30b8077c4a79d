"""Reference-compatible function surface for the ICP path, backed by libicp_b200.so.

Same names, argument meaning and error behaviour as the reference's
``functions/general.py`` for the functions on the hot path:

    make_pairs(align_obj, base_obj, base_bvh, vlist, thresh, sample=0, calc_stats=False)
        -> (A, B, d_stats)                       reference general.py:257-329
    affine_matrix_from_points(v0, v1, shear=True, scale=True, usesvd=True)
        -> ndarray(4, 4)                         reference general.py:105-217

plus the fused entry points BASELINE.json's north_star names (they do not exist in the
reference -- SURVEY.md D1 -- and are additions):

    calc_target(align_obj, base_obj, base_bvh, vlist, sample=0) -> (idx, dist)
    icp_step(align_obj, base_obj, base_bvh, vlist, thresh, sample=0, ...) -> (M, stats)

``base_bvh`` is a :class:`VertexNN`, the drop-in for
``BVHTree.FromObject(base_obj, depsgraph)`` (reference operators/icp_align.py:53): it uploads the
base vertices to the GPU once and answers nearest-VERTEX queries (the
reference's BVH answers closest point on the triangulated surface).

Nothing here computes on the CPU; without the CUDA library every call raises.
"""
from __future__ import annotations

import numpy as np

from . import _native as N
from .engine import IcpEngine, IcpError

__all__ = ["VertexNN", "SurfaceNN", "base_index_from_object", "extract_triangles", "make_pairs", "affine_matrix_from_points", "calc_target", "icp_step",
           "extract_coords", "matrix_to_array", "array_to_matrix_like"]


# ----------------------------------------------------------------------------- marshalling
def extract_coords(obj, out=None) -> np.ndarray:
    """(n, 3) float32 vertex coordinates of ``obj.data.vertices`` (bulk foreach_get if present).
    ``out``: a flat float32 buffer of 3 n elements to fill instead of a fresh array (pinned staging)."""
    verts = obj.data.vertices
    n = len(verts)
    buf = np.empty(n * 3, dtype=np.float32) if out is None else out
    if hasattr(verts, "foreach_get"):
        verts.foreach_get("co", buf)
    else:
        for i in range(n):
            co = verts[i].co
            buf[3 * i], buf[3 * i + 1], buf[3 * i + 2] = co[0], co[1], co[2]
    return buf.reshape(n, 3)


def extract_triangles(obj) -> np.ndarray:
    """(F, 3) int32 vertex indices of the mesh's loop triangles (what BVHTree.FromObject walks);
    empty for a face-less mesh (a point cloud)."""
    mesh = obj.data
    lt = getattr(mesh, "loop_triangles", None)
    if lt is None:
        return np.zeros((0, 3), dtype=np.int32)
    if hasattr(mesh, "calc_loop_triangles"):
        mesh.calc_loop_triangles()
    n = len(lt)
    buf = np.empty(n * 3, dtype=np.int32)
    if n:
        lt.foreach_get("vertices", buf)
    return buf.reshape(n, 3)


def matrix_to_array(m) -> np.ndarray:
    """mathutils.Matrix (or anything indexable [r][c]) -> (4, 4) float32, row-major."""
    if isinstance(m, np.ndarray):
        return np.ascontiguousarray(m, dtype=np.float32).reshape(4, 4)
    return np.array([[m[r][c] for c in range(4)] for r in range(4)], dtype=np.float32)


def array_to_matrix_like(a: np.ndarray, like):
    """Build a matrix of the host's own type (mathutils.Matrix in Blender) from a (4,4) array."""
    if isinstance(like, np.ndarray):
        return np.array(a, dtype=like.dtype)
    rows = [[float(a[r][c]) for c in range(4)] for r in range(4)]
    return type(like)(rows)


# Search used when the caller does not choose one: the exact BVH search (row N2) returns the same
# keys as the brute-force kernel and is ~400x faster at 1M x 1M; set to N.NN_BRUTEFORCE to force
# the O(M N) north-star kernel.
DEFAULT_NN_MODE = N.NN_BVH


class VertexNN:
    """GPU nearest-vertex index over a base object; stands where the reference passes
    ``base_bvh`` (operators/icp_align.py:53,101; functions/general.py:297)."""

    def __init__(self, base_coords, **engine_kwargs):
        engine_kwargs.setdefault("nn_mode", DEFAULT_NN_MODE)
        if engine_kwargs.get("world_size", 1) > 1:
            # the O(log N) searches shard the QUERIES (base replicated); only brute force shards the base
            engine_kwargs.setdefault("shard_mode", N.SHARD_BASE if engine_kwargs["nn_mode"] == N.NN_BRUTEFORCE
                                     else N.SHARD_QUERIES)
        self.engine = IcpEngine(**engine_kwargs)
        base_coords = np.ascontiguousarray(base_coords, dtype=np.float32).reshape(-1, 3)
        self.n_base = base_coords.shape[0]
        self._upload_base(base_coords)

    def _upload_base(self, base_coords):
        if self._base_is_sharded():
            self.engine.set_base_sharded(base_coords)
        else:
            self.engine.set_base(base_coords)

    def _base_is_sharded(self):
        p = self.engine.params
        return p.world_size > 1 and p.shard_mode == N.SHARD_BASE

    def connect(self, dist=None):
        """Collective (world_size > 1): build the NCCL communicator of this index's engine."""
        from .dist import connect_engine
        connect_engine(self.engine, dist)

    @classmethod
    def FromObject(cls, obj, depsgraph=None, **engine_kwargs):
        return cls(extract_coords(obj), **engine_kwargs)

    def invalidate(self):
        """Kept for callers of the first version; the coordinates are re-read on every call now."""

    # internal: the reference reads vert.co on every call (general.py:284), so do we -- the C side keeps
    # its Morton slot order and the previous winners when the vertex count and vlist are unchanged
    def _bind_moving(self, align_obj, vlist):
        vl = np.ascontiguousarray(vlist, dtype=np.int32)
        coords = extract_coords(align_obj)
        p = self.engine.params
        if p.world_size > 1 and p.shard_mode == N.SHARD_QUERIES:
            self.engine.update_params(stride=self._stride)
            self.engine.set_moving_sharded(coords, vl)
        else:
            self.engine.set_moving(coords, vl)

    def _bind(self, align_obj, base_obj, vlist, thresh=None, sample=0, **params):
        self._stride = int(sample) if sample and sample > 1 else 1
        upd = dict(stride=self._stride)
        if thresh is not None:
            upd["thresh"] = float(thresh)
        upd.update(params)
        self.engine.update_params(**upd)
        self._bind_moving(align_obj, vlist)      # query-sharded engines fold the stride into their slice here
        self.engine.set_matrices(matrix_to_array(align_obj.matrix_world),
                                 matrix_to_array(base_obj.matrix_world))

    def close(self):
        self.engine.close()


class SurfaceNN(VertexNN):
    """GPU closest-point-on-surface index over the base object's TRIANGLES: the reference's real
    ``BVHTree.FromObject(base_obj, depsgraph)`` semantics (row N1).  ``find_nearest`` answers
    come back as (triangle index, closest point) inside make_pairs / icp_step."""

    def __init__(self, base_coords, triangles, **engine_kwargs):
        engine_kwargs["nn_mode"] = N.NN_SURFACE
        if engine_kwargs.get("world_size", 1) > 1:
            engine_kwargs.setdefault("shard_mode", N.SHARD_QUERIES)
        self.engine = IcpEngine(**engine_kwargs)
        coords = np.ascontiguousarray(base_coords, dtype=np.float32).reshape(-1, 3)
        tris = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3)
        self.n_base = tris.shape[0]
        corners = np.ascontiguousarray(coords[tris].reshape(-1, 9))
        if self._base_is_sharded():
            self.engine.set_base_triangles_sharded(corners)
        else:
            self.engine.set_base_triangles(corners)

    @classmethod
    def FromObject(cls, obj, depsgraph=None, **engine_kwargs):
        return cls(extract_coords(obj), extract_triangles(obj), **engine_kwargs)


# "auto": triangles when the base mesh has faces (the reference's behaviour), vertices for a
# face-less point cloud (which the reference cannot align at all); or force one.
DEFAULT_CORRESPONDENCE = "auto"


def base_index_from_object(base_obj, depsgraph=None, correspondence=None, **engine_kwargs):
    """The drop-in for ``BVHTree.FromObject(base_obj, depsgraph)`` (operators/icp_align.py:53)."""
    mode = correspondence or DEFAULT_CORRESPONDENCE
    if mode == "surface" or (mode == "auto" and len(extract_triangles(base_obj)) > 0):
        return SurfaceNN.FromObject(base_obj, depsgraph, **engine_kwargs)
    return VertexNN.FromObject(base_obj, depsgraph, **engine_kwargs)


def _require_vertex_nn(base_bvh):
    if not isinstance(base_bvh, VertexNN):
        raise TypeError(
            "base_bvh must be object_alignment_b200.VertexNN.FromObject(base_obj, depsgraph); "
            "got %r (the CPU BVHTree path is not part of this engine)" % type(base_bvh).__name__)


# ----------------------------------------------------------------------------- reference API
def make_pairs(align_obj, base_obj, base_bvh, vlist, thresh, sample=0, calc_stats=False):
    """Drop-in for reference make_pairs (functions/general.py:257-329).

    Returns ``(A, B, d_stats)``: A = kept align-local vertices, B = their nearest base vertices
    brought into align-local space (both 3xK float64), d_stats = [mean, std] of the kept
    world-space distances or None.  Like the reference, returns None when ``thresh <= 0``
    (general.py:277 -- the caller's tuple-unpack then raises TypeError).
    """
    _require_vertex_nn(base_bvh)
    if not thresh > 0:
        return None
    base_bvh._bind(align_obj, base_obj, vlist, thresh, sample)
    return base_bvh.engine.make_pairs(calc_stats=bool(calc_stats))


_default_engine = None


def _solver_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = IcpEngine()
    return _default_engine


def affine_matrix_from_points(v0, v1, shear=True, scale=True, usesvd=True):
    """Drop-in for reference affine_matrix_from_points (functions/general.py:105-217) on the
    branch every ICP caller uses: ``shear=False, usesvd=True`` with 3-D points
    (operators/icp_align.py:107,109).  The affine (shear=True) and quaternion (usesvd=False)
    branches are never reached by the path this engine replaces and raise NotImplementedError.
    """
    v0 = np.array(v0, dtype=np.float64, copy=True)
    v1 = np.array(v1, dtype=np.float64, copy=True)
    if v0.ndim != 2 or v1.ndim != 2:
        raise ValueError("input arrays are of wrong shape or type")
    ndims = v0.shape[0]
    if ndims < 2 or v0.shape[1] < ndims or v0.shape != v1.shape:      # general.py:149-157
        raise ValueError("input arrays are of wrong shape or type")
    if shear or not usesvd or ndims != 3:
        raise NotImplementedError(
            "only shear=False, usesvd=True, 3-D is on the ICP path (operators/icp_align.py:107,109)")
    try:
        return _solver_engine().solve_pairs(v0, v1, scale=bool(scale))
    except IcpError as e:
        if e.code == N.ICP_ERR_TOO_FEW_PAIRS:
            raise ValueError("input arrays are of wrong shape or type") from None
        raise


# ----------------------------------------------------------------------------- fused additions
def calc_target(align_obj, base_obj, base_bvh, vlist, sample=0):
    """NN half of make_pairs (general.py:274-297): for every sampled vertex the index of the
    nearest base vertex and the base-local distance.  -> (idx int32[L], dist float32[L])."""
    _require_vertex_nn(base_bvh)
    base_bvh._bind(align_obj, base_obj, vlist, None, sample)
    idx, d2 = base_bvh.engine.nn()
    return idx, np.sqrt(d2)


def icp_step(align_obj, base_obj, base_bvh, vlist, thresh, sample=0, scale=False, calc_stats=True,
             target_d=0.01, reject_fraction=0.0):
    """One loop body of execute (operators/icp_align.py:98-149) fused on the GPU:
    pairs -> M -> ``align_obj.matrix_world = align_obj.matrix_world @ fp32(M)``.
    Returns (M float64 (4,4), IterStats).  Raises ValueError like the reference's solver when
    fewer than 3 pairs survive, TypeError when thresh <= 0 (icp_align.py:101 unpacking None)."""
    _require_vertex_nn(base_bvh)
    if not thresh > 0:
        raise TypeError("cannot unpack non-iterable NoneType object")
    base_bvh._bind(align_obj, base_obj, vlist, thresh, sample, align_meth=1 if scale else 0,
                   use_target=1 if calc_stats else 0, target_d=float(target_d),
                   reject_fraction=float(reject_fraction))
    try:
        M, stats = base_bvh.engine.step()
    except IcpError as e:
        if e.code == N.ICP_ERR_TOO_FEW_PAIRS:
            raise ValueError("input arrays are of wrong shape or type") from None
        raise
    align_obj.matrix_world = array_to_matrix_like(base_bvh.engine.matrix_world(),
                                                  align_obj.matrix_world)
    return M, stats
