"""IcpEngine -- thin owner of one `icp_handle` (one B200, one stream).

All arithmetic happens in libicp_b200.so; this file only marshals numpy buffers.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _native as N

_STATUS_NAMES = {
    N.ICP_ERR_BAD_ARG: "ICP_ERR_BAD_ARG", N.ICP_ERR_TOO_FEW_PAIRS: "ICP_ERR_TOO_FEW_PAIRS",
    N.ICP_ERR_CUDA: "ICP_ERR_CUDA", N.ICP_ERR_NCCL: "ICP_ERR_NCCL", N.ICP_ERR_OOM: "ICP_ERR_OOM",
    N.ICP_ERR_STATE: "ICP_ERR_STATE", N.ICP_ERR_SINGULAR: "ICP_ERR_SINGULAR",
}


class IcpError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s: %s" % (_STATUS_NAMES.get(code, code), message))
        self.code = code
        self.message = message


def stride_from_sample_fraction(sample_fraction: float) -> int:
    """factor = round(1/sample)  (reference operators/icp_align.py:89; Python banker's rounding,
    ZeroDivisionError when sample_fraction == 0, exactly like the reference)."""
    return round(1 / sample_fraction)


def shard_range(n_global: int, rank: int, world_size: int):
    """Contiguous equal split of the base index range [lo, hi) for `rank`."""
    lo = (n_global * rank) // world_size
    hi = (n_global * (rank + 1)) // world_size
    return lo, hi


def _f32c(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


class IterStats:
    __slots__ = ("pairs", "mean_dist", "std_dist", "trans", "angle", "M")

    def __init__(self, s: N.IcpIterStats):
        self.pairs = int(s.pairs)
        self.mean_dist = float(s.mean_dist)
        self.std_dist = float(s.std_dist)
        self.trans = float(s.trans)
        self.angle = float(s.angle)
        self.M = np.array(s.M, dtype=np.float64).reshape(4, 4)


class IcpEngine:
    """One GPU's ICP state: base shard, moving vertices, pose, convergence ring.

    Parameters mirror ObjectAlignmentPreferences (reference lib/preferences.py:31-72):
    thresh=min_start, target_d, use_target, align_meth ('0' rigid / '1' rot+loc+scale),
    stride=round(1/sample_fraction); plus reject_fraction (new) and sharding.
    """

    def __init__(self, *, thresh=0.5, target_d=0.01, use_target=True, align_meth="0", stride=1,
                 reject_fraction=0.0, device=0, rank=0, world_size=1, nn_mode=N.NN_BRUTEFORCE,
                 nn_variant=0, shard_mode=N.SHARD_BASE):
        self._lib = N.lib()
        self._h = C.c_void_p()
        self.params = N.IcpParams()
        self._fill_params(thresh, target_d, use_target, align_meth, stride, reject_fraction,
                          device, rank, world_size, nn_mode, nn_variant)
        self.params.shard_mode = int(shard_mode)
        rc = self._lib.icp_create(C.byref(self.params), C.byref(self._h))
        if rc != N.ICP_OK:
            msg = self._lib.icp_last_error(None).decode()
            self._h = C.c_void_p()
            raise IcpError(rc, msg)

    # ------------------------------------------------------------------ plumbing
    def _fill_params(self, thresh, target_d, use_target, align_meth, stride, reject_fraction,
                     device, rank, world_size, nn_mode, nn_variant):
        p = self.params
        p.struct_size = C.sizeof(N.IcpParams)
        p.device = int(device)
        p.thresh = float(thresh)
        p.target_d = float(target_d)
        p.reject_fraction = float(reject_fraction)
        p.stride = int(stride)
        p.use_target = 1 if use_target else 0
        p.align_meth = int(align_meth)
        p.nn_mode = int(nn_mode)
        p.rank = int(rank)
        p.world_size = int(world_size)
        p.nn_variant = int(nn_variant)

    def _check(self, rc):
        if rc != N.ICP_OK:
            raise IcpError(rc, self._lib.icp_last_error(self._h).decode())

    def update_params(self, **kw):
        p = self.params
        for k, v in kw.items():
            if k == "use_target":
                v = 1 if v else 0
            if k == "align_meth":
                v = int(v)
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        self._check(self._lib.icp_set_params(self._h, C.byref(p)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.icp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ inputs
    def set_base(self, xyz, global_offset=0, n_global=None):
        """Upload this rank's shard of base vertices (replaces BVHTree.FromObject, icp_align.py:53)."""
        xyz = _f32c(xyz, (-1, 3))
        n = xyz.shape[0]
        self._base_keepalive = xyz
        self._check(self._lib.icp_set_base(self._h, xyz.ctypes.data, n, int(global_offset),
                                           int(n if n_global is None else n_global)))

    def set_base_triangles(self, tri_xyz, global_offset=0, n_global=None):
        """Surface mode (nn_mode=NN_SURFACE): upload this rank's shard of the base mesh's triangles
        as (F, 9) corner coordinates -- the reference's real BVHTree.FromObject input."""
        tri = np.ascontiguousarray(tri_xyz, dtype=np.float32).reshape(-1, 9)
        n = tri.shape[0]
        self._check(self._lib.icp_set_base_triangles(self._h, tri.ctypes.data, n, int(global_offset),
                                                     int(n if n_global is None else n_global)))

    def set_base_triangles_sharded(self, tri_xyz_full):
        tri = np.asarray(tri_xyz_full).reshape(-1, 9)
        lo, hi = shard_range(tri.shape[0], self.params.rank, self.params.world_size)
        self.set_base_triangles(tri[lo:hi], lo, tri.shape[0])

    def set_base_sharded(self, xyz_full):
        """Upload only this rank's contiguous slice of the full base array."""
        xyz_full = np.asarray(xyz_full).reshape(-1, 3)
        lo, hi = shard_range(xyz_full.shape[0], self.params.rank, self.params.world_size)
        self.set_base(xyz_full[lo:hi], lo, xyz_full.shape[0])

    def set_moving(self, xyz, vlist=None):
        xyz = _f32c(xyz, (-1, 3))
        if vlist is not None:
            vlist = np.ascontiguousarray(vlist, dtype=np.int32)
            self._check(self._lib.icp_set_moving(self._h, xyz.ctypes.data, xyz.shape[0],
                                                 vlist.ctypes.data, vlist.shape[0]))
        else:
            self._check(self._lib.icp_set_moving(self._h, xyz.ctypes.data, xyz.shape[0], None, 0))

    def set_moving_collective(self, xyz, vlist=None):
        """Collective upload (every rank calls it with the same array): 1/world_size of the coordinates
        over this rank's host link, the rest over NVLink (ncclAllGather)."""
        xyz = _f32c(xyz, (-1, 3))
        if vlist is not None:
            vlist = np.ascontiguousarray(vlist, dtype=np.int32)
        self._check(self._lib.icp_set_moving_collective(
            self._h, xyz.ctypes.data, xyz.shape[0], None if vlist is None else vlist.ctypes.data,
            0 if vlist is None else vlist.shape[0]))

    def set_moving_sharded(self, xyz, vlist=None):
        """ICP_SHARD_QUERIES: this rank keeps only its contiguous slice of the sampled query
        positions (the stride is applied here, once, so every rank samples the same vertices)."""
        xyz = _f32c(xyz, (-1, 3))
        pos = np.arange(xyz.shape[0], dtype=np.int32) if vlist is None else np.ascontiguousarray(vlist, np.int32)
        stride = self.params.stride if self.params.stride > 1 else 1
        pos = pos[0::stride]
        lo, hi = shard_range(pos.shape[0], self.params.rank, self.params.world_size)
        self.update_params(stride=1)
        self.set_moving(xyz, pos[lo:hi])

    def set_matrices(self, mw_align, mw_base):
        a = _f32c(mw_align, (4, 4))
        b = _f32c(mw_base, (4, 4))
        self._check(self._lib.icp_set_matrices(self._h, a.ctypes.data, b.ctypes.data))

    # ------------------------------------------------------------------ multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        lib = N.lib()
        rc = lib.icp_comm_unique_id(buf)
        if rc != N.ICP_OK:
            raise IcpError(rc, lib.icp_last_error(None).decode())
        return buf.raw

    def comm_init(self, unique_id: bytes):
        assert len(unique_id) == 128
        self._check(self._lib.icp_comm_init(self._h, C.c_char_p(unique_id)))

    # ------------------------------------------------------------------ hot path
    def query_count(self) -> int:
        n = C.c_int64()
        self._check(self._lib.icp_query_count(self._h, C.byref(n)))
        return n.value

    def nn(self):
        """calc_target: (idx int32[L], d2 float32[L]) -- nearest base vertex per sampled vertex."""
        n = self.query_count()
        idx = np.empty(n, dtype=np.int32)
        d2 = np.empty(n, dtype=np.float32)
        self._check(self._lib.icp_nn(self._h, idx.ctypes.data, d2.ctypes.data))
        return idx, d2

    def make_pairs(self, calc_stats=False):
        """(A, B, d_stats) of the reference's make_pairs (general.py:257-329).  ONE search: the output
        arrays are sized for every query (row stride = capacity) and sliced to the K kept pairs."""
        cap = max(self.query_count(), 1)
        k = C.c_int64()
        A = np.empty((3, cap), dtype=np.float64)
        B = np.empty((3, cap), dtype=np.float64)
        ds = np.zeros(2, dtype=np.float64)
        self._check(self._lib.icp_make_pairs(self._h, 1 if calc_stats else 0, A.ctypes.data,
                                             B.ctypes.data, cap, C.byref(k), ds.ctypes.data))
        K = k.value
        A = np.ascontiguousarray(A[:, :K])
        B = np.ascontiguousarray(B[:, :K])
        return A, B, ([float(ds[0]), float(ds[1])] if calc_stats else None)

    def solve_pairs(self, v0, v1, scale=False):
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        v1 = np.ascontiguousarray(v1, dtype=np.float64)
        M = np.empty((4, 4), dtype=np.float64)
        self._check(self._lib.icp_solve_pairs(self._h, v0.ctypes.data, v1.ctypes.data,
                                              v0.shape[1], 1 if scale else 0, M.ctypes.data))
        return M

    def step(self):
        M = np.empty((4, 4), dtype=np.float64)
        s = N.IcpIterStats()
        self._check(self._lib.icp_step(self._h, M.ctypes.data, C.byref(s)))
        return M, IterStats(s)

    def run(self, max_iters):
        done, conv = C.c_int32(), C.c_int32()
        recs = (N.IcpIterStats * max(1, int(max_iters)))()
        self._check(self._lib.icp_run(self._h, int(max_iters), C.byref(done), C.byref(conv), recs))
        return done.value, bool(conv.value), [IterStats(recs[i]) for i in range(done.value)]

    def reset_convergence(self):
        self._check(self._lib.icp_reset_convergence(self._h))

    def matrix_world(self) -> np.ndarray:
        m = np.empty((4, 4), dtype=np.float32)
        self._check(self._lib.icp_get_matrix_world(self._h, m.ctypes.data))
        return m

    # ------------------------------------------------------------------ measurement
    def last_elapsed_ms(self) -> float:
        ms = C.c_double()
        self._check(self._lib.icp_last_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def set_profiling(self, on: bool):
        self._check(self._lib.icp_set_profiling(self._h, 1 if on else 0))

    def kernel_time_ms(self, which: int):
        ms, n = C.c_double(), C.c_int64()
        self._check(self._lib.icp_kernel_time_ms(self._h, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self) -> int:
        n = C.c_int64()
        self._check(self._lib.icp_launch_count(self._h, C.byref(n)))
        return n.value

    def synchronize(self):
        self._check(self._lib.icp_synchronize(self._h))

    def measure_fp32_peak(self) -> float:
        v = C.c_double()
        self._check(self._lib.icp_measure_fp32_peak(self._h, C.byref(v)))
        return v.value

    def measure_copy_bandwidth(self) -> float:
        v = C.c_double()
        self._check(self._lib.icp_measure_copy_bandwidth(self._h, C.byref(v)))
        return v.value

    def flush_l2(self):
        self._check(self._lib.icp_flush_l2(self._h))


class PinnedArray:
    """A numpy view over cudaHostAlloc'ed memory (uploads from it run at full PCIe rate)."""

    def __init__(self, shape, dtype):
        self._lib = N.lib()
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * self.dtype.itemsize
        p = C.c_void_p()
        rc = self._lib.icp_alloc_pinned(n, C.byref(p))
        if rc != N.ICP_OK:
            raise IcpError(rc, self._lib.icp_last_error(None).decode())
        self._ptr = p
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        if self._ptr:
            self.array = None
            self._lib.icp_free_pinned(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
