"""World-size-2 (and 3) gloo tests on CPU for the N>1 protocol of SURVEY.md section 8e:

  * base split into contiguous shards (shard_range)
  * collective 1: element-wise MIN over uint64 keys float_bits(d2) << 32 | global index
  * owner-computes moments (a rank accumulates only the queries whose winner lies in its shard)
  * collective 2: SUM over the 20 moment doubles

The per-rank compute is done by the CPU oracle here (no GPU in this container); the collectives
and the sharding helpers are the product's own (object_alignment_b200/dist.py).  The result must
equal the unsharded oracle: indices bit-exact, solved M to 1e-12.
"""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _moments(A, B, shift):
    a = A.T - shift
    b = B.T - shift
    out = np.zeros(20)
    out[0] = a.shape[0]
    out[1:4] = a.sum(0)
    out[4:7] = b.sum(0)
    out[7:16] = (b.T @ a).reshape(-1)
    out[16] = (a * a).sum()
    out[17] = (b * b).sum()
    return out


def _solve_from_moments(m, shift):
    K = m[0]
    ca, cb = m[1:4] / K, m[4:7] / K
    H = m[7:16].reshape(3, 3) - K * np.outer(cb, ca)
    u, s, vh = np.linalg.svd(H)
    R = u @ vh
    if np.linalg.det(R) < 0:
        R -= np.outer(u[:, 2], vh[2] * 2)
    M = np.identity(4)
    M[:3, :3] = R
    M[:3, 3] = (cb + shift) - R @ (ca + shift)
    return M


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from object_alignment_b200 import dist as icpdist
    from object_alignment_b200 import meshgen, shard_range
    from oracle import icp_oracle as orc

    dist = icpdist.init_rendezvous("gloo")
    cfg = meshgen.config_cloud(3000, 5001, 41, 42)
    mov, base = cfg["moving"], cfg["base"]
    lo, hi = shard_range(base.shape[0], rank, world)
    eye = np.eye(4, dtype=np.float32)
    res = orc.make_pairs_vec(mov, np.arange(3000), eye, eye, base[lo:hi], 0.5, 1, index_offset=lo,
                             return_all=True)
    keys = icpdist.allreduce_min_keys_host(orc.pack_key(res["d2"], res["idx"]), dist)
    idx, d2 = orc.unpack_key(keys)
    # owner-computes: only winners inside my shard, evaluated with the full-precision rules
    mine = (idx >= lo) & (idx < hi)
    full = orc.make_pairs_vec(mov[mine], np.arange(int(mine.sum())), eye, eye, base[lo:hi], 0.5, 1,
                              index_offset=lo, return_all=True)
    assert np.array_equal(full["idx"], idx[mine])          # the local NN of an owned query is the winner
    shift = mov[0].astype(np.float64)
    mom = icpdist.allreduce_sum_host(_moments(full["A"], full["B"], shift), dist)
    M = _solve_from_moments(mom, shift)
    tmax = icpdist.allreduce_max_scalar(float(rank), dist)
    assert tmax == world - 1
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), idx=idx, d2=d2, M=M, K=mom[0])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_protocol_equals_unsharded(world, tmp_path):
    from object_alignment_b200 import meshgen
    from oracle import icp_oracle as orc

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / "out.npz")
    cfg = meshgen.config_cloud(3000, 5001, 41, 42)
    eye = np.eye(4, dtype=np.float32)
    ref = orc.make_pairs_vec(cfg["moving"], np.arange(3000), eye, eye, cfg["base"], 0.5, 1,
                             return_all=True)
    assert np.array_equal(z["idx"], ref["idx"])
    assert np.array_equal(z["d2"].view(np.uint32), ref["d2"].view(np.uint32))
    assert int(z["K"]) == int(ref["keep"].sum())
    np.testing.assert_allclose(z["M"], orc.affine_matrix_from_points_port(ref["A"], ref["B"]),
                               rtol=0, atol=1e-12)


def _worker_queries(rank, world, port, out_dir):
    """ICP_SHARD_QUERIES protocol (tree searches): base replicated, every rank owns a contiguous
    slice of the query positions; the only exchange is the SUM over the 20 moment doubles, taken
    about the same reference point (the first moving vertex) on every rank."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from object_alignment_b200 import dist as icpdist
    from object_alignment_b200 import meshgen, shard_range
    from oracle import icp_oracle as orc

    dist = icpdist.init_rendezvous("gloo")
    cfg = meshgen.config_cloud(3001, 4000, 43, 44)
    mov, base = cfg["moving"], cfg["base"]
    lo, hi = shard_range(mov.shape[0], rank, world)
    eye = np.eye(4, dtype=np.float32)
    res = orc.make_pairs_vec(mov, np.arange(lo, hi), eye, eye, base, 0.5, 1, return_all=True)
    shift = mov[0].astype(np.float64)                     # shared reference point, not mov[lo]
    mom = icpdist.allreduce_sum_host(_moments(res["A"], res["B"], shift), dist)
    M = _solve_from_moments(mom, shift)
    if rank == 0:
        np.savez(os.path.join(out_dir, "outq.npz"), M=M, K=mom[0])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_query_sharded_protocol_equals_unsharded(world, tmp_path):
    from object_alignment_b200 import meshgen, shard_range
    from oracle import icp_oracle as orc

    # the slices tile the query positions exactly, in order
    cuts = [shard_range(3001, r, world) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == 3001 and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
    mp.spawn(_worker_queries, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / "outq.npz")
    cfg = meshgen.config_cloud(3001, 4000, 43, 44)
    eye = np.eye(4, dtype=np.float32)
    ref = orc.make_pairs_vec(cfg["moving"], np.arange(3001), eye, eye, cfg["base"], 0.5, 1, return_all=True)
    assert int(z["K"]) == int(ref["keep"].sum())
    np.testing.assert_allclose(z["M"], orc.affine_matrix_from_points_port(ref["A"], ref["B"]),
                               rtol=0, atol=1e-12)
