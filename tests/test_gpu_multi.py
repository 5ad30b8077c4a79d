"""GPU multi-rank test: N processes (one per GPU) with the base sharded and resolved through the
library's own NCCL all-reduces must reproduce the single-GPU result -- indices bit-exact, pose to
1e-12 relative on the fp64 moments path.  Skipped with < 2 GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, %r)
from object_alignment_b200 import IcpEngine, meshgen
from object_alignment_b200 import dist as icpdist
rank, world, local = icpdist.env_rank_world()
dist = icpdist.init_rendezvous("gloo")
cfg = meshgen.config_cloud(60_000, 90_001, 51, 52)
rf = float(os.environ.get("ICP_RF", "0"))
shard = int(os.environ.get("ICP_SHARD", "0"))
eng = IcpEngine(thresh=0.5, use_target=True, target_d=1e-4, device=local, rank=rank, world_size=world,
                reject_fraction=rf, nn_mode=int(os.environ.get("ICP_NN_MODE", "0")), shard_mode=shard)
if shard == 1:
    eng.set_base(cfg["base"])                    # base replicated, queries sliced
    icpdist.connect_engine(eng, dist)
    eng.set_moving_sharded(cfg["moving"])
else:
    eng.set_base_sharded(cfg["base"])
    icpdist.connect_engine(eng, dist)
    if os.environ.get("ICP_COLLECTIVE") == "1" and world > 1:
        eng.set_moving_collective(cfg["moving"])  # 1/world of the coordinates per host link + ncclAllGather
    else:
        eng.set_moving(cfg["moving"])
eng.set_matrices(np.eye(4), np.eye(4))
idx, d2 = eng.nn()
done, conv, recs = eng.run(12)
mw = eng.matrix_world()
# seeded pass under the final pose, after a wild pose jump (the global pruning bound must stay valid:
# it is the previous winner's distance plus the query's displacement), and after new coordinates
idx_f, d2_f = eng.nn()
jump = meshgen.rigid_matrix(140.0, (0.3, -0.2, 0.4), axis=(0.1, -1.0, 0.3))
eng.set_matrices(jump, np.eye(4))
idx_j, d2_j = eng.nn()
other = (cfg["moving"][::-1] * np.float32(0.8)).astype(np.float32)
if shard == 1:
    eng.set_moving_sharded(other)
elif os.environ.get("ICP_COLLECTIVE") == "1" and world > 1:
    eng.set_moving_collective(other)
else:
    eng.set_moving(other)
idx_o, d2_o = eng.nn()
if rank == 0:
    np.savez(os.environ["ICP_OUT"], idx=idx, d2=d2, mw=mw, done=done, conv=conv,
             K=[r.pairs for r in recs], M=[r.M for r in recs], idx_f=idx_f, d2_f=d2_f, idx_j=idx_j, d2_j=d2_j,
             idx_o=idx_o, d2_o=d2_o)
dist.barrier()
eng.close()
""" % ROOT

SURFACE_WORKER = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, %r)
import object_alignment_b200 as oab
from object_alignment_b200 import meshgen
from object_alignment_b200 import dist as icpdist
from oracle import mathutils_shim as ms
rank, world, local = icpdist.env_rank_world()
dist = icpdist.init_rendezvous("gloo")
cfg = meshgen.config_cloud(30_000, 50_000, 61, 62)
align, base = ms.Object("align", cfg["moving"]), ms.Object("base", cfg["base"])
mode = int(os.environ.get("ICP_NN_MODE", "1"))
nn = oab.VertexNN.FromObject(base, None, device=local, rank=rank, world_size=world, nn_mode=mode)
nn.connect(dist)
vlist = list(range(0, 30_000, 1))
Ms = []
for it in range(5):
    M, st = oab.icp_step(align, base, nn, vlist, 0.5, 2, calc_stats=False)   # the function surface, stride 2
    Ms.append(M)
if rank == 0:
    np.savez(os.environ["ICP_OUT"], M=np.array(Ms), mw=align.matrix_world.a, shard_mode=nn.engine.params.shard_mode)
dist.barrier()
nn.close()
""" % ROOT


def _gpu_count():
    from object_alignment_b200 import _native
    return _native.lib().icp_device_count()


def _torchrun(script, w, env, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(w),
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    subprocess.run(cmd, check=True, env=env, timeout=600, cwd=ROOT)


@pytest.mark.parametrize("nn_mode", [0, 1])
def test_function_surface_shards_by_search_kind(tmp_path, nn_mode):
    """VertexNN picks the split itself: base shards for brute force, query shards for the tree search;
    icp_step through the reference-shaped surface gives the single-GPU pose either way."""
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(SURFACE_WORKER)
    outs = {}
    for w in (1, 2):
        out = tmp_path / ("out%d.npz" % w)
        _torchrun(script, w, dict(os.environ, ICP_OUT=str(out), ICP_NN_MODE=str(nn_mode)), 29650 + w)
        outs[w] = np.load(out)
    assert int(outs[2]["shard_mode"]) == (0 if nn_mode == 0 else 1)
    np.testing.assert_allclose(outs[1]["M"], outs[2]["M"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(outs[1]["mw"], outs[2]["mw"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("rf,nn_mode,shard,collective", [(0.0, 0, 0, 0), (0.2, 0, 0, 0), (0.0, 0, 0, 1), (0.0, 1, 0, 0),
                                                         (0.2, 1, 0, 0), (0.0, 1, 1, 0)])
def test_sharded_base_matches_single_gpu(tmp_path, rf, nn_mode, shard, collective):
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 4)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    outs = {}
    for w in (1, world):
        out = tmp_path / ("out%d.npz" % w)
        env = dict(os.environ, ICP_OUT=str(out), ICP_RF=str(rf), ICP_NN_MODE=str(nn_mode),
                   ICP_SHARD=str(shard if w > 1 else 0), ICP_COLLECTIVE=str(collective))
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(w),
               "--master-addr", "127.0.0.1", "--master-port", str(29600 + w), str(script)]
        subprocess.run(cmd, check=True, env=env, timeout=600, cwd=ROOT)
        outs[w] = np.load(out)
    a, b = outs[1], outs[world]
    for tag in ("", "_f", "_j", "_o"):
        ia, da, ib, db = a["idx" + tag], a["d2" + tag], b["idx" + tag], b["d2" + tag]
        if shard == 1:      # rank 0 holds only its slice of the queries: compare that slice
            n0 = ib.shape[0]
            ia, da = ia[:n0], da[:n0]
        if tag == "_f" and not np.array_equal(a["mw"], b["mw"]):
            continue        # the final poses differ in the last bit (fp64 summation order): different queries
        assert np.array_equal(ia, ib) and np.array_equal(da.view(np.uint32), db.view(np.uint32)), tag
    assert int(a["done"]) == int(b["done"]) and bool(a["conv"]) == bool(b["conv"])
    assert a["K"].tolist() == b["K"].tolist()
    np.testing.assert_allclose(a["M"], b["M"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(a["mw"], b["mw"], rtol=0, atol=1e-6)
