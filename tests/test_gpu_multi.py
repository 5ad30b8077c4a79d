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
    eng.set_moving(cfg["moving"])
eng.set_matrices(np.eye(4), np.eye(4))
idx, d2 = eng.nn()
done, conv, recs = eng.run(12)
mw = eng.matrix_world()
if rank == 0:
    np.savez(os.environ["ICP_OUT"], idx=idx, d2=d2, mw=mw, done=done, conv=conv,
             K=[r.pairs for r in recs], M=[r.M for r in recs])
dist.barrier()
eng.close()
""" % ROOT


def _gpu_count():
    from object_alignment_b200 import _native
    return _native.lib().icp_device_count()


@pytest.mark.parametrize("rf,nn_mode,shard", [(0.0, 0, 0), (0.2, 0, 0), (0.0, 1, 0), (0.2, 1, 0), (0.0, 1, 1)])
def test_sharded_base_matches_single_gpu(tmp_path, rf, nn_mode, shard):
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
                   ICP_SHARD=str(shard if w > 1 else 0))
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(w),
               "--master-addr", "127.0.0.1", "--master-port", str(29600 + w), str(script)]
        subprocess.run(cmd, check=True, env=env, timeout=600, cwd=ROOT)
        outs[w] = np.load(out)
    a, b = outs[1], outs[world]
    if shard == 1:      # rank 0 holds only its slice of the queries: compare that slice
        n0 = b["idx"].shape[0]
        assert np.array_equal(a["idx"][:n0], b["idx"]) and np.array_equal(a["d2"][:n0], b["d2"])
    else:
        assert np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["d2"], b["d2"])
    assert int(a["done"]) == int(b["done"]) and bool(a["conv"]) == bool(b["conv"])
    assert a["K"].tolist() == b["K"].tolist()
    np.testing.assert_allclose(a["M"], b["M"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(a["mw"], b["mw"], rtol=0, atol=1e-6)
