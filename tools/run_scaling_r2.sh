#!/bin/bash
# usage: tools/run_scaling_r2.sh <N> [steps]   (run on the GPU box under gpurun --gpus N)
# Default (brute-force, base sharded; carries parity + c4) and tree (queries sharded) lines at N GPUs.
N=$1; S=${2:-20}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N))"
$TR bench.py --gpus $N --steps $S --warmup 5 2>$O/r2_scale_c3_n$N.err | tail -1 > $O/r2_scale_c3_n$N.json
$TR bench.py --gpus $N --steps $S --warmup 5 --nn-mode bvh --no-c4 2>$O/r2_scale_c3_bvh_n$N.err | tail -1 > $O/r2_scale_c3_bvh_n$N.json
python - <<PY
import json
for f in ["r2_scale_c3_n$N", "r2_scale_c3_bvh_n$N"]:
    try:
        d = json.load(open("$O/" + f + ".json"))
        print(f, "it/s=%.3f ms=%.4f e2e=%.3f parity=%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("parity") or {}).get("ok")), d.get("per_rank"))
        if d.get("c4"): print("  c4 it/s=%.3f ms=%.3f parity=%s" % (d["c4"]["value"], d["c4"]["ms_per_step"], (d["c4"].get("parity") or {}).get("ok")))
    except Exception as e:
        print(f, "FAILED", e); print(open("$O/" + f + ".err").read()[-1500:])
PY
