#!/bin/bash
# usage: tools/run_scaling.sh <workload> <steps> <N...>   (run on the GPU box)
W=$1; S=$2; shift 2
for N in "$@"; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps $S --warmup 3 --workload $W --no-cpu --no-secondary 2>gpurun_out/scale_${W}_n$N.err | tail -1 > gpurun_out/scale_${W}_n$N.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps $S --warmup 3 --workload $W 2>gpurun_out/scale_${W}_n$N.err | tail -1 > gpurun_out/scale_${W}_n$N.json
  fi
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_${W}_n$N.json"))
    print("${W} N=$N", "it/s=%.3f ms/step=%.3f e2e=%.3f nn_ms=%.3f coll_ms=%.3f frac=%.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_ms"], d["kernel_ms_total"]["collectives"]/d["steps"], d["roofline"]["frac"]))
except Exception as e:
    print("${W} N=$N FAILED", e); print(open("gpurun_out/scale_${W}_n$N.err").read()[-1500:])
PY
done
