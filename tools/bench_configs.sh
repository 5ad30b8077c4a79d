#!/bin/bash
# All BASELINE.json configurations on ONE GPU, both NN modes (run on the GPU box); lines land in gpurun_out/.
O=gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-secondary "$@" 2>$O/cfg_$name.err | tail -1 > $O/cfg_$name.json; python - <<PY
import json
try:
    d=json.load(open("$O/cfg_$name.json"))
    print("$name it/s=%.3f ms=%.3f e2e=%.3f nn_ms=%.4f launches=%d m_eff=%d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_ms"], d["gpu_launches"], d["config"]["m_eff"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
run c2_bf --workload c2
run c2_bvh --workload c2 --nn-mode bvh
run c2s --workload c2s
run c4_bf --workload c4 --steps 3
run c4_bvh --workload c4 --nn-mode bvh
run c5_bf --workload c5 --steps 3
run c5_bvh --workload c5 --nn-mode bvh --steps 5
