#!/bin/bash
# tile-tree walk check: tree-mode parity tests, then the A/B of the two generations at C3
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -x -q -m gpu -k "bvh or tree or trajectory or modal or operator or stepwise" 2>&1 | tail -15 ) > gpurun_out/tt_tests.txt
tail -5 gpurun_out/tt_tests.txt
VARIANTS=${VARIANTS:-13,0} python tools/exp_tree.py 2> gpurun_out/tt_exp.err | tee gpurun_out/tt_exp.jsonl
tail -3 gpurun_out/tt_exp.err
