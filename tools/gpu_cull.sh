#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -x -q -m gpu -k "seeded_passes_stay_exact or c5 or nn_bit_exact or reject or c3_seeded" 2>&1 | tail -3
for k in 1 2 3 16; do ICP_Q16_HARD_CLASS=$k python bench.py --workload c5 --steps 3 --no-cpu --no-secondary --no-c4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 hard_class $k', d['value'], d['ms_per_step'], d['pairs_kept_last'])"; done
for k in 1 3; do ICP_Q16_HARD_CLASS=$k VARIANTS=71 python tools/exp_q16.py; done
VARIANTS=77,78 python tools/exp_q16.py
N=125000 VARIANTS=71 python tools/exp_q16.py
