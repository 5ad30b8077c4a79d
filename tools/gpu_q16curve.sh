#!/bin/bash
# order of the q16 tiles: Morton (default) vs Hilbert (ICP_Q16_TILE_CURVE=hilbert); slots and tree tiles stay Hilbert
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -x -q -m gpu -k "nn_bit_exact or c3_seeded or (seeded_passes_stay_exact and (0] or 71 or 50)) or class_order" 2>&1 | tail -2
for c in morton hilbert morton hilbert; do
  export ICP_Q16_TILE_CURVE=$c; echo "q16 tile curve $c"
  for w in c2 c3; do python bench.py --workload $w --no-cpu --no-secondary --no-c4 --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], d['ms_per_step'])"; done
done
