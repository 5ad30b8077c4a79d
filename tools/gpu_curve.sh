#!/bin/bash
# Hilbert (default build) vs Morton (tools/_build/libicp_morton.so) ordering: tree search, brute force, surface
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for l in morton hilbert; do
  if [ $l = morton ]; then export ICP_B200_LIB=$PWD/tools/_build/libicp_morton.so; else unset ICP_B200_LIB; fi
  echo "curve $l"
  VARIANTS=0 python tools/exp_tree.py
  VARIANTS=71 python tools/exp_q16.py
  python bench.py --workload c3s --no-cpu --no-secondary --no-c4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3s', d['value'], d['ms_per_step'])"
  python bench.py --workload c4 --nn-mode bvh --no-cpu --no-secondary --no-c4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4 bvh', d['value'], d['ms_per_step'])"
done
