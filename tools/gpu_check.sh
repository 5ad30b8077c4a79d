#!/bin/bash
# Round-end style check on the GPU box: GPU tests, smoke, both bench arms, the ncu launch list of the
# default bench command and one --set full capture per dominant kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
BENCH_NCU="python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary"

( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?  $(tail -1 $O/pytest_gpu.log)"
python __graft_entry__.py --smoke > $O/smoke.log 2>&1
echo "smoke rc=$?  $(tail -1 $O/smoke.log)"

python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_c3.json 2> $O/bench_ref_c3.err
echo "reference arm rc=$?"
( time python bench.py ) > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err
echo "bench rc=$?  $(tail -3 $O/bench_c3_n1.err | tr '\n' ' ')"
python bench.py --nn-mode bvh --no-cpu > $O/bench_c3_bvh_n1.json 2> $O/bench_c3_bvh_n1.err
python bench.py --workload c3s --no-cpu > $O/bench_c3s_n1.json 2> $O/bench_c3s_n1.err
echo "bvh/surface bench rc=$?"

ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file $O/launches_c3_default.csv $BENCH_NCU > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nn_bruteforce_slab -s 3 -c 1 \
    -f -o $O/ncu_slab $BENCH_NCU > $O/ncu_slab.log 2>&1
ncu -i $O/ncu_slab.ncu-rep --page details --csv > $O/ncu_details_slab.csv 2>/dev/null
ncu -i $O/ncu_slab.ncu-rep --page raw --csv > $O/ncu_raw_slab.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_nn_bvh -s 4 -c 1 \
    -f -o $O/ncu_bvh $BENCH_NCU --nn-mode bvh > $O/ncu_bvh.log 2>&1
ncu -i $O/ncu_bvh.ncu-rep --page details --csv > $O/ncu_details_bvh.csv 2>/dev/null
ncu -i $O/ncu_bvh.ncu-rep --page raw --csv > $O/ncu_raw_bvh.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_nn_surface -s 4 -c 1 \
    -f -o $O/ncu_surf $BENCH_NCU --workload c3s > $O/ncu_surf.log 2>&1
ncu -i $O/ncu_surf.ncu-rep --page details --csv > $O/ncu_details_surf.csv 2>/dev/null
ncu -i $O/ncu_surf.ncu-rep --page raw --csv > $O/ncu_raw_surf.csv 2>/dev/null
ls -la $O
