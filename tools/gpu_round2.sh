#!/bin/bash
# Round-2 check on the GPU box (run under gpurun): GPU tests, smoke, both bench arms, the ncu launch list
# of the default bench command, one --set full capture per dominant kernel, sanitizer passes.
# Everything lands in gpurun_out/ with an ${P}_ prefix; the summaries worth keeping are copied to profiles/.
mkdir -p gpurun_out
O=gpurun_out
P=${PREFIX:-r2f}
BENCH_NCU="python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --no-c4"
WHAT=${1:-all}

if [[ $WHAT == all || $WHAT == tests ]]; then
  ( time python -m pytest tests -m gpu -x -q ) > $O/${P}_pytest_gpu.log 2>&1
  echo "pytest rc=$?  $(tail -4 $O/${P}_pytest_gpu.log | head -1)"
  python __graft_entry__.py --smoke > $O/${P}_smoke.log 2>&1
  echo "smoke rc=$?  $(tail -1 $O/${P}_smoke.log)"
fi
if [[ $WHAT == all || $WHAT == bench ]]; then
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/${P}_clocks_bench.csv &
  SMI=$!
  ( time python bench.py --steps 20 --warmup 5 ) > $O/${P}_bench_c3_n1.json 2> $O/${P}_bench_c3_n1.err
  echo "bench rc=$?  $(tail -3 $O/${P}_bench_c3_n1.err | tr '\n' ' ')"
  kill $SMI
  python bench.py --impl reference --steps 3 --warmup 1 > $O/${P}_bench_ref_c3.json 2> $O/${P}_bench_ref_c3.err
  echo "reference arm rc=$?"
  python bench.py --nn-mode bvh --no-cpu --steps 20 --warmup 5 > $O/${P}_bench_c3_bvh_n1.json 2> $O/${P}_bench_c3_bvh_n1.err
  python bench.py --workload c3s --no-cpu > $O/${P}_bench_c3s_n1.json 2> $O/${P}_bench_c3s_n1.err
  python bench.py --workload c2 --no-cpu --no-c4 > $O/${P}_bench_c2_n1.json 2> $O/${P}_bench_c2_n1.err
  echo "bvh/surface/c2 bench rc=$?"
fi
if [[ $WHAT == all || $WHAT == ncu ]]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file $O/${P}_launches_c3_default.csv $BENCH_NCU > $O/${P}_ncu_launch.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file $O/${P}_launches_c3_bvh.csv $BENCH_NCU --nn-mode bvh > $O/${P}_ncu_launch_bvh.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_nn_bruteforce_q16 -s 3 -c 1 \
      -f -o $O/${P}_ncu_q16 $BENCH_NCU > $O/${P}_ncu_q16.log 2>&1
  ncu -i $O/${P}_ncu_q16.ncu-rep --page details --csv > $O/${P}_ncu_details_k2_q16.csv 2>/dev/null
  ncu --set full --clock-control none --import-source on -k regex:k_nn_tiletree_p -s 4 -c 1 \
      -f -o $O/${P}_ncu_tiletree $BENCH_NCU --nn-mode bvh > $O/${P}_ncu_tiletree.log 2>&1
  ncu -i $O/${P}_ncu_tiletree.ncu-rep --page details --csv > $O/${P}_ncu_details_tiletree_fused.csv 2>/dev/null
  ncu --set full --clock-control none --import-source on -k regex:k_filter_moments -s 3 -c 1 \
      -f -o $O/${P}_ncu_k4 $BENCH_NCU > $O/${P}_ncu_k4.log 2>&1
  ncu -i $O/${P}_ncu_k4.ncu-rep --page details --csv > $O/${P}_ncu_details_k4.csv 2>/dev/null
  echo "ncu done: $(ls $O/*.ncu-rep | wc -l) reports"
fi
if [[ $WHAT == all || $WHAT == sanitize ]]; then
  T=tests/test_gpu_parity.py
  MEM="$T::test_nn_bit_exact_small $T::test_bvh_ragged_sizes $T::test_bruteforce_seeded_passes_stay_exact[0] $T::test_bruteforce_seeded_passes_stay_exact[62] $T::test_bruteforce_seeded_passes_stay_exact[50] $T::test_bvh_warm_start_stays_exact_over_iterations $T::test_reject_fraction_matches_definition $T::test_surface_search_matches_oracle $T::test_tree_warm_start_on_a_lattice_resolves_ties_to_the_lowest_index tests/test_gpu_baseline_configs.py::test_c5_shape_outliers_mask_reject_stride"
  ( time compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -x -q -m gpu $MEM ) > $O/${P}_memcheck.txt 2>&1
  echo "memcheck rc=$?  $(grep -E 'passed|failed|ERROR SUMMARY' $O/${P}_memcheck.txt | tail -2 | tr '\n' ' ')"
  # racecheck / synccheck: the default kernels of the three modes at small sizes (racecheck runs ~100x slower)
  RACE="$T::test_bvh_ragged_sizes[0-1025-1025] $T::test_bvh_ragged_sizes[0-4097-2049] $T::test_tree_warm_start_on_a_lattice_resolves_ties_to_the_lowest_index[0] $T::test_nn_bit_exact_small[0] $T::test_surface_search_matches_oracle[icosphere]"
  ( time compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -x -q -m gpu $RACE ) > $O/${P}_racecheck.txt 2>&1
  echo "racecheck rc=$?  $(grep -E 'passed|failed|RACECHECK SUMMARY' $O/${P}_racecheck.txt | tail -2 | tr '\n' ' ')"
  ( time compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest -x -q -m gpu $RACE $T::test_bruteforce_seeded_passes_stay_exact[0] ) > $O/${P}_synccheck.txt 2>&1
  echo "synccheck rc=$?  $(grep -E 'passed|failed|ERROR SUMMARY' $O/${P}_synccheck.txt | tail -2 | tr '\n' ' ')"
fi
ls $O | grep ${P}_ | wc -l
