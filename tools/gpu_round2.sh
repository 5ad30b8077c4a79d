#!/bin/bash
# Round-2 check on the GPU box (run under gpurun): GPU tests, smoke, both bench arms, the ncu launch list
# of the default bench command, one --set full capture per dominant kernel, sanitizer passes.
# Everything lands in gpurun_out/ with an r2_ prefix; the summaries worth keeping are copied to profiles/.
mkdir -p gpurun_out
O=gpurun_out
BENCH_NCU="python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --no-c4"
WHAT=${1:-all}

if [[ $WHAT == all || $WHAT == tests ]]; then
  ( time python -m pytest tests -m gpu -x -q ) > $O/r2_pytest_gpu.log 2>&1
  echo "pytest rc=$?  $(tail -4 $O/r2_pytest_gpu.log | head -1)"
  python __graft_entry__.py --smoke > $O/r2_smoke.log 2>&1
  echo "smoke rc=$?  $(tail -1 $O/r2_smoke.log)"
fi
if [[ $WHAT == all || $WHAT == bench ]]; then
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/r2_clocks_bench.csv &
  SMI=$!
  ( time python bench.py --steps 20 --warmup 5 ) > $O/r2_bench_c3_n1.json 2> $O/r2_bench_c3_n1.err
  echo "bench rc=$?  $(tail -3 $O/r2_bench_c3_n1.err | tr '\n' ' ')"
  kill $SMI
  python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_ref_c3.json 2> $O/r2_bench_ref_c3.err
  echo "reference arm rc=$?"
  python bench.py --nn-mode bvh --no-cpu --steps 20 --warmup 5 > $O/r2_bench_c3_bvh_n1.json 2> $O/r2_bench_c3_bvh_n1.err
  python bench.py --workload c3s --no-cpu > $O/r2_bench_c3s_n1.json 2> $O/r2_bench_c3s_n1.err
  python bench.py --workload c2 --no-cpu --no-c4 > $O/r2_bench_c2_n1.json 2> $O/r2_bench_c2_n1.err
  echo "bvh/surface/c2 bench rc=$?"
fi
if [[ $WHAT == all || $WHAT == ncu ]]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file $O/r2_launches_c3_default.csv $BENCH_NCU > $O/r2_ncu_launch.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file $O/r2_launches_c3_bvh.csv $BENCH_NCU --nn-mode bvh > $O/r2_ncu_launch_bvh.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_nn_bruteforce_q16 -s 3 -c 1 \
      -f -o $O/r2_ncu_q16 $BENCH_NCU > $O/r2_ncu_q16.log 2>&1
  ncu -i $O/r2_ncu_q16.ncu-rep --page details --csv > $O/r2_ncu_details_k2_q16.csv 2>/dev/null
  ncu --set full --clock-control none --import-source on -k regex:k_nn_tiletree -s 4 -c 1 \
      -f -o $O/r2_ncu_tiletree $BENCH_NCU --nn-mode bvh > $O/r2_ncu_tiletree.log 2>&1
  ncu -i $O/r2_ncu_tiletree.ncu-rep --page details --csv > $O/r2_ncu_details_tiletree_fused.csv 2>/dev/null
  ncu --set full --clock-control none --import-source on -k regex:k_filter_moments -s 3 -c 1 \
      -f -o $O/r2_ncu_k4 $BENCH_NCU > $O/r2_ncu_k4.log 2>&1
  ncu -i $O/r2_ncu_k4.ncu-rep --page details --csv > $O/r2_ncu_details_k4.csv 2>/dev/null
  echo "ncu done: $(ls $O/*.ncu-rep | wc -l) reports"
fi
if [[ $WHAT == all || $WHAT == sanitize ]]; then
  SAN="python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k"
  SEL="test_nn_bit_exact_small or test_bvh_ragged_sizes or (seeded_passes_stay_exact and (0 or 62 or 50)) or test_bvh_warm_start or test_reject_fraction or test_surface_search_matches_oracle"
  ( time compute-sanitizer --tool memcheck --error-exitcode 9 $SAN "$SEL" ) > $O/r2_memcheck.txt 2>&1
  echo "memcheck rc=$?  $(grep -E 'passed|failed|ERROR SUMMARY' $O/r2_memcheck.txt | tail -2 | tr '\n' ' ')"
  SEL2="test_nn_bit_exact_small or (test_bvh_ragged_sizes and 1025) or (seeded_passes_stay_exact and (0 or 62)) or (test_surface_search_matches_oracle and icosphere)"
  ( time compute-sanitizer --tool racecheck --error-exitcode 9 $SAN "$SEL2" ) > $O/r2_racecheck.txt 2>&1
  echo "racecheck rc=$?  $(grep -E 'passed|failed|RACECHECK SUMMARY' $O/r2_racecheck.txt | tail -2 | tr '\n' ' ')"
  ( time compute-sanitizer --tool synccheck --error-exitcode 9 $SAN "$SEL2" ) > $O/r2_synccheck.txt 2>&1
  echo "synccheck rc=$?  $(grep -E 'passed|failed|ERROR SUMMARY' $O/r2_synccheck.txt | tail -2 | tr '\n' ' ')"
fi
ls $O | grep r2_ | wc -l
