#!/bin/bash
# Final GPU call of a round: the -m gpu suite as the driver runs it, the full-size tests at the real size, the default
# bench line, the ncu captures the roofline numbers come from, compute-sanitizer over a subset.
tag=${1:-final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.txt 2>&1; echo "smoke rc=$?"
( timeout 1800 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider --durations=12 2>&1 | tail -60 ) > gpurun_out/${tag}_pytest.txt
tail -4 gpurun_out/${tag}_pytest.txt
( ZGPU_FULLSIZE_SCALE=1.0 timeout 1500 python -m pytest tests/test_gpu_fullsize.py -q --timeout=900 -p no:cacheprovider -k "bit_exact or thousand_update" 2>&1 | tail -30 ) > gpurun_out/${tag}_pytest_fullsize_1.0.txt
tail -3 gpurun_out/${tag}_pytest_fullsize_1.0.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
for wl in cfg3 cfg4; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_${wl}.csv \
    python bench.py --workload $wl --configs '' --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > /dev/null 2> gpurun_out/${tag}_ncu_list_${wl}.err
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:check_kernel -s 5 -c 1 -f -o gpurun_out/${tag}_prof_${wl} \
    python bench.py --workload $wl --configs '' --steps 2 --warmup 3 --no-cpu-baseline --sustain-s 0 > /dev/null 2> gpurun_out/${tag}_ncu_full_${wl}.err
done
timeout 600 python scripts/write_bench.py --writes 6 > gpurun_out/${tag}_write_bench.json 2> gpurun_out/${tag}_write_bench.err; cut -c1-400 gpurun_out/${tag}_write_bench.json
K="two_level_meet or golden or fixed_schemas or depth_cap or expiration or incremental_publish_equals or known_divergence or device_resident_sharded_store_depth"
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_new_paths.py -q -x -p no:cacheprovider -k "$K" 2>&1 | tail -25 ) > gpurun_out/${tag}_compute_sanitizer_memcheck.log
tail -4 gpurun_out/${tag}_compute_sanitizer_memcheck.log
( timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_new_paths.py -q -x -p no:cacheprovider -k "two_level_meet or fixed_schemas or baseline_configs_scaled or depth_cap" 2>&1 | tail -25 ) > gpurun_out/${tag}_compute_sanitizer_racecheck.log
tail -4 gpurun_out/${tag}_compute_sanitizer_racecheck.log
# config 5 at the C ABI on this one GPU (the 8-GPU run is scripts/gpu_multi2.sh) and filtered kube lists on real bodies
for sc in 0.1 1.0; do
  timeout 600 python scripts/cfg5_replay.py --scale $sc --devices 1 --rounds 4 > gpurun_out/${tag}_cfg5_1gpu_s$sc.json 2> gpurun_out/${tag}_cfg5_1gpu_s$sc.err
  echo "cfg5 $sc rc=$?"; python -c "
import json; b=json.load(open('gpurun_out/${tag}_cfg5_1gpu_s$sc.json')); print('CFG5 $sc lists/s', round(b['postfilter']['filtered_lists_per_s']), 'lookups/s', round(b['prefilter']['lookups_per_s']), 'mixed', round(b['mixed']['filtered_lists_per_s']))"
done
timeout 600 python scripts/list_replay.py --clients 256 --rounds 2 > gpurun_out/${tag}_list_replay.json 2> gpurun_out/${tag}_list_replay.err; echo "list_replay rc=$?"; cut -c1-500 gpurun_out/${tag}_list_replay.json
