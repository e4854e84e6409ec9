#!/bin/bash
# One GPU call: smoke, parity suite, the default bench line, and (when the suite is green) the ncu captures
# the roofline numbers come from. Outputs under gpurun_out/<tag>_*.
tag=${1:-ci}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${tag}_smoke.txt
( ZGPU_FULLSIZE_SCALE=${ZGPU_FULLSIZE_SCALE:-0.25} timeout 2400 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider --durations=12 2>&1 | tail -150 ) > gpurun_out/${tag}_pytest.txt
tail -15 gpurun_out/${tag}_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err; cut -c1-3000 gpurun_out/${tag}_bench.json
# A/B of tuning variants (scripts/build_variant.sh), cfg3 only, kernel numbers only
for v in spicedb-kubeapi-proxy_b200/variants/libzgpu_*.so; do
  [ -f "$v" ] || continue
  n=$(basename $v .so)
  ZGPU_LIB=$PWD/$v timeout 300 python bench.py --configs '' --no-cpu-baseline --sustain-s 0 --steps 20 --warmup 5 > gpurun_out/${tag}_ab_${n}.json 2> gpurun_out/${tag}_ab_${n}.err
  python -c "import json,sys; b=json.load(open('gpurun_out/${tag}_ab_${n}.json')); print('$n', b['value'], b['e2e']['value'], b['roofline']['alg_bytes_per_check'])"
done
if true; then
  for wl in cfg3 cfg4; do
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_${wl}.csv \
      python bench.py --workload $wl --configs '' --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > /dev/null 2> gpurun_out/${tag}_ncu_list_${wl}.err
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:check_kernel -s 5 -c 1 -f -o gpurun_out/${tag}_prof_${wl} \
      python bench.py --workload $wl --configs '' --steps 2 --warmup 3 --no-cpu-baseline --sustain-s 0 > /dev/null 2> gpurun_out/${tag}_ncu_full_${wl}.err
  done
  ls -la gpurun_out/${tag}_prof_* 2>/dev/null
fi
