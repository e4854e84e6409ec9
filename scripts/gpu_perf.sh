#!/bin/bash
# Perf-only GPU call: quick parity smoke, default bench line, A/B of tuning variants, ncu captures.
tag=${1:-perf}
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.txt | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
b=json.load(open('gpurun_out/${tag}_bench.json'))
print('MAIN cfg3', round(b['value']), 'e2e', round(b['e2e']['value']), 'B/chk', round(b['roofline']['alg_bytes_per_check']), 'mism', b['cpu_baseline']['parity_mismatches_vs_gpu'])
for k,v in b['configs'].items(): print('MAIN', k, round(v['value']), 'e2e', round(v['e2e']['value']), 'B/chk', round(v['roofline']['alg_bytes_per_check']), 'mism', v['cpu_baseline']['parity_mismatches_vs_gpu'])
PY
for v in spicedb-kubeapi-proxy_b200/variants/libzgpu_*.so; do
  [ -f "$v" ] || continue
  n=$(basename $v .so)
  for wl in cfg3 cfg4; do
    ZGPU_LIB=$PWD/$v timeout 400 python bench.py --workload $wl --configs '' --no-cpu-baseline --sustain-s 0 --steps 20 --warmup 5 > gpurun_out/${tag}_ab_${n}_${wl}.json 2> gpurun_out/${tag}_ab_${n}_${wl}.err
    python -c "import json; b=json.load(open('gpurun_out/${tag}_ab_${n}_${wl}.json')); print('AB $n $wl', round(b['value']), round(b['e2e']['value']), round(b['roofline']['alg_bytes_per_check']))"
  done
done
for wl in cfg3 cfg4; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_${wl}.csv \
    python bench.py --workload $wl --configs '' --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0 > /dev/null 2> gpurun_out/${tag}_ncu_list_${wl}.err
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:check_kernel -s 5 -c 1 -f -o gpurun_out/${tag}_prof_${wl} \
    python bench.py --workload $wl --configs '' --steps 2 --warmup 3 --no-cpu-baseline --sustain-s 0 > /dev/null 2> gpurun_out/${tag}_ncu_full_${wl}.err
done
ls -la gpurun_out/${tag}_prof_* 2>/dev/null
