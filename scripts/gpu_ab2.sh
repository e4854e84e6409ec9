#!/bin/bash
# A/B call: kernel variants of the cooperative two-level meet, and the batcher (targeted wake-ups, parallel gather)
# against the previous library, on one GPU.
tag=${1:-ab2}
mkdir -p gpurun_out
V=spicedb-kubeapi-proxy_b200/variants
ab() { # name workload
  ZGPU_LIB=$PWD/$V/libzgpu_$1.so timeout 400 python bench.py --workload $2 --configs '' --no-cpu-baseline --sustain-s 0 --no-sharded --steps 30 --warmup 5 > gpurun_out/${tag}_ab_$1_$2.json 2> gpurun_out/${tag}_ab_$1_$2.err
  python -c "import json; b=json.load(open('gpurun_out/${tag}_ab_$1_$2.json')); print('AB $1 $2', round(b['value']), round(b['e2e']['value']), round(b['roofline']['alg_bytes_per_check']))"
}
for n in base filt split filtsplit base filtsplit; do ab $n cfg3; done
for n in base filtsplit; do ab $n cfg4; done
( timeout 900 python -m pytest tests/test_cabi_harness.py tests/test_gpu_parity.py -m gpu -q -x --timeout=600 -p no:cacheprovider -k "harness or concurrent or coalesc or lookups or multi_device or batch" 2>&1 | tail -8 ) > gpurun_out/${tag}_pytest_batcher.txt
tail -3 gpurun_out/${tag}_pytest_batcher.txt
for lib in base new; do
  if [ $lib = base ]; then export ZGPU_LIB=$PWD/$V/libzgpu_base.so; else unset ZGPU_LIB; fi
  timeout 600 python scripts/cfg5_replay.py --scale 0.1 --devices 1 --rounds 4 > gpurun_out/${tag}_cfg5_${lib}_s0.1.json 2> gpurun_out/${tag}_cfg5_${lib}_s0.1.err
  echo "cfg5 $lib rc=$?"; python -c "
import json; b=json.load(open('gpurun_out/${tag}_cfg5_${lib}_s0.1.json')); print('CFG5 $lib lists/s', round(b['postfilter']['filtered_lists_per_s']), 'lookups/s', round(b['prefilter']['lookups_per_s']), 'mixed', round(b['mixed']['filtered_lists_per_s']))"
done
unset ZGPU_LIB
timeout 600 python scripts/list_replay.py --clients 256 --rounds 2 > gpurun_out/${tag}_list_replay.json 2> gpurun_out/${tag}_list_replay.err; echo "list_replay rc=$?"; cut -c1-600 gpurun_out/${tag}_list_replay.json
