#!/bin/bash
# Multi-GPU call (gpurun --gpus N): one-handle engine test, the bench under torchrun (replicas + sharded cfg4 leg),
# config 5 replayed with one engine owning every GPU.
tag=${1:-multi}; n=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${tag}_smi.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout=600 -p no:cacheprovider -k "one_engine_owning" --tb=long 2>&1 | tail -60 ) > gpurun_out/${tag}_pytest.txt
tail -5 gpurun_out/${tag}_pytest.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29571 \
  bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
echo "bench rc=$?"; tail -5 gpurun_out/${tag}_bench_n${n}.err; cut -c1-600 gpurun_out/${tag}_bench_n${n}.json
python -c "
import json; b=json.load(open('gpurun_out/${tag}_bench_n${n}.json')); print('value', b['value'], 'e2e', b['e2e']['value']); print({k:(v.get('value'), v.get('error'), v.get('mismatches_vs_replica')) for k,v in b.get('configs',{}).items()})"
timeout 1500 python scripts/cfg5_replay.py --devices $n --clients 1000 --scale 1.0 --rounds 2 > gpurun_out/${tag}_cfg5_n${n}.json 2> gpurun_out/${tag}_cfg5_n${n}.err
echo "cfg5 rc=$?"; tail -3 gpurun_out/${tag}_cfg5_n${n}.err; cut -c1-1500 gpurun_out/${tag}_cfg5_n${n}.json
