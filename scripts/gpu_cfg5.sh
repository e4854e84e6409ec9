#!/bin/bash
# Config 5 only (1 000 native clients, ONE engine handle owning n GPUs), the 95 M store first.
tag=${1:-cfg5}; n=${2:-8}
mkdir -p gpurun_out
for sc in 1.0 0.1; do
  timeout 400 python scripts/cfg5_replay.py --devices $n --clients 1000 --scale $sc --rounds 2 > gpurun_out/${tag}_cfg5_n${n}_s${sc}.json 2> gpurun_out/${tag}_cfg5_n${n}_s${sc}.err
  echo "cfg5 scale $sc rc=$?"; tail -2 gpurun_out/${tag}_cfg5_n${n}_s${sc}.err | cut -c1-300
  python -c "
import json; b=json.load(open('gpurun_out/${tag}_cfg5_n${n}_s${sc}.json')); print('CFG5 n$n s$sc lists/s', round(b['postfilter']['filtered_lists_per_s']), 'lookups/s', round(b['prefilter']['lookups_per_s']), 'mixed', round(b['mixed']['filtered_lists_per_s']), 'devices', b['devices'])"
done
