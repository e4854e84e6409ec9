#!/bin/bash
# Short re-validation after a late change: the whole -m gpu suite and the write bench.
tag=${1:-recheck}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest.txt
tail -3 gpurun_out/${tag}_pytest.txt
timeout 600 python scripts/write_bench.py --writes 8 > gpurun_out/${tag}_write_bench.json 2> gpurun_out/${tag}_write_bench.err; cut -c1-700 gpurun_out/${tag}_write_bench.json
