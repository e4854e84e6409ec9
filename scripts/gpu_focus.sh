#!/bin/bash
# focused GPU run: named tests with full tracebacks
tag=${1:-focus}; shift
mkdir -p gpurun_out
( timeout 2400 python -m pytest "$@" -q --timeout=900 -p no:cacheprovider -x --tb=long 2>&1 | tail -250 ) > gpurun_out/${tag}_pytest.txt
tail -30 gpurun_out/${tag}_pytest.txt
