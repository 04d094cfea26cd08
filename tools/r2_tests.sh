#!/bin/bash
# the -m gpu suite with the failures' detail kept: tools/r2_tests.sh [tag]
tag=${1:-t}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${tag}_tests.log | head -40
