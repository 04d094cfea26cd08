#!/bin/bash
# session-2 run h: pitch tests + C4 device-only timing (occupancy variants of the lag / refine kernels)
tag=${1:-s2h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pitch.py tests/test_gpu_round2.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider -k "pitch" > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${tag}_tests.log | head
timeout 600 python bench.py --config c4 --device-only --steps 10 --warmup 3 2>&1 | cut -c1-600 | tee gpurun_out/${tag}_c4.txt
