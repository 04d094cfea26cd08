#!/bin/bash
# session-2 run b: roots tests + C3 device-only timing (approximate MUFU ops in the fp32 Laguerre step)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_formants.py tests/test_gpu_round2.py tests/test_gpu_full_size.py tests/test_gpu_real_speech.py -m gpu -q -p no:cacheprovider > gpurun_out/s2b_tests.log 2>&1
echo "tests rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/s2b_tests.log | head
timeout 300 python bench.py --config c3 --device-only --steps 10 --warmup 3 > gpurun_out/s2b_c3.txt 2>&1
cut -c1-600 gpurun_out/s2b_c3.txt
timeout 300 python tools/bench_formants.py > gpurun_out/s2b_stages.txt 2>&1; cat gpurun_out/s2b_stages.txt
