#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_real_speech.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_run6_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2_run6_tests.log
timeout 1500 python tools/parity_scale.py > gpurun_out/r2_parity_scale.txt 2>&1
echo "parity rc=$?"; cat gpurun_out/r2_parity_scale.txt
