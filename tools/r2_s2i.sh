#!/bin/bash
# MFCC tests + C5 device-only timing (no ncu)
tag=${1:-s2i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "mfcc or golden or smoke or pcm16 or nan or real_speech or host_pipeline or multi" > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${tag}_tests.log | head -20
timeout 300 python bench.py --config c5 --device-only --steps 20 --warmup 3 2>&1 | cut -c1-300 | tee gpurun_out/${tag}_c5.txt
