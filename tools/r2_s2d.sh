#!/bin/bash
# session-2 run d: mfcc_lane5_kernel — MFCC tests, then C5 device-only timing old vs new
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "mfcc or golden or smoke or pcm16 or nan or real_speech or host_pipeline or multi" > gpurun_out/s2d_tests.log 2>&1
echo "tests rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/s2d_tests.log | head -20
for v in 0 1; do
  echo "== VBX_MFCC_LANE5=$v" >> gpurun_out/s2d_c5.txt
  VBX_MFCC_LANE5=$v timeout 300 python bench.py --config c5 --device-only --steps 20 --warmup 3 >> gpurun_out/s2d_c5.txt 2>&1
done
cut -c1-500 gpurun_out/s2d_c5.txt
