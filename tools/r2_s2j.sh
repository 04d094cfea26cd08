#!/bin/bash
# final robustness pass: the lane5 MFCC tests five times over, then memcheck / racecheck / synccheck on them
mkdir -p gpurun_out
for i in 1 2 3 4 5; do timeout 600 python -m pytest tests/test_gpu_mfcc_waves.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider -k "mfcc" 2>&1 | tail -1; done
: > gpurun_out/r2f_sanitizer_lane5.txt
for tool in memcheck racecheck synccheck; do
  echo "## $tool" >> gpurun_out/r2f_sanitizer_lane5.txt
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_mfcc_waves.py -m gpu -q -p no:cacheprovider -k "lane5" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Hazard|Invalid" | head -20 >> gpurun_out/r2f_sanitizer_lane5.txt
done
cat gpurun_out/r2f_sanitizer_lane5.txt
