#!/bin/bash
# session-2 run f: whole -m gpu suite; racecheck / memcheck / synccheck on the lane5 MFCC tests
mkdir -p gpurun_out
bash tools/r2_tests.sh s2f
for tool in memcheck racecheck synccheck; do
  echo "## $tool" >> gpurun_out/s2f_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_mfcc_waves.py -m gpu -q -p no:cacheprovider -k "lane5" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Hazard|Invalid|=========  " | head -20 >> gpurun_out/s2f_sanitizer.txt
done
cat gpurun_out/s2f_sanitizer.txt
