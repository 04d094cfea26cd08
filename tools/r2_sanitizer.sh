#!/bin/bash
# compute-sanitizer over the -m gpu suite (racecheck / memcheck / synccheck): tools/r2_sanitizer.sh > gpurun_out/r2_sanitizer.txt
mkdir -p gpurun_out
echo "# compute-sanitizer (round 2 kernels): for tool in racecheck memcheck synccheck: compute-sanitizer --tool \$tool python -m pytest tests -m gpu -q"
for tool in memcheck racecheck synccheck; do
  echo "## $tool"
  timeout 1500 compute-sanitizer --tool $tool python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -E "passed|failed|error|ERROR SUMMARY|RACECHECK SUMMARY|Hazard|Invalid|=========  " | head -40
done
