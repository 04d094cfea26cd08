#!/bin/bash
# round-2 GPU run 1: the whole -m gpu suite, then the default bench line and the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_run1_smi.txt 2>&1
nproc >> gpurun_out/r2_run1_smi.txt; free -g >> gpurun_out/r2_run1_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_run1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_run1_tests.log
tail -5 gpurun_out/r2_run1_tests.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_run1_bench.json 2> gpurun_out/r2_run1_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_run1_bench.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2_run1_ref.json 2> gpurun_out/r2_run1_ref.err
echo "ref rc=$?"; tail -3 gpurun_out/r2_run1_ref.err
