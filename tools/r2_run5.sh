#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_run5_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r2_run5_tests.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_run5_bench.json 2> gpurun_out/r2_run5_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_run5_bench.err
ncu --set full --clock-control none --import-source on -k regex:lpc_fusedp -c 1 -o gpurun_out/prof_lpcp_v1 env VBX_FORMANT_CHUNKS=1 python bench.py --config c3 --utts 1125 --steps 1 --warmup 0 --device-only > gpurun_out/ncu_lpcp_v1.log 2>&1
ls -la gpurun_out/*.ncu-rep
