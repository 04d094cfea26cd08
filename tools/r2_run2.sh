#!/bin/bash
# round-2 GPU run 2: new tests, then C3 device-only timing for several plans of the aligned LPC kernel and chunk counts
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_formants.py tests/test_gpu_lpc.py tests/test_gpu_host_pipeline.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_run2_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2_run2_tests.log
: > gpurun_out/r2_run2_c3.txt
run() { echo "== $*" >> gpurun_out/r2_run2_c3.txt; env "$@" timeout 300 python bench.py --config c3 --device-only --steps 10 --warmup 3 >> gpurun_out/r2_run2_c3.txt 2>&1; }
run VBX_LPCA=0 VBX_FORMANT_CHUNKS=1
run VBX_LPCA=0
run VBX_FORMANT_CHUNKS=1
run VBX_LPCA_PLAN=32:8
run VBX_LPCA_PLAN=16:8
run VBX_LPCA_PLAN=32:4
run VBX_LPCA_PLAN=24:8
run VBX_LPCA_PLAN=16:16
run VBX_LPCA_PLAN=64:4
run VBX_LPCA_PLAN=32:8 VBX_FORMANT_CHUNKS=4
run VBX_LPCA_PLAN=32:8 VBX_FORMANT_CHUNKS=16
cat gpurun_out/r2_run2_c3.txt | cut -c1-400
