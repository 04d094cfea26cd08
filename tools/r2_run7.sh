#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pitch_lag64 -c 1 -o gpurun_out/prof_lag64_v1 python bench.py --config c4 --utts 48 --steps 1 --warmup 0 --device-only > gpurun_out/ncu_lag64_v1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pitch_refine8q -c 1 -o gpurun_out/prof_refine2_v1 python bench.py --config c4 --utts 48 --steps 1 --warmup 0 --device-only > gpurun_out/ncu_refine2_v1.log 2>&1
ls -la gpurun_out/*.ncu-rep
python -m pytest tests/test_gpu_pitch.py tests/test_gpu_round2.py tests/test_gpu_real_speech.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
