#!/bin/bash
# one `ncu --set full` capture of lpc_roots_pair_kernel on one find_formants chunk of the C3 shape
tag=${1:-roots_v2}
mkdir -p gpurun_out
VBX_FORMANT_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:lpc_roots_pair -c 1 -f -o gpurun_out/prof_$tag python bench.py --config c3 --utts 1125 --steps 1 --warmup 0 --no-cpu --device-only > gpurun_out/ncu_$tag.log 2>&1
bash tools/ncu_summary.sh gpurun_out/prof_$tag.ncu-rep gpurun_out/${tag}_full.txt
head -34 gpurun_out/${tag}_full.txt
