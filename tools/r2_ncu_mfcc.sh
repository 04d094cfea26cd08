#!/bin/bash
# one `ncu --set full` capture of the MFCC kernel on 360 utterances of the C5 shape; summary to gpurun_out/$1_full.txt
tag=${1:-mfcc_l5}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mfcc_ -c 1 -f -o gpurun_out/prof_$tag python bench.py --config c5 --utts 360 --steps 1 --warmup 0 --no-cpu --device-only > gpurun_out/ncu_$tag.log 2>&1
bash tools/ncu_summary.sh gpurun_out/prof_$tag.ncu-rep gpurun_out/${tag}_full.txt
head -40 gpurun_out/${tag}_full.txt
