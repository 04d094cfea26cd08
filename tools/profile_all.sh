#!/bin/bash
# GPU box: launch lists (ncu --metrics gpu__time_duration.sum) for the bench commands and one `--set full` capture per kernel.
set -x
mkdir -p gpurun_out
for c in c2 c3 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${c}.csv python bench.py --config $c --steps 2 --warmup 3 --device-only > gpurun_out/ncu_bench_${c}.log 2>&1
done
full() { # name regex command...
  local name=$1 rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -o gpurun_out/prof_${name}_final "$@" > gpurun_out/ncu_${name}_final.log 2>&1
}
full lpc16 lpc_fused16 python bench.py --config c2 --steps 1 --warmup 0 --no-cpu
full lpc "lpc_fused_kernel" python bench.py --config c3 --utts 1125 --steps 1 --warmup 0 --no-cpu
full roots lpc_roots_ python bench.py --config c3 --utts 1125 --steps 1 --warmup 0 --no-cpu
full tracker tracker_idx python bench.py --config c3 --utts 1125 --steps 1 --warmup 0 --no-cpu
full lag pitch_lag python bench.py --config c4 --utts 48 --steps 1 --warmup 0 --no-cpu
full refine pitch_refine8 python bench.py --config c4 --utts 48 --steps 1 --warmup 0 --no-cpu
full mfcc mfcc_warp python bench.py --config c5 --utts 360 --steps 1 --warmup 0 --no-cpu
ls -la gpurun_out/*_final.ncu-rep
