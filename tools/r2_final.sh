#!/bin/bash
# round 2, final single-GPU run: suite, default bench + reference arm + smoke, launch lists, parity at scale, lag64 capture
mkdir -p gpurun_out
bash tools/r2_tests.sh r2f
( time timeout 1200 python bench.py ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2f_bench.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"
python __graft_entry__.py smoke > gpurun_out/r2f_smoke.txt 2>&1; tail -1 gpurun_out/r2f_smoke.txt | cut -c1-200
for c in c2 c3 c4 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_${c}.csv python bench.py --config $c --steps 2 --warmup 3 --device-only > gpurun_out/r2f_ncu_bench_${c}.log 2>&1
done
timeout 900 python tools/parity_scale.py > gpurun_out/r2f_parity_scale.txt 2>&1; grep -E "mismatch|positional|mfcc|lpc:" gpurun_out/r2f_parity_scale.txt | cut -c1-250
ncu --set full --clock-control none --import-source on -k regex:pitch_lag64 -c 1 -f -o gpurun_out/prof_lag64_v2 python bench.py --config c4 --utts 48 --steps 1 --warmup 0 --no-cpu --device-only > gpurun_out/ncu_lag64_v2.log 2>&1
bash tools/ncu_summary.sh gpurun_out/prof_lag64_v2.ncu-rep gpurun_out/lag64_v2_full.txt | tail -1
python tools/show_bench.py gpurun_out/r2f_bench.json 2>/dev/null | head -12 | cut -c1-260
