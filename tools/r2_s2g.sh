#!/bin/bash
# session-2 run g: default bench line (all configs) + reference arm + smoke, as the driver runs them
mkdir -p gpurun_out
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > gpurun_out/s2g_bench.json 2> gpurun_out/s2g_bench.err
echo "bench rc=$?"; tail -4 gpurun_out/s2g_bench.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/s2g_ref.json 2> gpurun_out/s2g_ref.err
echo "ref rc=$?"; tail -4 gpurun_out/s2g_ref.err
python __graft_entry__.py smoke > gpurun_out/s2g_smoke.txt 2>&1; tail -2 gpurun_out/s2g_smoke.txt
python tools/show_bench.py gpurun_out/s2g_bench.json 2>/dev/null | head -60
