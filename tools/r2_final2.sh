#!/bin/bash
# round 2, final run part 2: the bare `python bench.py` / `python bench.py --impl reference` (no flags), the suite once more
mkdir -p gpurun_out
bash tools/r2_tests.sh r2f2
( time timeout 1200 python bench.py ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2f_bench.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/r2f_ref.err
python tools/show_bench.py gpurun_out/r2f_bench.json 2>/dev/null | head -12 | cut -c1-260
