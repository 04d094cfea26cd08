#!/bin/bash
# session-2 run c: whole -m gpu suite + parity at scale with the approximate-MUFU Laguerre step
mkdir -p gpurun_out
bash tools/r2_tests.sh s2c
timeout 900 python tools/parity_scale.py > gpurun_out/s2c_parity_scale.txt 2>&1; grep -E "formants|Laguerre|pitch gpu|positional" gpurun_out/s2c_parity_scale.txt
