#!/bin/bash
# racecheck of the whole -m gpu suite with the mbarrier / TMA persistent LPC kernel switched off (VBX_LPCP=0): racecheck does not
# model mbarrier-ordered async-proxy writes and reports that kernel's hand-offs as hazards (see tests/test_gpu_round2.py::
# test_lpc_persistent_kernel_repeatable_under_load for how that kernel is checked instead)
mkdir -p gpurun_out
VBX_LPCP=0 timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -p no:cacheprovider -k "not persistent" > gpurun_out/r2f_racecheck_raw.txt 2>&1
{
  echo "# VBX_LPCP=0 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -k 'not persistent'"
  grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r2f_racecheck_raw.txt
  echo "# distinct hazard sites:"
  grep -oE "(Read|Write) access at .* in [a-z_0-9]+\.(cu|cuh):[0-9]+" gpurun_out/r2f_racecheck_raw.txt | sed -E 's/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -20
} > gpurun_out/r2f_racecheck.txt
cat gpurun_out/r2f_racecheck.txt | cut -c1-250
