#!/bin/bash
# C3 device-only timing for several numbers of SMs left to the co-running tracker
for r in ${RS:-9 7 6 5}; do
  echo "== VBX_FORMANT_RESERVE_SMS=$r"; VBX_FORMANT_RESERVE_SMS=$r timeout 300 python bench.py --config c3 --device-only --steps 10 --warmup 3 2>&1 | cut -c1-420
done
