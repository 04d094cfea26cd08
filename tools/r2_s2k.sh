#!/bin/bash
# C3 device-only timing for several chunk counts of vbx_find_formants (tapered chunks)
for k in ${KS:-6 8 10 12}; do
  echo "== VBX_FORMANT_CHUNKS=$k"; VBX_FORMANT_CHUNKS=$k timeout 300 python bench.py --config c3 --device-only --steps 10 --warmup 3 2>&1 | cut -c1-160
done
