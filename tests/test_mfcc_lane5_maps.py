"""CPU check of mfcc_lane5_kernel's index maps (vbx_mfcc_lane5.cuh) through their numpy emulation (tools/mfcc_lane5_emulation.py):
the in-place radix 8-5-5 decimation-in-frequency passes over five lanes, the mirrored columns of blocks 5..7, the pairing of
butterfly (kA, kB) with (8 - kA, 4 - kB) for the fused untangle and the kA-major spectrum layout reproduce numpy's rfft, every
spectrum slot is written, and only the self-paired bin 100 twice.  The kernel itself is compared with the oracle in
tests/test_gpu_mfcc_waves.py (-m gpu)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import mfcc_lane5_emulation as emu  # noqa: E402


def test_lane5_transform_matches_rfft():
    for seed in (1, 2):
        err, every, twice = emu.check(seed)
        assert err < 1e-12 and every and twice == [emu.slot(100)]


def test_lane5_slot_map_is_a_bijection():
    slots = [emu.slot(k) for k in range(emu.MC + 1)]
    assert sorted(slots) == list(range(emu.MC + 1))
    # consecutive bins are 25 slots (one 16-byte bank group modulo 8) apart except where k mod 8 wraps
    assert all((emu.slot(k + 1) - emu.slot(k)) % 8 == 1 for k in range(emu.MC - 1) if k % 8 != 7)


def test_lane5_frame_stride_keeps_unit_stride_accesses_conflict_free():
    assert emu.wavefronts(205, lambda s: s) == 4          # 30 lanes, 16 bytes each: four quarter-warp wavefronts is the minimum
    assert emu.wavefronts(205, lambda s: 100 - s) == 8    # (the mirrored direction would pay double: hence the mirrored columns)
