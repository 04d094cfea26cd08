"""Helpers shared by the -m gpu parity tests (all calls go through the C ABI via ctypes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
import voxbox_b200 as vb  # noqa: E402
from voxbox_b200 import synth  # noqa: E402

_ctx = None


def ctx():
    """One context for the test session; raises (fails the test loudly) without a GPU."""
    global _ctx
    if _ctx is None:
        _ctx = vb.Context(0)
    return _ctx


def normwise(gpu, ref):
    """max|gpu − ref| / max|ref| per frame vector (SURVEY §8d tolerance definition)."""
    gpu, ref = np.asarray(gpu, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    num = np.max(np.abs(gpu - ref), axis=-1)
    den = np.maximum(np.max(np.abs(ref), axis=-1), 1e-300)
    return num / den
