"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/voxbox_b200.h declares, refuses to run without a GPU (no CPU fallback), and its host-side
window tables equal the oracle's bit for bit.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
import voxbox_b200 as vb  # noqa: E402


def _declared_symbols():
    src = open(vb.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    fns = re.findall(r"VBX_API\s+(?!extern)[\w\s\*]+?\b(vbx_\w+)\s*\(", src)
    data = re.findall(r"VBX_API\s+extern\s+const\s+\w+\s+(VBX_\w+)\s*\[", src)
    return sorted(set(fns)), sorted(set(data))


def test_library_built_and_loads():
    assert os.path.exists(vb.LIB_PATH), "run `python __graft_entry__.py` (build()) first"
    lib = vb.load_library()
    assert lib.vbx_version() >= 100


def test_exports_every_declared_symbol():
    fns, data = _declared_symbols()
    assert len(fns) >= 20
    out = subprocess.check_output(["nm", "-D", "--defined-only", vb.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [s for s in fns + data if s not in exported]
    assert not missing, f"declared in include/voxbox_b200.h but not exported: {missing}"
    # and nothing undocumented leaks out of the vbx_ namespace
    extra = [s for s in exported if s.startswith("vbx_") and s not in fns]
    assert not extra, f"exported but not declared in the header: {extra}"


def test_python_binding_declares_every_function():
    fns, _ = _declared_symbols()
    lib = vb.load_library()
    undeclared = [f for f in fns if getattr(lib, f).argtypes is None and f not in ("vbx_version",)]
    assert not undeclared, f"ctypes signatures missing for: {undeclared}"


def test_rust_shim_declares_every_function():
    """rust/vox_box_b200 (source only: no Rust toolchain here) binds every entry point the header declares."""
    fns, _ = _declared_symbols()
    rs = open(os.path.join(ROOT, "rust", "vox_box_b200", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (vbx_\w+)\s*\(", rs))
    missing = [f for f in fns if f not in bound]
    assert not missing, f"missing from the Rust extern block: {missing}"
    assert not [f for f in bound if f not in fns], "Rust extern block names a function the header does not declare"


def test_formant_estimate_constants():  # lib.rs:27-28
    lib = vb.load_library()
    male = (C.c_double * 4).in_dll(lib, "VBX_MALE_FORMANT_ESTIMATES")
    female = (C.c_double * 4).in_dll(lib, "VBX_FEMALE_FORMANT_ESTIMATES")
    assert list(male) == [320., 1440., 2760., 3200.]
    assert list(female) == [480., 1760., 3200., 3520.]


def test_status_strings():  # error.rs:25-32
    lib = vb.load_library()
    assert lib.vbx_status_str(vb.ERR_LPC) == b"Denum was <= 0.0"
    assert lib.vbx_status_str(vb.ERR_WORKSPACE) == b"Not enough workspace allocated"
    assert lib.vbx_status_str(vb.ERR_POLYNOMIAL) == b"Failed to find roots"


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(vb.VoxBoxError) as e:
        vb.Context(0)
    assert e.value.status == vb.ERR_CUDA


def test_window_tables_match_oracle(oracle):
    for n in (2, 3, 16, 400, 640, 1024, 1102, 2048):
        assert np.array_equal(vb.window_table(vb.WINDOW_HANN_SYMMETRIC, n), oracle.hanning_window(n))
        assert np.array_equal(vb.window_table(vb.WINDOW_HANN_PERIODIC, n), oracle.hanning_periodic(n))
        assert np.array_equal(vb.window_table(vb.WINDOW_NONE, n), np.ones(n))


def test_product_does_not_reference_oracle():
    """The product tree must not include, link or import anything under oracle/."""
    pkg = os.path.join(ROOT, "vox_box.rs_b200")
    for dp, _, fs in os.walk(pkg):
        if "_build" in dp or "__pycache__" in dp:
            continue
        for f in fs:
            if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp", ".py", ".rs")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "vox_box_oracle" not in txt and "libvoxbox_oracle" not in txt, f
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
    out = subprocess.check_output(["ldd", vb.LIB_PATH], text=True)
    assert "oracle" not in out
