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


# The reference's public surface (src/*.rs of andrewcsmith/vox_box.rs; the line numbers are the reference's): every trait with
# its methods, every public struct / enum / free function on the hot path.  The shim must spell them identically.
REFERENCE_SURFACE = {
    "traits": {
        "Autocorrelate": ["autocorrelate_mut", "autocorrelate"],                          # periodic.rs:265-274
        "Pitched": ["pitch"],                                                             # periodic.rs:356-358
        "LagType": [],                                                                    # periodic.rs:232-234
        "LPC": ["lpc_mut", "lpc", "lpc_praat_mut", "lpc_praat"],                          # spectrum.rs:50-55
        "ToResonance": ["to_resonance"],                                                  # spectrum.rs:195-197
        "EstimateFormants": ["estimate_formants"],                                        # spectrum.rs:216-219
        "MFCC": ["mfcc"],                                                                 # spectrum.rs:371-373
        "Polynomial": ["degree", "off_low", "laguerre", "find_roots_work_size", "find_roots", "find_roots_mut",
                       "div_polynomial", "div_polynomial_mut"],                            # polynomial.rs:10-21
        "RMS": ["rms"], "Amplitude": ["amplitude"], "MaxAmplitude": ["max_amplitude"],    # waves.rs:10-41
        "Normalize": ["normalize_with_max", "normalize"], "Filter": ["preemphasis"],      # waves.rs:61-84
    },
    "types": ["Pitch", "PitchExtractor", "Interpolation", "HanningLag", "LPCSolver", "Resonance", "FormantExtractor",
              "VoxBoxError", "VoxBoxResult"],
    "functions": ["interpolate_sinc", "improve_extremum", "hz_to_mel", "mel_to_hz", "dct", "dct_mut", "find_formants",
                  "find_formants_real_work_size", "find_formants_complex_work_size", "from_root", "with_context"],
    "constants": ["MAX_RESONANCES", "MALE_FORMANT_ESTIMATES", "FEMALE_FORMANT_ESTIMATES"],
    "signatures": [  # spelled as in the reference (whitespace-insensitive)
        "fn autocorrelate_mut(&self, coeffs: &mut [T]);",
        "fn lpc_mut(&self, n_coeffs: usize, ac: &mut [T], kc: &mut [T], tmp: &mut [T]);",
        "fn lpc_praat_mut(&self, n_coeffs: usize, coeffs: &mut [T], work: &mut [T]) -> VoxBoxResult<()>;",
        "fn lpc_praat(&self, n_coeffs: usize) -> VoxBoxResult<Vec<T>>;",
        "fn pitch<W: LagType>(&self, sample_rate: T, threshold: T, local_peak: S, global_peak: S, min: T, max: T) -> Vec<Pitch<T>>;",
        "fn to_resonance(&self, sample_rate: T) -> Vec<Resonance<T>>;",
        "fn estimate_formants(&mut self, resonances: &[Resonance<T>]);",
        "fn mfcc(&self, num_coeffs: usize, freq_bounds: (f64, f64), sample_rate: f64) -> Vec<T>;",
        "fn laguerre(&self, z: Complex<T>) -> Complex<T>;",
        "fn find_roots(&self) -> VoxBoxResult<Vec<Complex<T>>>;",
        "fn div_polynomial(&mut self, other: Complex<T>) -> VoxBoxResult<Vec<Complex<T>>>;",
        "fn div_polynomial_mut(&'a mut self, other: Complex<T>, rem: &'a mut [Complex<T>]) -> VoxBoxResult<()>;",
        "fn normalize_with_max(&mut self, max: Option<S>);",
        "fn preemphasis(&mut self, factor: f64) -> &mut Self;",
        "pub fn new(num_formants: usize, resonances: I, starting_estimates: Vec<Resonance<T>>) -> Self",
        "pub fn new(candidates: &'a [&'a [Pitch<T>]], voiced_unvoiced_cost: T, voicing_threshold: T) -> Self",
        "pub fn interpolate_sinc<S: Elem>(y: &[S], offset: isize, nx: usize, x: S, max_depth: usize) -> f64",
        "pub fn improve_extremum<S: Elem>(y: &[S], offset: isize, nx: usize, ixmid: f64, interp: Interpolation, is_max: bool) -> (f64, f64)",
    ],
}


def test_rust_shim_mirrors_the_reference_trait_surface():
    """north_star: "the existing Rust trait/function surface stays unchanged".  The shim (source only: no Rust toolchain here)
    defines every trait of the reference with the same method names and signatures, implemented for slices / VecDeque, plus the
    public types, free functions and constants of the path."""
    rs = open(os.path.join(ROOT, "rust", "vox_box_b200", "src", "lib.rs")).read()
    flat = re.sub(r"\s+", " ", rs)
    for trait, methods in REFERENCE_SURFACE["traits"].items():
        assert re.search(rf"pub trait {trait}\b", rs), f"trait {trait} missing"
        m = re.search(rf"pub trait {trait}\b[^{{]*\{{(.*?)\n    \}}", rs, flags=re.S)
        body = m.group(1) if m else ""
        for meth in methods:
            assert re.search(rf"fn {meth}\b", body), f"{trait}::{meth} missing from the trait definition"
        if trait != "LagType":
            assert re.search(rf"impl<[^>]*>\s+{trait}(<[^>]*>)?\s+for\s+(\[|VecDeque|S\b)", rs), f"no slice impl of {trait}"
    assert re.search(r"impl<T: Elem> Autocorrelate<T> for VecDeque<T>", rs)   # periodic.rs:291-304
    for t in REFERENCE_SURFACE["types"]:
        assert re.search(rf"pub (struct|enum|type) {t}\b", rs), f"type {t} missing"
    for f in REFERENCE_SURFACE["functions"]:
        assert re.search(rf"pub fn {f}\b", rs), f"function {f} missing"
    for c in REFERENCE_SURFACE["constants"]:
        assert re.search(rf"pub const {c}\b", rs), f"constant {c} missing"
    for sig in REFERENCE_SURFACE["signatures"]:
        assert re.sub(r"\s+", " ", sig) in flat, f"signature not found verbatim: {sig}"
    for mod in ("periodic", "spectrum", "polynomial", "waves", "error"):
        assert re.search(rf"pub mod {mod}\b", rs), f"module {mod} missing"
    # braces balance (a cheap stand-in for the compiler this image lacks)
    code = re.sub(r"//[^\n]*", "", rs)
    code = re.sub(r'"(\\.|[^"\\])*"', '""', code)
    code = re.sub(r"'(\\.|[^'\\])'", "' '", code)
    for a, b in ("{}", "()", "[]"):
        assert code.count(a) == code.count(b), f"unbalanced {a}{b}: {code.count(a)} vs {code.count(b)}"


def test_multi_partition_rule():
    """vbx_multi_partition (host arithmetic, no GPU): contiguous, exhaustive, balanced to within one unit."""
    for n, parts in ((4500, 8), (10, 3), (7, 8), (0, 4), (360000, 8)):
        lo_hi = [vb.multi_partition(n, parts, p) for p in range(parts)]
        assert lo_hi[0][0] == 0 and lo_hi[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(lo_hi, lo_hi[1:]))
        sizes = [h - l for l, h in lo_hi]
        assert max(sizes) - min(sizes) <= 1


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
