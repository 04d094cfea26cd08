"""ctypes binding of libvoxbox_b200.so (the C ABI in include/voxbox_b200.h).

Thin by design: the product is the CUDA library; this module only marshals
pointers for the test-suite, bench.py and Python callers.  There is no CPU
fallback — importing works anywhere (so symbol/export checks can run without a
GPU), but creating a `Context` fails loudly when the library or a CUDA device is
missing.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_PKG))  # vox_box.rs_b200/
LIB_PATH = os.path.join(ROOT, "libvoxbox_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(ROOT), "include", "voxbox_b200.h")

OK, ERR_LPC, ERR_PITCH, ERR_POLYNOMIAL, ERR_WORKSPACE, ERR_CUDA, ERR_BADARG, ERR_NOMEM = range(8)
F32, F64, I16 = 0, 1, 2
WINDOW_NONE, WINDOW_HANN_SYMMETRIC, WINDOW_HANN_PERIODIC = 0, 1, 2
MAX_RESONANCES = 32
_NP = {F32: np.float32, F64: np.float64, I16: np.int16}


class VoxBoxError(RuntimeError):
    """Mirror of error.rs VoxBoxError + the CUDA/BADARG/NOMEM codes of the C ABI."""

    def __init__(self, status, message):
        super().__init__(f"vbx status {status}: {message}")
        self.status = status


class Frames(C.Structure):
    """struct vbx_frames"""
    _fields_ = [("base", C.c_void_p), ("n_frames", C.c_int64), ("frame_stride", C.c_int64),
                ("frames_per_segment", C.c_int64), ("segment_stride", C.c_int64), ("frame_len", C.c_int32), ("dtype", C.c_int32), ("window", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def load_library():
    """Load libvoxbox_b200.so; raises if it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VoxBoxError(ERR_CUDA, f"{LIB_PATH} not built — run `python __graft_entry__.py` "
                                        "(make -C vox_box.rs_b200); there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


_vp, _i32, _i64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
_frp = C.POINTER(Frames)


def _declare(L):
    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("vbx_ctx_create", C.c_int, C.c_int, C.POINTER(_vp))
    sig("vbx_ctx_destroy", C.c_int, _vp)
    sig("vbx_sync", C.c_int, _vp)
    sig("vbx_ctx_stream", _vp, _vp)
    sig("vbx_last_error", C.c_char_p, _vp)
    sig("vbx_status_str", C.c_char_p, C.c_int)
    sig("vbx_version", C.c_int)
    sig("vbx_device_sm_count", C.c_int, _vp)
    sig("vbx_kernel_launches", _i64, _vp)
    sig("vbx_malloc", C.c_int, _vp, _sz, C.POINTER(_vp))
    sig("vbx_free", C.c_int, _vp, _vp)
    sig("vbx_malloc_host", C.c_int, _vp, _sz, C.POINTER(_vp))
    sig("vbx_free_host", C.c_int, _vp, _vp)
    sig("vbx_memcpy_h2d", C.c_int, _vp, _vp, _vp, _sz)
    sig("vbx_memcpy_d2h", C.c_int, _vp, _vp, _vp, _sz)
    sig("vbx_memset", C.c_int, _vp, _vp, C.c_int, _sz)
    sig("vbx_timer_start", C.c_int, _vp)
    sig("vbx_timer_stop_ms", C.c_int, _vp, C.POINTER(C.c_float))
    sig("vbx_measure_peaks", C.c_int, _vp, C.POINTER(C.c_double), C.POINTER(C.c_double))
    sig("vbx_window_table_host", C.c_int, C.c_int, _i32, C.POINTER(C.c_double))
    sig("vbx_autocorrelate", C.c_int, _vp, _frp, _i32, _vp, _i32)
    sig("vbx_autocorrelate_host", C.c_int, _vp, _frp, _i32, _vp, _i32)
    sig("vbx_lpc_levinson", C.c_int, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _i32)
    sig("vbx_lpc", C.c_int, _vp, _frp, _i32, _vp, _vp, _vp, _i32)
    sig("vbx_lpc_host", C.c_int, _vp, _frp, _i32, _vp, _vp, _vp, _i32)


def window_table(window, n):
    out = np.empty(n, dtype=np.float64)
    st = load_library().vbx_window_table_host(window, n, out.ctypes.data_as(C.POINTER(C.c_double)))
    if st:
        raise VoxBoxError(st, "vbx_window_table_host")
    return out


class DeviceArray:
    """Caller-owned HBM buffer (vbx_malloc) with numpy shape/dtype metadata."""

    def __init__(self, ctx, shape, dtype):
        self.ctx, self.shape, self.dtype = ctx, tuple(int(s) for s in np.atleast_1d(shape)), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = _vp()
        ctx._check(ctx.lib.vbx_malloc(ctx.h, max(self.nbytes, 1), C.byref(p)), "vbx_malloc")
        self.ptr = p.value

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.nbytes == self.nbytes
        self.ctx._check(self.ctx.lib.vbx_memcpy_h2d(self.ctx.h, self.ptr, host.ctypes.data, self.nbytes), "h2d")
        self.ctx.sync()  # `host` may be a temporary
        return self

    def to_host(self):
        out = np.empty(self.shape, dtype=self.dtype)
        self.ctx._check(self.ctx.lib.vbx_memcpy_d2h(self.ctx.h, out.ctypes.data, self.ptr, self.nbytes), "d2h")
        self.ctx.sync()
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.vbx_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """vbx_ctx: one device + stream + scratch arena.  Fails loudly without the library / a GPU."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = _vp()
        st = self.lib.vbx_ctx_create(device, C.byref(h))
        if st != OK:
            raise VoxBoxError(st, "vbx_ctx_create failed: no usable CUDA device (there is no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.vbx_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st, what):
        if st != OK:
            raise VoxBoxError(st, f"{what}: {self.lib.vbx_last_error(self.h).decode()} "
                                  f"[{self.lib.vbx_status_str(st).decode()}]")

    # -- plumbing ---------------------------------------------------------------------------
    def sync(self):
        self._check(self.lib.vbx_sync(self.h), "vbx_sync")

    def empty(self, shape, dtype):
        return DeviceArray(self, shape, dtype)

    def to_device(self, host):
        host = np.ascontiguousarray(host)
        return DeviceArray(self, host.shape, host.dtype).upload(host)

    def timer_start(self):
        self._check(self.lib.vbx_timer_start(self.h), "vbx_timer_start")

    def timer_stop_ms(self):
        ms = C.c_float(0)
        self._check(self.lib.vbx_timer_stop_ms(self.h, C.byref(ms)), "vbx_timer_stop_ms")
        return ms.value

    @property
    def kernel_launches(self):
        return self.lib.vbx_kernel_launches(self.h)

    @property
    def sm_count(self):
        return self.lib.vbx_device_sm_count(self.h)

    def measure_peaks(self):
        a, b = C.c_double(0), C.c_double(0)
        self._check(self.lib.vbx_measure_peaks(self.h, C.byref(a), C.byref(b)), "vbx_measure_peaks")
        return dict(fp32_tflops=a.value, fp64_tflops=b.value)

    @staticmethod
    def frames(base_ptr, n_frames, frame_len, frame_stride, window=WINDOW_NONE, dtype=F32, frames_per_segment=0,
               segment_stride=0):
        """vbx_frames: frame f = u*J + j starts at base[u*segment_stride + j*frame_stride] (J = frames_per_segment)."""
        return Frames(base_ptr, n_frames, frame_stride, frames_per_segment, segment_stride, frame_len, dtype, window, 0)

    @staticmethod
    def n_frames_of(n_samples, frame_len, hop):
        """Windower::{hanning,rectangle}: a frame while bin <= remaining, advance by hop."""
        return 0 if n_samples < frame_len else (n_samples - frame_len) // hop + 1

    # -- device-pointer ops -------------------------------------------------------------------
    def autocorrelate(self, frames, n_lags, out_dtype=F64, out=None):
        r = out if out is not None else self.empty((frames.n_frames, n_lags), _NP[out_dtype])
        self._check(self.lib.vbx_autocorrelate(self.h, C.byref(frames), n_lags, r.ptr, out_dtype), "vbx_autocorrelate")
        return r

    def lpc(self, frames, p, out_dtype=F64, want_r=True, want_kc=True, out=None):
        """Fused window→autocorrelate(p+1)→lpc(p).  Returns (r, ac, kc) device arrays (None if not wanted)."""
        F = frames.n_frames
        if out is not None:
            r, ac, kc = out
        else:
            r = self.empty((F, p + 1), _NP[out_dtype]) if want_r else None
            ac = self.empty((F, p + 1), _NP[out_dtype])
            kc = self.empty((F, p), _NP[out_dtype]) if want_kc else None
        self._check(self.lib.vbx_lpc(self.h, C.byref(frames), p, r.ptr if r else None, ac.ptr if ac else None,
                                     kc.ptr if kc else None, out_dtype), "vbx_lpc")
        return r, ac, kc

    def lpc_levinson(self, r, p, out_dtype=F64):
        F, stride = r.shape
        ac = self.empty((F, p + 1), _NP[out_dtype])
        kc = self.empty((F, p), _NP[out_dtype])
        rd = F64 if r.dtype == np.float64 else F32
        self._check(self.lib.vbx_lpc_levinson(self.h, r.ptr, rd, F, stride, p, ac.ptr, kc.ptr, out_dtype),
                    "vbx_lpc_levinson")
        return ac, kc

    # -- host-pointer twins (H2D + kernels + D2H inside the call) ------------------------------
    def autocorrelate_host(self, audio, n_frames, frame_len, stride, window, n_lags, out_dtype=F64):
        audio = np.ascontiguousarray(audio)
        fr = self.frames(audio.ctypes.data, n_frames, frame_len, stride, window, I16 if audio.dtype == np.int16 else F32)
        r = np.empty((n_frames, n_lags), dtype=_NP[out_dtype])
        self._check(self.lib.vbx_autocorrelate_host(self.h, C.byref(fr), n_lags, r.ctypes.data, out_dtype),
                    "vbx_autocorrelate_host")
        return r

    def lpc_host(self, audio, n_frames, frame_len, stride, window, p, out_dtype=F64, out=None):
        audio = np.ascontiguousarray(audio)
        fr = self.frames(audio.ctypes.data, n_frames, frame_len, stride, window, I16 if audio.dtype == np.int16 else F32)
        if out is None:
            out = (np.empty((n_frames, p + 1), dtype=_NP[out_dtype]), np.empty((n_frames, p + 1), dtype=_NP[out_dtype]),
                   np.empty((n_frames, p), dtype=_NP[out_dtype]))
        r, ac, kc = out
        self._check(self.lib.vbx_lpc_host(self.h, C.byref(fr), p, r.ctypes.data if r is not None else None,
                                          ac.ctypes.data if ac is not None else None,
                                          kc.ctypes.data if kc is not None else None, out_dtype), "vbx_lpc_host")
        return r, ac, kc
