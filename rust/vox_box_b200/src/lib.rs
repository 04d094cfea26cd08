//! vox_box_b200 — vox_box's trait/function surface on top of libvoxbox_b200 (C ABI, sm_100a CUDA).
//!
//! NOT COMPILED IN THE BUILD IMAGE (no cargo/rustc there): source only, see README.md.
//!
//! Layout:
//!  * `ffi`     — the `extern "C"` block, one declaration per entry point of include/voxbox_b200.h;
//!  * `Context` — RAII wrapper of `vbx_ctx` (one device + stream + scratch arena, not `Sync`);
//!  * slice traits with the reference's names and signatures (a call == a batch of one frame,
//!    host pointers, through the `_host` twins where they exist);
//!  * `Batch`   — device-resident frame tensors for real workloads (thousands of frames per call).
//!
//! Error mapping: `vbx_status` → `VoxBoxError` exactly as error.rs:6-16, plus `Cuda`/`BadArg`/`NoMem`.
#![allow(non_camel_case_types)]

extern crate num_complex;

use num_complex::Complex;
use std::os::raw::{c_char, c_int, c_void};
use std::ptr;

pub mod ffi {
    use super::*;

    #[repr(C)]
    pub struct vbx_ctx {
        _private: [u8; 0],
    }

    /// struct vbx_frames (include/voxbox_b200.h)
    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct vbx_frames {
        pub base: *const c_void,
        pub n_frames: i64,
        pub frame_stride: i64,
        pub frames_per_segment: i64,
        pub segment_stride: i64,
        pub frame_len: i32,
        pub dtype: i32,
        pub window: i32,
        pub reserved: i32,
    }

    pub const VBX_OK: c_int = 0;
    pub const VBX_ERR_LPC: c_int = 1;
    pub const VBX_ERR_PITCH: c_int = 2;
    pub const VBX_ERR_POLYNOMIAL: c_int = 3;
    pub const VBX_ERR_WORKSPACE: c_int = 4;
    pub const VBX_ERR_CUDA: c_int = 5;
    pub const VBX_ERR_BADARG: c_int = 6;
    pub const VBX_ERR_NOMEM: c_int = 7;
    pub const VBX_F32: i32 = 0;
    pub const VBX_F64: i32 = 1;
    pub const VBX_I16: i32 = 2;
    pub const VBX_WINDOW_NONE: i32 = 0;
    pub const VBX_WINDOW_HANN_SYMMETRIC: i32 = 1;
    pub const VBX_WINDOW_HANN_PERIODIC: i32 = 2;
    pub const VBX_LPC_BURG: i32 = 0;
    pub const VBX_LPC_AUTOCORR: i32 = 1;

    #[link(name = "voxbox_b200")]
    extern "C" {
        pub fn vbx_ctx_create(device: c_int, out: *mut *mut vbx_ctx) -> c_int;
        pub fn vbx_ctx_destroy(ctx: *mut vbx_ctx) -> c_int;
        pub fn vbx_sync(ctx: *mut vbx_ctx) -> c_int;
        pub fn vbx_last_error(ctx: *mut vbx_ctx) -> *const c_char;
        pub fn vbx_status_str(status: c_int) -> *const c_char;
        pub fn vbx_malloc(ctx: *mut vbx_ctx, bytes: usize, dev_out: *mut *mut c_void) -> c_int;
        pub fn vbx_free(ctx: *mut vbx_ctx, dev: *mut c_void) -> c_int;
        pub fn vbx_memcpy_h2d(ctx: *mut vbx_ctx, dev: *mut c_void, host: *const c_void, bytes: usize) -> c_int;
        pub fn vbx_memcpy_d2h(ctx: *mut vbx_ctx, host: *mut c_void, dev: *const c_void, bytes: usize) -> c_int;
        pub fn vbx_memcpy_d2d(ctx: *mut vbx_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
        pub fn vbx_mfcc_set_fft_precision(ctx: *mut vbx_ctx, dtype: i32) -> c_int;
        pub fn vbx_profile_begin(ctx: *mut vbx_ctx) -> c_int;
        pub fn vbx_profile_end(ctx: *mut vbx_ctx) -> c_int;
        pub fn vbx_profile_count(ctx: *mut vbx_ctx) -> c_int;
        pub fn vbx_profile_entry(ctx: *mut vbx_ctx, index: c_int, name_out: *mut c_char, name_len: c_int, ms_total: *mut f64,
                                 launches: *mut i64) -> c_int;

        pub fn vbx_autocorrelate(ctx: *mut vbx_ctx, frames: *const vbx_frames, n_lags: i32, r_out: *mut c_void, out_dtype: i32) -> c_int;
        pub fn vbx_autocorrelate_host(ctx: *mut vbx_ctx, frames: *const vbx_frames, n_lags: i32, r_out: *mut c_void, out_dtype: i32) -> c_int;
        pub fn vbx_lpc_levinson(ctx: *mut vbx_ctx, r: *const c_void, r_dtype: i32, n_frames: i64, r_stride: i32, p: i32,
                                ac_out: *mut c_void, kc_out: *mut c_void, out_dtype: i32) -> c_int;
        pub fn vbx_lpc(ctx: *mut vbx_ctx, frames: *const vbx_frames, p: i32, r_out: *mut c_void, ac_out: *mut c_void,
                       kc_out: *mut c_void, out_dtype: i32) -> c_int;
        pub fn vbx_lpc_host(ctx: *mut vbx_ctx, frames: *const vbx_frames, p: i32, r_out: *mut c_void, ac_out: *mut c_void,
                            kc_out: *mut c_void, out_dtype: i32) -> c_int;
        pub fn vbx_lpc_burg(ctx: *mut vbx_ctx, frames: *const vbx_frames, p: i32, coeffs_out: *mut c_void, status_out: *mut u8,
                            out_dtype: i32) -> c_int;
        pub fn vbx_find_roots(ctx: *mut vbx_ctx, coeffs: *const c_void, dtype: i32, n_polys: i64, len: i32, roots_out: *mut c_void,
                              status_out: *mut u8) -> c_int;
        pub fn vbx_find_roots_work_size(len: i64) -> i64;
        pub fn vbx_laguerre(ctx: *mut vbx_ctx, coeffs: *const c_void, dtype: i32, n_polys: i64, len: i32, start_re: f64,
                            start_im: f64, z_out: *mut c_void) -> c_int;
        pub fn vbx_div_polynomial(ctx: *mut vbx_ctx, coeffs_inout: *mut c_void, dtype: i32, n_polys: i64, len: i32,
                                  other: *const c_void, other_per_poly: i32, rem_out: *mut c_void, status_out: *mut u8) -> c_int;
        pub fn vbx_roots_to_resonances(ctx: *mut vbx_ctx, roots: *const c_void, dtype: i32, n_frames: i64, n_roots: i32,
                                       sample_rate: f64, strict_im: i32, res_out: *mut c_void, res_slots: i32,
                                       nres_out: *mut i32, out_dtype: i32) -> c_int;
        pub fn vbx_estimate_formants(ctx: *mut vbx_ctx, resonances: *const c_void, res_dtype: i32, res_slots: i32,
                                     n_resonances: i32, n_segments: i64, frames_per_segment: i64, status_in: *const u8,
                                     est_inout: *mut c_void, n_estimates: i32, tracks_out: *mut c_void, dtype: i32) -> c_int;
        pub fn vbx_find_formants_real_work_size(buf_len: i64, n_coeffs: i64) -> i64;
        pub fn vbx_find_formants_complex_work_size(n_coeffs: i64) -> i64;
        pub fn vbx_find_formants(ctx: *mut vbx_ctx, frames: *const vbx_frames, sample_rate: f64, n_coeffs: i32, lpc_method: i32,
                                 est_inout: *mut c_void, n_formants: i32, tracks_out: *mut c_void, resonances_out: *mut c_void,
                                 nres_out: *mut i32, status_out: *mut u8, dtype: i32) -> c_int;
        pub fn vbx_find_formants_host(ctx: *mut vbx_ctx, frames: *const vbx_frames, sample_rate: f64, n_coeffs: i32,
                                      lpc_method: i32, est_inout: *mut c_void, n_formants: i32, tracks_out: *mut c_void,
                                      resonances_out: *mut c_void, nres_out: *mut i32, status_out: *mut u8, dtype: i32) -> c_int;
        pub fn vbx_find_formants_resampled(ctx: *mut vbx_ctx, frames: *const vbx_frames, sample_rate: f64, resample_ratio: f64,
                                           n_coeffs: i32, est_inout: *mut c_void, n_formants: i32, tracks_out: *mut c_void,
                                           resonances_out: *mut c_void, nres_out: *mut i32, status_out: *mut u8, dtype: i32) -> c_int;
        pub fn vbx_pitch(ctx: *mut vbx_ctx, frames: *const vbx_frames, sample_rate: f64, threshold: f64, min_hz: f64, max_hz: f64,
                         max_candidates: i32, cand_out: *mut c_void, n_cand_out: *mut i32, status_out: *mut u8, out_dtype: i32) -> c_int;
        pub fn vbx_pitch_host(ctx: *mut vbx_ctx, frames: *const vbx_frames, sample_rate: f64, threshold: f64, min_hz: f64,
                              max_hz: f64, max_candidates: i32, cand_out: *mut c_void, n_cand_out: *mut i32, status_out: *mut u8,
                              out_dtype: i32) -> c_int;
        pub fn vbx_pitch_extract(ctx: *mut vbx_ctx, cand: *const c_void, dtype: i32, n_frames: i64, max_candidates: i32,
                                 out: *mut c_void) -> c_int;
        pub fn vbx_pitch_viterbi(ctx: *mut vbx_ctx, cand: *const c_void, dtype: i32, n_cand: *const i32, n_segments: i64,
                                 frames_per_segment: i64, max_candidates: i32, voiced_unvoiced_cost: f64, octave_jump_cost: f64,
                                 octave_cost: f64, ceiling_hz: f64, path_out: *mut c_void, index_out: *mut i32) -> c_int;
        pub fn vbx_autocorrelate_ring(ctx: *mut vbx_ctx, rings: *const c_void, dtype: i32, n_rings: i64, capacity: i64,
                                      heads: *const i64, n: i32, n_lags: i32, r_out: *mut c_void, out_dtype: i32) -> c_int;
        pub fn vbx_lpc_to_resonances(ctx: *mut vbx_ctx, lpc: *const c_void, lpc_dtype: i32, n_frames: i64, lpc_stride: i32, p: i32,
                                     lpc_has_leading_one: i32, sample_rate: f64, strict_im: i32, status_in: *const u8,
                                     res_out: *mut c_void, res_slots: i32, nres_out: *mut i32, roots_out: *mut c_void,
                                     status_out: *mut u8, out_dtype: i32, precision: i32) -> c_int;
        pub fn vbx_window_table_host(window: c_int, n: i32, out: *mut f64) -> c_int;
        pub fn vbx_malloc_host(ctx: *mut vbx_ctx, bytes: usize, host_out: *mut *mut c_void) -> c_int;
        pub fn vbx_free_host(ctx: *mut vbx_ctx, host: *mut c_void) -> c_int;
        pub fn vbx_memset(ctx: *mut vbx_ctx, dev: *mut c_void, value: c_int, bytes: usize) -> c_int;
        pub fn vbx_ctx_stream(ctx: *mut vbx_ctx) -> *mut c_void;
        pub fn vbx_version() -> c_int;
        pub fn vbx_device_sm_count(ctx: *mut vbx_ctx) -> c_int;
        pub fn vbx_kernel_launches(ctx: *mut vbx_ctx) -> i64;
        pub fn vbx_timer_start(ctx: *mut vbx_ctx) -> c_int;
        pub fn vbx_timer_stop_ms(ctx: *mut vbx_ctx, ms_out: *mut f32) -> c_int;
        pub fn vbx_measure_peaks(ctx: *mut vbx_ctx, fp32_tflops: *mut f64, fp64_tflops: *mut f64) -> c_int;
        pub fn vbx_interpolate_sinc(ctx: *mut vbx_ctx, y: *const f64, n_series: i64, y_len: i64, offset: i64, nx: i64,
                                    x: *const f64, n_points: i64, max_depth: i64, out: *mut f64) -> c_int;
        pub fn vbx_improve_extremum(ctx: *mut vbx_ctx, y: *const f64, n_series: i64, y_len: i64, offset: i64, nx: i64,
                                    ixmid: *const f64, n_points: i64, interpolation: i32, sinc_depth: i64, is_max: i32,
                                    xmid_out: *mut f64, ymid_out: *mut f64) -> c_int;
        pub fn vbx_mfcc(ctx: *mut vbx_ctx, frames: *const vbx_frames, num_coeffs: i32, n_keep: i32, freq_lo: f64, freq_hi: f64,
                        sample_rate: f64, out: *mut c_void, energies_out: *mut c_void, out_dtype: i32) -> c_int;
        pub fn vbx_mfcc_host(ctx: *mut vbx_ctx, frames: *const vbx_frames, num_coeffs: i32, n_keep: i32, freq_lo: f64,
                             freq_hi: f64, sample_rate: f64, out: *mut c_void, out_dtype: i32) -> c_int;
        pub fn vbx_hz_to_mel(hz: f64) -> f64;
        pub fn vbx_mel_to_hz(mel: f64) -> f64;
        pub fn vbx_dct(ctx: *mut vbx_ctx, signal: *const c_void, dtype: i32, n_signals: i64, n: i32, coeffs: *mut c_void) -> c_int;
        pub fn vbx_rms(ctx: *mut vbx_ctx, x: *const c_void, dtype: i32, n_signals: i64, n: i32, stride: i64, out: *mut c_void) -> c_int;
        pub fn vbx_max_amplitude(ctx: *mut vbx_ctx, x: *const c_void, dtype: i32, n_signals: i64, n: i32, stride: i64,
                                 out: *mut c_void) -> c_int;
        pub fn vbx_normalize(ctx: *mut vbx_ctx, x_inout: *mut c_void, dtype: i32, n_signals: i64, n: i32, stride: i64,
                             maxes: *const c_void) -> c_int;
        pub fn vbx_preemphasis(ctx: *mut vbx_ctx, x_inout: *mut c_void, dtype: i32, n_signals: i64, n: i32, stride: i64,
                               factor: f64) -> c_int;
    }
}

// ---- error.rs:4-38 -------------------------------------------------------------------------
pub type VoxBoxResult<T> = Result<T, VoxBoxError>;

#[derive(Debug)]
pub enum VoxBoxError {
    LPC(&'static str),
    Pitch(&'static str),
    Polynomial(&'static str),
    Workspace,
    /// no reference analogue: the device/library failed (there is no CPU fallback)
    Cuda(String),
    BadArg(String),
    NoMem,
}

pub const MAX_RESONANCES: usize = 32; // lib.rs:26
pub const MALE_FORMANT_ESTIMATES: [f64; 4] = [320., 1440., 2760., 3200.]; // lib.rs:27
pub const FEMALE_FORMANT_ESTIMATES: [f64; 4] = [480., 1760., 3200., 3520.]; // lib.rs:28

/// spectrum.rs:149-154
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct Resonance<T> {
    pub frequency: T,
    pub bandwidth: T,
}

/// periodic.rs:306-310
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct Pitch<T> {
    pub frequency: T,
    pub strength: T,
}

pub struct Context {
    raw: *mut ffi::vbx_ctx,
}

impl Context {
    pub fn new(device: i32) -> VoxBoxResult<Context> {
        let mut raw = ptr::null_mut();
        let st = unsafe { ffi::vbx_ctx_create(device, &mut raw) };
        if st != ffi::VBX_OK {
            return Err(VoxBoxError::Cuda("vbx_ctx_create failed: no usable CUDA device (no CPU fallback)".into()));
        }
        Ok(Context { raw })
    }

    fn check(&self, st: c_int) -> VoxBoxResult<()> {
        use std::ffi::CStr;
        let msg = || unsafe { CStr::from_ptr(ffi::vbx_last_error(self.raw)).to_string_lossy().into_owned() };
        match st {
            ffi::VBX_OK => Ok(()),
            ffi::VBX_ERR_LPC => Err(VoxBoxError::LPC("Denum was <= 0.0")),
            ffi::VBX_ERR_PITCH => Err(VoxBoxError::Pitch("pitch candidate strength is NaN")),
            ffi::VBX_ERR_POLYNOMIAL => Err(VoxBoxError::Polynomial("Failed to find roots")),
            ffi::VBX_ERR_WORKSPACE => Err(VoxBoxError::Workspace),
            ffi::VBX_ERR_NOMEM => Err(VoxBoxError::NoMem),
            ffi::VBX_ERR_BADARG => Err(VoxBoxError::BadArg(msg())),
            _ => Err(VoxBoxError::Cuda(msg())),
        }
    }

    fn one_frame(x: &[f32], window: i32) -> ffi::vbx_frames {
        ffi::vbx_frames {
            base: x.as_ptr() as *const c_void,
            n_frames: 1,
            frame_stride: x.len() as i64,
            frames_per_segment: 0,
            segment_stride: 0,
            frame_len: x.len() as i32,
            dtype: ffi::VBX_F32,
            window,
            reserved: 0,
        }
    }

    // ---- periodic.rs:265-289 Autocorrelate --------------------------------------------------
    pub fn autocorrelate_mut(&self, x: &[f32], coeffs: &mut [f64]) -> VoxBoxResult<()> {
        let fr = Self::one_frame(x, ffi::VBX_WINDOW_NONE);
        self.check(unsafe {
            ffi::vbx_autocorrelate_host(self.raw, &fr, coeffs.len() as i32, coeffs.as_mut_ptr() as *mut c_void, ffi::VBX_F64)
        })
    }
    pub fn autocorrelate(&self, x: &[f32], n_coeffs: usize) -> VoxBoxResult<Vec<f64>> {
        let mut out = vec![0f64; n_coeffs];
        self.autocorrelate_mut(x, &mut out)?;
        Ok(out)
    }

    // ---- periodic.rs:291-304 `impl Autocorrelate for VecDeque<T>`: the same fold on a ring buffer --------------
    // The deque's storage is uploaded as it lies in memory (two slices = one ring with a head) and unrolled on the device.
    pub fn autocorrelate_deque(&self, x: &std::collections::VecDeque<f32>, n_coeffs: usize) -> VoxBoxResult<Vec<f64>> {
        let (front, back) = x.as_slices();
        // ring = back ++ front, logical element i = ring[(head + i) mod capacity] with head = back.len()
        let mut ring: Vec<f32> = Vec::with_capacity(x.len());
        ring.extend_from_slice(back);
        ring.extend_from_slice(front);
        let heads = [back.len() as i64];
        let dev = DeviceBuf::from_host(self, &ring)?;
        let dheads = DeviceBuf::from_host(self, &heads)?;
        let out = DeviceBuf::<f64>::new(self, n_coeffs)?;
        self.check(unsafe {
            ffi::vbx_autocorrelate_ring(self.raw, dev.ptr, ffi::VBX_F32, 1, ring.len() as i64, dheads.ptr as *const i64,
                                        ring.len() as i32, n_coeffs as i32, out.ptr, ffi::VBX_F64)
        })?;
        out.to_host(self)
    }

    // ---- spectrum.rs:50-92 LPC::{lpc_mut, lpc} on an already autocorrelated buffer ------------------
    pub fn lpc(&self, r: &[f64], n_coeffs: usize) -> VoxBoxResult<Vec<f64>> {
        let dev = DeviceBuf::from_host(self, r)?;
        let ac = DeviceBuf::<f64>::new(self, n_coeffs + 1)?;
        self.check(unsafe {
            ffi::vbx_lpc_levinson(self.raw, dev.ptr, ffi::VBX_F64, 1, r.len() as i32, n_coeffs as i32, ac.ptr, ptr::null_mut(), ffi::VBX_F64)
        })?;
        ac.to_host(self)
    }

    // ---- spectrum.rs:94-146 LPC::lpc_praat ----------------------------------------------------------
    pub fn lpc_praat(&self, x: &[f32], n_coeffs: usize) -> VoxBoxResult<Vec<f64>> {
        let dev = DeviceBuf::from_host(self, x)?;
        let mut fr = Self::one_frame(x, ffi::VBX_WINDOW_NONE);
        fr.base = dev.ptr;
        let co = DeviceBuf::<f64>::new(self, n_coeffs)?;
        let st = DeviceBuf::<u8>::new(self, 1)?;
        self.check(unsafe { ffi::vbx_lpc_burg(self.raw, &fr, n_coeffs as i32, co.ptr, st.ptr as *mut u8, ffi::VBX_F64) })?;
        self.check(st.to_host(self)?[0] as c_int)?; // Err(LPC("Denum was <= 0.0"))
        co.to_host(self)
    }

    // ---- polynomial.rs:79-152 Polynomial::find_roots --------------------------------------------------
    pub fn find_roots(&self, coeffs: &[Complex<f64>]) -> VoxBoxResult<Vec<Complex<f64>>> {
        let dev = DeviceBuf::from_host(self, coeffs)?;
        let out = DeviceBuf::<Complex<f64>>::new(self, coeffs.len())?;
        let st = DeviceBuf::<u8>::new(self, 1)?;
        self.check(unsafe {
            ffi::vbx_find_roots(self.raw, dev.ptr, ffi::VBX_F64, 1, coeffs.len() as i32, out.ptr, st.ptr as *mut u8)
        })?;
        self.check(st.to_host(self)?[0] as c_int)?;
        let mut roots = out.to_host(self)?;
        while roots.last().map_or(false, |z| z.re == 0. && z.im == 0.) {
            roots.pop(); // polynomial.rs:85-87
        }
        Ok(roots)
    }

    // ---- lib.rs:40-116 find_formants (one frame; `formants` is the tracker state, in/out) ----------------------
    pub fn find_formants(&self, buf: &[f32], sample_rate: f64, n_coeffs: usize, formants: &mut [Resonance<f64>]) -> VoxBoxResult<()> {
        let fr = Self::one_frame(buf, ffi::VBX_WINDOW_HANN_PERIODIC);
        let mut status = [0u8; 1];
        self.check(unsafe {
            ffi::vbx_find_formants_host(self.raw, &fr, sample_rate, n_coeffs as i32, ffi::VBX_LPC_BURG,
                                        formants.as_mut_ptr() as *mut c_void, formants.len() as i32, ptr::null_mut(),
                                        ptr::null_mut(), ptr::null_mut(), status.as_mut_ptr(), ffi::VBX_F64)
        })?;
        self.check(status[0] as c_int)
    }

    // ---- periodic.rs:356-456 Pitched::pitch::<Hanning> (frame already windowed by the caller) ---------------------
    pub fn pitch(&self, windowed: &[f32], sample_rate: f64, threshold: f64, min: f64, max: f64) -> VoxBoxResult<Vec<Pitch<f64>>> {
        let fr = Self::one_frame(windowed, ffi::VBX_WINDOW_NONE);
        let cap = windowed.len() / 4 + 2; // every other lag below N/2 a maximum, plus the unvoiced candidate
        let mut cand = vec![Pitch { frequency: 0f64, strength: 0f64 }; cap];
        let (mut n, mut st) = ([0i32; 1], [0u8; 1]);
        self.check(unsafe {
            ffi::vbx_pitch_host(self.raw, &fr, sample_rate, threshold, min, max, cap as i32, cand.as_mut_ptr() as *mut c_void,
                                n.as_mut_ptr(), st.as_mut_ptr(), ffi::VBX_F64)
        })?;
        self.check(st[0] as c_int)?;
        cand.truncate(n[0] as usize);
        Ok(cand)
    }

    // ---- periodic.rs:320-354 PitchExtractor: the strongest candidate of every frame (`candidates[frame][0]`) --------------
    // `candidates` = one `pitch()` result per frame.  With `viterbi = true` the opt-in path finder runs instead (the
    // reference declares `voiced_unvoiced_cost` but never uses it: periodic.rs:394-395); `false` is the reference's behaviour.
    pub fn pitch_extract(&self, candidates: &[Vec<Pitch<f64>>], voiced_unvoiced_cost: f64, viterbi: bool) -> VoxBoxResult<Vec<Pitch<f64>>> {
        let frames = candidates.len();
        let cap = candidates.iter().map(|c| c.len()).max().unwrap_or(0).max(1);
        let mut flat = vec![Pitch { frequency: 0f64, strength: 0f64 }; frames * cap];
        let mut counts = vec![0i32; frames];
        for (f, c) in candidates.iter().enumerate() {
            flat[f * cap..f * cap + c.len()].copy_from_slice(c);
            counts[f] = c.len() as i32;
        }
        let dev = DeviceBuf::from_host(self, &flat)?;
        let out = DeviceBuf::<Pitch<f64>>::new(self, frames)?;
        if viterbi {
            let dn = DeviceBuf::from_host(self, &counts)?;
            self.check(unsafe {
                ffi::vbx_pitch_viterbi(self.raw, dev.ptr, ffi::VBX_F64, dn.ptr as *const i32, 1, frames as i64, cap as i32,
                                       voiced_unvoiced_cost, 0.35, 0.01, 600.0, out.ptr, ptr::null_mut())
            })?;
        } else {
            self.check(unsafe { ffi::vbx_pitch_extract(self.raw, dev.ptr, ffi::VBX_F64, frames as i64, cap as i32, out.ptr) })?;
        }
        out.to_host(self)
    }

    // ---- spectrum.rs:371-441 MFCC::mfcc (frame already windowed by the caller) ----------------------------------------
    pub fn mfcc(&self, windowed: &[f32], num_coeffs: usize, freq_bounds: (f64, f64), sample_rate: f64) -> VoxBoxResult<Vec<f64>> {
        let fr = Self::one_frame(windowed, ffi::VBX_WINDOW_NONE);
        let mut out = vec![0f64; num_coeffs];
        self.check(unsafe {
            ffi::vbx_mfcc_host(self.raw, &fr, num_coeffs as i32, num_coeffs as i32, freq_bounds.0, freq_bounds.1, sample_rate,
                               out.as_mut_ptr() as *mut c_void, ffi::VBX_F64)
        })?;
        Ok(out)
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::vbx_ctx_destroy(self.raw) };
    }
}

pub fn hz_to_mel(hz: f64) -> f64 {
    unsafe { ffi::vbx_hz_to_mel(hz) }
}
pub fn mel_to_hz(mel: f64) -> f64 {
    unsafe { ffi::vbx_mel_to_hz(mel) }
}
pub fn find_formants_real_work_size(buf_len: usize, n_coeffs: usize) -> usize {
    unsafe { ffi::vbx_find_formants_real_work_size(buf_len as i64, n_coeffs as i64) as usize }
}
pub fn find_formants_complex_work_size(n_coeffs: usize) -> usize {
    unsafe { ffi::vbx_find_formants_complex_work_size(n_coeffs as i64) as usize }
}

/// Caller-owned HBM buffer (vbx_malloc / vbx_free).
pub struct DeviceBuf<T> {
    pub ptr: *mut c_void,
    pub len: usize,
    ctx: *mut ffi::vbx_ctx,
    _t: std::marker::PhantomData<T>,
}

impl<T: Clone + Default> DeviceBuf<T> {
    pub fn new(ctx: &Context, len: usize) -> VoxBoxResult<Self> {
        let mut p = ptr::null_mut();
        ctx.check(unsafe { ffi::vbx_malloc(ctx.raw, len * std::mem::size_of::<T>(), &mut p) })?;
        Ok(DeviceBuf { ptr: p, len, ctx: ctx.raw, _t: std::marker::PhantomData })
    }
    pub fn from_host(ctx: &Context, host: &[T]) -> VoxBoxResult<Self> {
        let b = Self::new(ctx, host.len())?;
        ctx.check(unsafe { ffi::vbx_memcpy_h2d(ctx.raw, b.ptr, host.as_ptr() as *const c_void, host.len() * std::mem::size_of::<T>()) })?;
        ctx.check(unsafe { ffi::vbx_sync(ctx.raw) })?;
        Ok(b)
    }
    pub fn to_host(&self, ctx: &Context) -> VoxBoxResult<Vec<T>> {
        let mut v = vec![T::default(); self.len];
        ctx.check(unsafe { ffi::vbx_memcpy_d2h(ctx.raw, v.as_mut_ptr() as *mut c_void, self.ptr, self.len * std::mem::size_of::<T>()) })?;
        ctx.check(unsafe { ffi::vbx_sync(ctx.raw) })?;
        Ok(v)
    }
}

impl<T> Drop for DeviceBuf<T> {
    fn drop(&mut self) {
        unsafe { ffi::vbx_free(self.ctx, self.ptr) };
    }
}

/// A batch of utterances resident in HBM: the intended way to drive the library (one call == all frames).
pub struct Batch<'a> {
    ctx: &'a Context,
    audio: DeviceBuf<f32>,
    pub n_utterances: usize,
    pub samples_per_utterance: usize,
}

impl<'a> Batch<'a> {
    pub fn upload(ctx: &'a Context, audio: &[f32], n_utterances: usize) -> VoxBoxResult<Batch<'a>> {
        Ok(Batch { ctx, audio: DeviceBuf::from_host(ctx, audio)?, n_utterances, samples_per_utterance: audio.len() / n_utterances })
    }
    /// `Windower::{hanning,rectangle}(.., bin, hop)` over every utterance as one strided view.
    pub fn frames(&self, bin: usize, hop: usize, window: i32) -> ffi::vbx_frames {
        let j = if self.samples_per_utterance < bin { 0 } else { (self.samples_per_utterance - bin) / hop + 1 };
        ffi::vbx_frames {
            base: self.audio.ptr,
            n_frames: (j * self.n_utterances) as i64,
            frame_stride: hop as i64,
            frames_per_segment: j as i64,
            segment_stride: self.samples_per_utterance as i64,
            frame_len: bin as i32,
            dtype: ffi::VBX_F32,
            window,
            reserved: 0,
        }
    }
    /// Hann → autocorrelate(p+1) → lpc(p) for every frame: returns ([F][p+1] r, [F][p+1] lpc) as fp32.
    pub fn lpc(&self, bin: usize, hop: usize, p: usize) -> VoxBoxResult<(Vec<f32>, Vec<f32>)> {
        let fr = self.frames(bin, hop, ffi::VBX_WINDOW_HANN_SYMMETRIC);
        let n = fr.n_frames as usize * (p + 1);
        let (r, ac) = (DeviceBuf::<f32>::new(self.ctx, n)?, DeviceBuf::<f32>::new(self.ctx, n)?);
        self.ctx.check(unsafe { ffi::vbx_lpc(self.ctx.raw, &fr, p as i32, r.ptr, ac.ptr, ptr::null_mut(), ffi::VBX_F32) })?;
        Ok((r.to_host(self.ctx)?, ac.to_host(self.ctx)?))
    }
}
