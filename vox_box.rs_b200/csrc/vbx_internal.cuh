// vbx_internal.cuh — shared internals of libvoxbox_b200 (context, error plumbing, device helpers).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/voxbox_b200.h"

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct vbx_mfcc_cache;  // vbx_mfcc.cu: device tables per (N, num_coeffs, bounds, fs)
void vbx_mfcc_cache_free(struct vbx_ctx* ctx);

struct vbx_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;  // max dynamic shared memory per block (opt-in)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    int64_t launches = 0;
    char err[512] = {0};
    // scratch arena (device) and pinned staging (host) — grow-only, reused across calls
    void* arena = nullptr;
    size_t arena_bytes = 0;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // A caller inside the library that already holds arena pointers (vbx_find_formants) hands its callees their scratch
    // as a sub-range of its own single reservation: while set, vbx_scratch_get serves from here and never touches the arena.
    void* sub_scratch = nullptr;
    size_t sub_scratch_bytes = 0;
    // host-call pipeline (vbx_pipeline.cuh): copy-in / copy-out streams, events, double-buffered device block
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_pipe[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void* pipe = nullptr;
    size_t pipe_bytes = 0;
    // side stream of vbx_find_formants (the latency-bound tracker of frame chunk c runs there while the main stream computes
    // the LPC / roots of chunk c + 1); ev_side: [0..1] chunk c's resonances ready, [2..3] chunk c's tracker done, by chunk parity
    cudaStream_t s_side = nullptr;
    cudaEvent_t ev_side[4] = {nullptr, nullptr, nullptr, nullptr};
    // window tables (device, f64), keyed by (kind << 32 | n)
    std::map<uint64_t, double*> windows;
    // per-kernel timing (vbx_profile_*): an event after every launch; a kernel's time = the gap to the previous event
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_events;   // pool, created lazily
    std::vector<const char*> prof_names;    // name of the launch that precedes event i+1 (event 0 = begin marker)
    size_t prof_used = 0;
    std::map<std::string, std::pair<double, int64_t>> prof_totals;  // name -> (ms, launches)
    // launches on the side stream are timed as explicit (start, stop) event pairs on that stream
    struct ProfRange { const char* name; cudaEvent_t e0, e1; };
    std::vector<ProfRange> prof_ranges;
    // executed-work counters of the data-dependent kernels (device, VBX_N_WORK_COUNTERS × u64; zeroed by vbx_profile_begin,
    // incremented only while profiling is on): see vbx_profile_counters in the header
    unsigned long long* work_counters = nullptr;
    int reserve_sms = 0;               // SMs the persistent kernels leave free for a co-running side-stream kernel
    bool roots_small = false;          // lpc_roots_pair_kernel in 96-thread CTAs (they fit beside a persistent LPC CTA)
    unsigned* tile_counter = nullptr;  // device: [0] the persistent kernels' dynamic tile cursor, [1] length of hard_list
    int* hard_list = nullptr;          // device: frames the pair-deflation roots kernel hands to the f64 fix-up launch
    static constexpr int kHardCap = 4096;
    vbx_mfcc_cache* mfcc_cache = nullptr;
    bool mfcc_fft_f32 = false;  // MFCC transform precision (default fp64)
};

enum { VBX_WORK_ROOTS_HORNER = 0, VBX_WORK_ROOTS_ROUNDS = 1, VBX_WORK_REFINE_TERMS = 2, VBX_WORK_REFINE_EVALS = 3, VBX_N_WORK_COUNTERS = 8 };
// device pointer of the counter block while profiling is on, else null (kernels skip the accounting)
static inline unsigned long long* vbx_work_ptr(vbx_ctx* ctx) { return ctx->prof_on ? ctx->work_counters : nullptr; }

int vbx_fail(vbx_ctx* ctx, int status, const char* fmt, ...);
void vbx_prof_mark(vbx_ctx* ctx, const char* name);        // records "launch `name` was just enqueued" (profiling on)
// side-stream launches: begin returns a slot (or -1), end closes it; counts the launch
int vbx_prof_range_begin(vbx_ctx* ctx, const char* name, cudaStream_t stream);
void vbx_prof_range_end(vbx_ctx* ctx, int slot, cudaStream_t stream);
int vbx_arena_reserve(vbx_ctx* ctx, size_t bytes);         // ensures ctx->arena has >= bytes
int vbx_pinned_reserve(vbx_ctx* ctx, size_t bytes);        // ensures ctx->pinned has >= bytes
// scratch for a kernel launcher: the caller-provided sub-range if one is set (see vbx_ctx::sub_scratch), else the arena
int vbx_scratch_get(vbx_ctx* ctx, size_t bytes, void** out);
// scratch the LPC stage of vbx_find_formants may ask for through vbx_scratch_get (0 when the fused kernels apply)
size_t vbx_lpc_scratch_bytes(vbx_ctx* ctx, const vbx_frames* fr, int n_lags);
size_t vbx_burg_scratch_bytes(vbx_ctx* ctx, const vbx_frames* fr);
int vbx_pipe_reserve(vbx_ctx* ctx, size_t bytes);          // ensures ctx->pipe has >= bytes
// cached device table (ones for NONE).  sample_dtype == VBX_I16 folds the PCM scale into the table: w[i] / 32767
// (tests/lib.rs:17-19 scale every sample by 1/(i32::MAX >> 16) before any arithmetic).
int vbx_get_window(vbx_ctx* ctx, int kind, int n, const double** dev_out, int sample_dtype = VBX_F32);
void vbx_window_fill_host(int kind, int n, double* out);

#define VBX_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return vbx_fail((ctx), VBX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                                     \
    } while (0)

#define VBX_CHECK_LAUNCH(ctx, name)                                                                  \
    do {                                                                                             \
        cudaError_t e__ = cudaGetLastError();                                                        \
        if (e__ != cudaSuccess)                                                                      \
            return vbx_fail((ctx), VBX_ERR_CUDA, "launch of %s failed: %s", (name), cudaGetErrorString(e__)); \
        (ctx)->launches++;                                                                           \
        if ((ctx)->prof_on) vbx_prof_mark((ctx), (name));                                            \
    } while (0)

#define VBX_REQUIRE(ctx, cond, ...)                                                                  \
    do {                                                                                             \
        if (!(cond)) return vbx_fail((ctx), VBX_ERR_BADARG, __VA_ARGS__);                            \
    } while (0)

static inline size_t vbx_dtype_size(int dt) { return dt == VBX_F64 ? 8 : (dt == VBX_I16 ? 2 : 4); }

// validates a vbx_frames descriptor (device or host pointers alike)
int vbx_check_frames(vbx_ctx* ctx, const vbx_frames* fr, bool allow_f64 = false);
// number of samples spanned by the strided view (0 if n_frames == 0)
static inline int64_t vbx_frames_per_segment(const vbx_frames* fr) {
    return fr->frames_per_segment > 0 ? fr->frames_per_segment : fr->n_frames;
}
static inline int64_t vbx_frames_extent(const vbx_frames* fr) {
    if (fr->n_frames <= 0) return 0;
    const int64_t J = vbx_frames_per_segment(fr), segs = fr->n_frames / J;
    return (segs - 1) * (fr->frames_per_segment > 0 ? fr->segment_stride : 0) + (J - 1) * fr->frame_stride + fr->frame_len;
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
template <typename T> __device__ __forceinline__ T vbx_ldg(const T* p) { return __ldg(p); }

// sample load with dtype dispatch (F32, or raw I16 PCM: the 1/32767 scale lives in the window table, vbx_get_window)
template <typename TIn> __device__ __forceinline__ float vbx_load_sample(const TIn* p);
template <> __device__ __forceinline__ float vbx_load_sample<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float vbx_load_sample<int16_t>(const int16_t* p) { return (float)__ldg(p); }
// double-valued loader for the paths that also accept f64 samples (Burg / find_formants)
template <typename TIn> __device__ __forceinline__ double vbx_load_sample_d(const TIn* p) { return (double)vbx_load_sample<TIn>(p); }
template <> __device__ __forceinline__ double vbx_load_sample_d<double>(const double* p) { return __ldg(p); }

template <typename TOut> __device__ __forceinline__ void vbx_store(TOut* p, double v) { *p = (TOut)v; }

__device__ __forceinline__ double vbx_shfl_xor(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }
#endif
