// vbx_waves.cu — waves.rs helpers, batched over signals (one warp per signal, fp64 arithmetic).
//
// Replaces:
//   waves.rs:10-23  RMS::rms                      sqrt(Σ x² / N)
//   waves.rs:25-37  Amplitude::amplitude          |x|
//   waves.rs:39-59  MaxAmplitude::max_amplitude   fold from |x[0]| with `>` (a NaN at index 0 sticks, later NaNs never win)
//   waves.rs:61-76  Normalize::{normalize, normalize_with_max}   x *= 1/max (no zero guard)
//   waves.rs:82-96  Filter::preemphasis           anti-causal additive IIR: y[N−1] = x[N−1]; y[i] = x[i] + 2π·factor·y[i+1]
// Signals are rows of a [n_signals][stride] array of f32 or f64 (stride >= n).
#include "vbx_internal.cuh"

namespace {

template <typename T> __device__ __forceinline__ double ldv(const T* p) { return (double)*p; }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += vbx_shfl_xor(v, m);
    return v;
}

// max_amplitude with the reference's fold semantics; all lanes return the value
template <typename T> __device__ __forceinline__ double max_amplitude_warp(const T* x, int n, int lane) {
    double pm = -1.0;
    for (int i = 1 + lane; i < n; i += 32) {
        const double a = fabs(ldv(x + i));
        if (a > pm) pm = a;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const double o = vbx_shfl_xor(pm, m);
        if (o > pm) pm = o;
    }
    const double a0 = fabs(ldv(x));
    return (a0 != a0) ? a0 : (pm > a0 ? pm : a0);
}

template <typename T>
__global__ void __launch_bounds__(128) rms_kernel(const T* __restrict__ x, int64_t n_signals, int n, int64_t stride, T* out, int op) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_signals) return;
    const T* row = x + s * stride;
    double v;
    if (op == 0) {
        double acc = 0.;
        for (int i = lane; i < n; i += 32) { const double a = ldv(row + i); acc = fma(a, a, acc); }
        v = sqrt(warp_sum_d(acc) / (double)n);
    } else {
        v = max_amplitude_warp(row, n, lane);
    }
    if (lane == 0) out[s] = (T)v;
}

template <typename T>
__global__ void __launch_bounds__(128) normalize_kernel(T* x, int64_t n_signals, int n, int64_t stride, const T* maxes) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_signals) return;
    T* row = x + s * stride;
    const double mx = maxes ? (double)maxes[s] : max_amplitude_warp(row, n, lane);
    const double scale = 1.0 / mx;
    __syncwarp();
    for (int i = lane; i < n; i += 32) row[i] = (T)(ldv(row + i) * scale);
}

// y[i] = x[i] + a·y[i+1] from the end.  Lane ℓ owns the contiguous chunk [ℓC, ℓC + C); pass 1 runs the recurrence
// with a zero carry to get the chunk's head value and a^len, a serial lane scan (31 → 0) resolves the carries, pass 2
// re-runs every chunk with its true carry and stores.
template <typename T>
__global__ void __launch_bounds__(128) preemphasis_kernel(T* x, int64_t n_signals, int n, int64_t stride, double a) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_signals) return;
    T* row = x + s * stride;
    const int C = (n + 31) / 32;
    const int lo = lane * C, hi = min(n, lo + C);  // [lo, hi)
    double head = 0., apow = 1.;
    for (int i = hi - 1; i >= lo; --i) {
        head = fma(a, head, ldv(row + i));
        apow *= a;
    }
    // carry into lane ℓ = y[(ℓ+1)C] = head_{ℓ+1} + a^{len_{ℓ+1}}·carry_{ℓ+1}
    double carry = 0.;  // lane 31 (or the last non-empty lane): y[N] does not exist — but y[N−1] = x[N−1] needs no `a` term
    double y_next = 0.;
    for (int l = 31; l >= 0; --l) {
        const double h = __shfl_sync(0xffffffffu, head, l), ap = __shfl_sync(0xffffffffu, apow, l);
        if (lane == l) carry = y_next;
        y_next = fma(ap, y_next, h);
    }
    double last = carry;
    for (int i = hi - 1; i >= lo; --i) {
        // waves.rs:91: x[i] + last·filter, except the final sample which is copied
        const double v = (i == n - 1) ? ldv(row + i) : fma(a, last, ldv(row + i));
        row[i] = (T)v;
        last = v;
    }
}

template <typename T>
int launch_reduce(vbx_ctx* ctx, const void* x, int64_t n_signals, int n, int64_t stride, void* out, int op) {
    const int64_t grid = (n_signals + 3) / 4;
    VBX_REQUIRE(ctx, grid <= 0x7fffffffLL, "too many signals for one launch");
    rms_kernel<T><<<(unsigned)grid, 128, 0, ctx->stream>>>((const T*)x, n_signals, n, stride, (T*)out, op);
    VBX_CHECK_LAUNCH(ctx, op == 0 ? "rms_kernel" : "max_amplitude_kernel");
    return VBX_OK;
}

int waves_check(vbx_ctx* ctx, const void* x, int dtype, int64_t n_signals, int n, int64_t stride) {
    VBX_REQUIRE(ctx, dtype == VBX_F32 || dtype == VBX_F64, "dtype must be VBX_F32 or VBX_F64");
    VBX_REQUIRE(ctx, n_signals >= 0, "n_signals < 0");
    VBX_REQUIRE(ctx, n >= 1, "signal length must be >= 1 (the reference indexes self[0])");
    VBX_REQUIRE(ctx, stride >= n, "stride must be >= n");
    VBX_REQUIRE(ctx, n_signals == 0 || x != nullptr, "x is NULL");
    return VBX_OK;
}

}  // namespace

extern "C" {

int vbx_rms(vbx_ctx* ctx, const void* x, int32_t dtype, int64_t n_signals, int32_t n, int64_t stride, void* out) {
    if (!ctx) return VBX_ERR_BADARG;
    int st = waves_check(ctx, x, dtype, n_signals, n, stride);
    if (st != VBX_OK || n_signals == 0) return st;
    VBX_REQUIRE(ctx, out != nullptr, "out is NULL");
    cudaSetDevice(ctx->device);
    return dtype == VBX_F64 ? launch_reduce<double>(ctx, x, n_signals, n, stride, out, 0)
                            : launch_reduce<float>(ctx, x, n_signals, n, stride, out, 0);
}

int vbx_max_amplitude(vbx_ctx* ctx, const void* x, int32_t dtype, int64_t n_signals, int32_t n, int64_t stride, void* out) {
    if (!ctx) return VBX_ERR_BADARG;
    int st = waves_check(ctx, x, dtype, n_signals, n, stride);
    if (st != VBX_OK || n_signals == 0) return st;
    VBX_REQUIRE(ctx, out != nullptr, "out is NULL");
    cudaSetDevice(ctx->device);
    return dtype == VBX_F64 ? launch_reduce<double>(ctx, x, n_signals, n, stride, out, 1)
                            : launch_reduce<float>(ctx, x, n_signals, n, stride, out, 1);
}

int vbx_normalize(vbx_ctx* ctx, void* x_inout, int32_t dtype, int64_t n_signals, int32_t n, int64_t stride, const void* maxes) {
    if (!ctx) return VBX_ERR_BADARG;
    int st = waves_check(ctx, x_inout, dtype, n_signals, n, stride);
    if (st != VBX_OK || n_signals == 0) return st;
    cudaSetDevice(ctx->device);
    const int64_t grid = (n_signals + 3) / 4;
    VBX_REQUIRE(ctx, grid <= 0x7fffffffLL, "too many signals for one launch");
    if (dtype == VBX_F64) normalize_kernel<double><<<(unsigned)grid, 128, 0, ctx->stream>>>((double*)x_inout, n_signals, n, stride, (const double*)maxes);
    else normalize_kernel<float><<<(unsigned)grid, 128, 0, ctx->stream>>>((float*)x_inout, n_signals, n, stride, (const float*)maxes);
    VBX_CHECK_LAUNCH(ctx, "normalize_kernel");
    return VBX_OK;
}

int vbx_preemphasis(vbx_ctx* ctx, void* x_inout, int32_t dtype, int64_t n_signals, int32_t n, int64_t stride, double factor) {
    if (!ctx) return VBX_ERR_BADARG;
    int st = waves_check(ctx, x_inout, dtype, n_signals, n, stride);
    if (st != VBX_OK || n_signals == 0) return st;
    cudaSetDevice(ctx->device);
    const int64_t grid = (n_signals + 3) / 4;
    VBX_REQUIRE(ctx, grid <= 0x7fffffffLL, "too many signals for one launch");
    const double a = 2.0 * 3.14159265358979323846264338327950288 * factor;  // waves.rs:88
    if (dtype == VBX_F64) preemphasis_kernel<double><<<(unsigned)grid, 128, 0, ctx->stream>>>((double*)x_inout, n_signals, n, stride, a);
    else preemphasis_kernel<float><<<(unsigned)grid, 128, 0, ctx->stream>>>((float*)x_inout, n_signals, n, stride, a);
    VBX_CHECK_LAUNCH(ctx, "preemphasis_kernel");
    return VBX_OK;
}

}  // extern "C"
