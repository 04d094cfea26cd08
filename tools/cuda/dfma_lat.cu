// DFMA / DADD dependent-issue latency and the ILP one warp needs: CH independent chains per thread, W warps per SM sub-partition;
// prints cycles per DFMA per warp (clock64) and the fraction of the FP64 pipe's peak (one warp-wide DFMA per 2 cycles per SMSP).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
    double v[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) v[c] = threadIdx.x + c;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) v[c] = fma(v[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += v[c];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH> void run(double* d, long long* c, int w) {
    const int iters = 2048;
    k<CH><<<148, w * 128>>>(d, c, iters, 0.999, 0.001);
    k<CH><<<148, w * 128>>>(d, c, iters, 0.999, 0.001);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / ((double)iters * 8 * CH);   // cycles per DFMA of one warp
    printf("chains=%2d warps/SMSP=%d: %.2f cycles per DFMA per warp, pipe use %.0f %%\n", CH, w, per, 100.0 * 2.0 * w / per);
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    for (int w : {1, 2, 4}) { run<1>(d, c, w); run<2>(d, c, w); run<4>(d, c, w); run<8>(d, c, w); run<16>(d, c, w); }
    return 0;
}
