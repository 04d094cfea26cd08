// How many warps per SM sub-partition does the FP64 pipe need?  13 independent DFMA chains per thread (the LPC kernel's
// accumulator count), W warps per SMSP resident; prints the achieved TFLOP/s per W.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters, double a, double b) {
    double v[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) v[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) v[c] = fma(v[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += v[c];
    if (s == 123.456) out[0] = s;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    for (int w : {1, 2, 3, 4, 5, 6, 8, 12, 16}) {   // warps per SMSP -> threads per SM = w*4*32, one block per SM
        int threads = w * 4 * 32; if (threads > 1024) { threads = 1024; }
        int blocks = sms * ((w * 4 * 32 + 1023) / 1024);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k<13><<<blocks, threads>>>(d, iters, 0.999, 0.001);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 13 * 4 * (double)iters * blocks * threads;
        printf("warps/SMSP=%2d  %.2f TFLOP/s\n", w, fl / ms / 1e9);
    }
    return 0;
}
