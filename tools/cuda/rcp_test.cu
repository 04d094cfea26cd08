// Accuracy of rcp.approx.ftz.f64 (MUFU.RCP64H seed) and of the seed + 1 / 2 Newton steps, on [1e-3, 1e6].
#include <cstdio>
#include <cmath>
__global__ void k(double* err) {
    double e0 = 0, e1 = 0, e2 = 0;
    for (int i = threadIdx.x; i < 4000000; i += blockDim.x) {
        double x = 1e-3 * pow(1.0000052, (double)i);  // up to ~1e6
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        double ex = 1.0 / x;
        e0 = fmax(e0, fabs(r - ex) / ex);
        double e = fma(-x, r, 1.0); r = fma(r, e, r);
        e1 = fmax(e1, fabs(r - ex) / ex);
        e = fma(-x, r, 1.0); r = fma(r, e, r);
        e2 = fmax(e2, fabs(r - ex) / ex);
    }
    atomicMax((unsigned long long*)&err[0], __double_as_longlong(e0));
    atomicMax((unsigned long long*)&err[1], __double_as_longlong(e1));
    atomicMax((unsigned long long*)&err[2], __double_as_longlong(e2));
}
int main() {
    double* d; cudaMalloc(&d, 24); cudaMemset(d, 0, 24);
    k<<<1, 256>>>(d);
    double h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("rcp.approx.ftz.f64 max rel err: seed %.3e, +1 Newton %.3e, +2 Newton %.3e\n", h[0], h[1], h[2]);
    return 0;
}
