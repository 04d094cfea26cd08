// Accuracy of rcp.approx.ftz.f64 (MUFU.RCP64H seed), of the seed + 1 / 2 Newton steps and of the seed + one third-order
// step r(1 + e + e^2) (what rcp_pos in vbx_pitch.cu uses), on [1e-3, 1e6].
#include <cstdio>
#include <cmath>
__global__ void k(double* err) {
    double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    for (int i = threadIdx.x; i < 4000000; i += blockDim.x) {
        double x = 1e-3 * pow(1.0000052, (double)i);  // up to ~1e6
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        double ex = 1.0 / x;
        e0 = fmax(e0, fabs(r - ex) / ex);
        {
            const double eh = fma(-x, r, 1.0);
            const double rh = fma(r, fma(eh, eh, eh), r);
            e3 = fmax(e3, fabs(rh - ex) / ex);
        }
        double e = fma(-x, r, 1.0); r = fma(r, e, r);
        e1 = fmax(e1, fabs(r - ex) / ex);
        e = fma(-x, r, 1.0); r = fma(r, e, r);
        e2 = fmax(e2, fabs(r - ex) / ex);
    }
    atomicMax((unsigned long long*)&err[0], __double_as_longlong(e0));
    atomicMax((unsigned long long*)&err[1], __double_as_longlong(e1));
    atomicMax((unsigned long long*)&err[2], __double_as_longlong(e2));
    atomicMax((unsigned long long*)&err[3], __double_as_longlong(e3));
}
int main() {
    double* d; cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
    k<<<1, 256>>>(d);
    double h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("rcp.approx.ftz.f64 max rel err: seed %.3e, +1 Newton %.3e, +2 Newton %.3e, third-order step %.3e\n", h[0], h[1], h[2], h[3]);
    return 0;
}
