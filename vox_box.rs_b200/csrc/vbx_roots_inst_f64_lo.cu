// instantiates lpc_roots_kernel<P, double> for P = 2..12 (see vbx_roots_kernel.cuh)
#include "vbx_roots_kernel.cuh"
namespace vbx_roots {
void fill_f64_lo(roots_kernel_t* t) { RootsFill<double, 2, 12>::fill(t); }
}  // namespace vbx_roots
