// instantiates lpc_roots_kernel<P, double> for P = 13..24 (see vbx_roots_kernel.cuh)
#include "vbx_roots_kernel.cuh"
namespace vbx_roots {
void fill_f64_hi(roots_kernel_t* t) { RootsFill<double, 13, 24>::fill(t); }
}  // namespace vbx_roots
