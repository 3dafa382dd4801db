// Instantiates the double kernels of the simple physics flavour, PWM control (see pdx_dispatch.cuh).
#include "pdx_dispatch.cuh"
namespace pdx {
cudaError_t launch_f64_simple(int kind, const LaunchArgs& la) { return launch_tu<double, PDX_PHYSICS_SIMPLE, false>(kind, la); }
}  // namespace pdx
