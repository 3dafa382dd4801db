// Instantiates the float kernels of the simple physics flavour, PID control modes (see pdx_dispatch.cuh).
#include "pdx_dispatch.cuh"
namespace pdx {
cudaError_t launch_f32_simple_pid(int kind, const LaunchArgs& la) { return launch_tu<float, PDX_PHYSICS_SIMPLE, true>(kind, la); }
}  // namespace pdx
