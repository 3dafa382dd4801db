// Instantiates the float kernels of the bullet physics flavour, PWM control (see pdx_dispatch.cuh).
#include "pdx_dispatch.cuh"
namespace pdx {
cudaError_t launch_f32_bullet(int kind, const LaunchArgs& la) { return launch_tu<float, PDX_PHYSICS_BULLET, false>(kind, la); }
}  // namespace pdx
