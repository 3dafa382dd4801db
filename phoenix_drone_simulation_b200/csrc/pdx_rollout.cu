// Rollout-collector kernels: GAE reverse scan and column moments.
//   pdx_gae     restates Buffer.finish_path / calculate_adv_and_value_targets /
//               discount_cumsum (algs/core.py:105-119, 458-479, 497-534) for a lock-step
//               [T][n] rollout: one thread per environment column walks time backwards and
//               restarts the recursions at episode boundaries.
//   pdx_moments restates the per-rank sums behind OnlineMeanStd.update
//               (utils/online_mean_std.py:70-84): sum x and sum (x - shift)^2 per column.
#include <cuda_runtime.h>
#include <cstdio>
#include "../../include/phoenix_b200.h"

namespace {

__global__ void __launch_bounds__(128) k_gae(int64_t T, int64_t n, const float* __restrict__ rew,
                                             const float* __restrict__ val, const uint8_t* __restrict__ done,
                                             const float* __restrict__ boot_val, const float* __restrict__ last_val,
                                             float gamma, float lam, float ret_scale, int use_scaling,
                                             float* __restrict__ adv, float* __restrict__ target_v,
                                             float* __restrict__ disc_ret) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float next_val = last_val[i];       // epoch cut: bootstrap with V(o_T)        (iwpg.py:376-378)
  float next_ret = next_val;          // rews = [..., last_val]                   (core.py:514)
  float next_adv = 0.0f;
  for (int64_t t = T - 1; t >= 0; --t) {
    const int64_t k = t * n + i;
    const uint8_t d = done[k];
    if (d == 1) { next_val = 0.0f; next_ret = 0.0f; next_adv = 0.0f; }          // terminated: v = 0
    else if (d == 2) { next_val = boot_val[k]; next_ret = next_val; next_adv = 0.0f; }  // time limit
    const float r = rew[k], v = val[k];
    const float ret = r + gamma * next_ret;                                      // core.py:518
    float rs = r;
    if (use_scaling) rs = fminf(fmaxf(r / ret_scale, -10.0f), 10.0f);            // core.py:527, oms clip
    const float delta = rs + gamma * next_val - v;                               // core.py:464
    const float a = delta + gamma * lam * next_adv;                              // core.py:465
    disc_ret[k] = ret;
    adv[k] = a;
    target_v[k] = a + v;                                                         // core.py:466
    next_val = v; next_ret = ret; next_adv = a;
  }
}

// blockDim = (cols padded to 32, rows per block); each block strides over row tiles.
__global__ void k_moments(int64_t rows, int dim, const float* __restrict__ x,
                          const double* __restrict__ shift, double* __restrict__ out) {
  extern __shared__ double sm[];                     // [2][blockDim.y][blockDim.x]
  const int d = threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  if (d < dim) {
    const double sh = shift ? shift[d] : 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += (int64_t)gridDim.x * blockDim.y) {
      const double v = (double)x[r * dim + d];
      s1 += v;
      s2 += (v - sh) * (v - sh);
    }
  }
  const int idx = threadIdx.y * blockDim.x + d;
  sm[idx] = s1;
  sm[blockDim.x * blockDim.y + idx] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && d < dim) {
    for (int y = 1; y < blockDim.y; ++y) {
      s1 += sm[y * blockDim.x + d];
      s2 += sm[blockDim.x * blockDim.y + y * blockDim.x + d];
    }
    atomicAdd(&out[d], s1);
    atomicAdd(&out[dim + d], s2);
  }
}

// Combines the episode-statistics vectors of all ranks after one all-gather:
// out[0..4) = column sums, out[4], out[6] = column minima, out[5], out[7] = column maxima.
__global__ void k_stats_combine(int world, const double* __restrict__ gathered, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= 8) return;
  double v = gathered[k];
  for (int r = 1; r < world; ++r) {
    const double x = gathered[r * 8 + k];
    v = k < 4 ? v + x : ((k & 1) ? fmax(v, x) : fmin(v, x));
  }
  out[k] = v;
}

// The library carries its own static CUDA runtime: select the device the data lives on.
int select_device_of(const void* ptr) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PDX_ERR_NO_DEVICE;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess || attr.type != cudaMemoryTypeDevice) {
    cudaGetLastError();
    return PDX_ERR_INVALID;
  }
  return cudaSetDevice(attr.device) == cudaSuccess ? PDX_OK : PDX_ERR_CUDA;
}

}  // namespace

extern "C" int pdx_gae(int64_t T, int64_t n, const float* rew, const float* val, const uint8_t* done,
                       const float* boot_val, const float* last_val, float gamma, float lam,
                       float ret_scale, int use_reward_scaling, float* adv, float* target_v,
                       float* disc_ret, void* stream) {
  if (T <= 0 || n <= 0 || !rew || !val || !done || !boot_val || !last_val || !adv || !target_v || !disc_ret)
    return PDX_ERR_INVALID;
  const int rc = select_device_of(rew);
  if (rc) return rc;
  const unsigned grid = (unsigned)((n + 127) / 128);
  k_gae<<<grid, 128, 0, (cudaStream_t)stream>>>(T, n, rew, val, done, boot_val, last_val, gamma, lam,
                                                ret_scale, use_reward_scaling, adv, target_v, disc_ret);
  return cudaGetLastError() == cudaSuccess ? PDX_OK : PDX_ERR_CUDA;
}

extern "C" int pdx_moments(int64_t rows, int32_t dim, const float* x, const double* shift, double* out,
                           void* stream) {
  if (rows <= 0 || dim <= 0 || dim > 1024 || !x || !out) return PDX_ERR_INVALID;
  const int rc = select_device_of(x);
  if (rc) return rc;
  const int bx = ((dim + 31) / 32) * 32;
  const int by = bx >= 256 ? 1 : 256 / bx;
  const dim3 block(bx, by);
  int64_t tiles = (rows + by - 1) / by;
  const unsigned grid = (unsigned)(tiles < 1184 ? tiles : 1184);      // 8 x 148 SMs
  const size_t smem = 2ull * bx * by * sizeof(double);
  k_moments<<<grid, block, smem, (cudaStream_t)stream>>>(rows, dim, x, shift, out);
  return cudaGetLastError() == cudaSuccess ? PDX_OK : PDX_ERR_CUDA;
}

extern "C" int pdx_stats_combine(int32_t world, const double* gathered, double* out, void* stream) {
  if (world <= 0 || !gathered || !out) return PDX_ERR_INVALID;
  const int rc = select_device_of(gathered);
  if (rc) return rc;
  k_stats_combine<<<1, 32, 0, (cudaStream_t)stream>>>(world, gathered, out);
  return cudaGetLastError() == cudaSuccess ? PDX_OK : PDX_ERR_CUDA;
}
