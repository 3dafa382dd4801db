// Kernels: init (constructor), reset (masked) and the fused multi-step rollout kernel.
//
// k_rollout advances every environment by n_steps env.steps in ONE launch:
//   * one thread owns one environment; its state quads are loaded once (coalesced 128-bit
//     accesses), live in registers across all steps and sub-steps, and are stored once;
//   * each step's observation rows are assembled in a shared-memory tile that has exactly the
//     layout of the block's slice of the row-major [n_envs][obs_dim] output, and leave the SM
//     as ONE bulk asynchronous copy (cp.async.bulk shared->global, the TMA engine; SASS UBLKCP)
//     issued by one thread, double buffered so the copy of step t overlaps step t+1;
//   * episodes that end are reset inside the same launch.  The reset path is as long as a step
//     (two noisy observation calls, ~20 Philox calls) and only a few lanes of a warp need it, so
//     the warp generates the random draws of its finished environments together (shared table,
//     one Philox call per lane and pass) and only the arithmetic runs on the owner lanes;
//   * episode statistics accumulate per thread in shared memory and are reduced once per launch
//     (warp shuffles, one set of atomics per block).
#pragma once
#include "pdx_model.cuh"

namespace pdx {

constexpr int kMaxBlock = 512;     // threads per block are chosen at run time (<= kMaxBlock)

template <class T>
struct KArgs {
  DevCfg<T> c;
  PdxBuffers b;
  const float* actions;
  const uint8_t* mask;
  uint64_t seed, counter;
  double* dump_step;
  double* dump_reset;
  double* dump_init;
  int n_steps;
  int n_tiles;      // 2: observation tiles double buffered; 1: one tile, shifted in place
};

template <class T, int RNG>
__device__ __forceinline__ Rng<T, RNG> make_rng(const KArgs<T>& a, uint64_t counter, int64_t i,
                                                const double* tape, double* dump) {
  Rng<T, RNG> r;
  const uint64_t env = (uint64_t)(a.b.env_offset + i);
  r.key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  r.env_lo = (uint32_t)env;
  r.env_hi = (uint32_t)(env >> 32) ^ ((uint32_t)(counter >> 32) << 8);
  r.ctr_lo = (uint32_t)counter;
  r.tape = tape ? tape + i : nullptr;
  r.stride = a.b.n_envs;
  r.dump = dump ? dump + i : nullptr;
  return r;
}

// float32 quantisation of the initial position (quirk A.6-6: `pos` is a float32 array)
template <class T> __device__ __forceinline__ T f32q(T x) { return (T)(float)x; }
template <class T> __device__ __forceinline__ T unif(T lo, T hi, T u) { return lo + (hi - lo) * u; }

// ---------------------------------------------------------------------------------------------
//  DroneBaseEnv.reset for one env (base.py:382-431).  `m` holds the state to reset: only the
//  gyro bias (and, for the caller, the OU state) survive a reset (quirk A.6-8); `stale` are the
//  body rates of the previous episode's last state (base.py:411, quirk A.6-5).
//  Outputs: m.w (complete new state, OU words untouched), o1 / o2 = the two observation calls
//  (base.py:420,429); the history is H-1 copies of (o1, last_action) and one (o2, last_action).
// ---------------------------------------------------------------------------------------------
template <class T, int TASK, int PHYS, bool NOISE, int RNG, bool PID, class RG>
__device__ __forceinline__ void reset_env(Model<T, TASK, PHYS, NOISE, RNG, PID>& m, const RG& rng,
                                          const T stale[3], T* o1, T* o2) {
  typedef Model<T, TASK, PHYS, NOISE, RNG, PID> Mo;
  constexpr Layout L = Mo::L;
  const DevCfg<T>& c = m.c;
  T* w = m.w;
  const T pi = T(3.14159265358979323846);

  T la[4] = {T(0), T(0), T(0), T(0)};                    // drone.last_action = ring[-1]
  T ring[8] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
  T x[4] = {T(0), T(0), T(0), T(0)};
  T pos[3] = {c.init_xyz[0], c.init_xyz[1], c.init_xyz[2]};
  T q[4] = {T(0), T(0), T(0), T(1)};
  T vel[3] = {T(0), T(0), T(0)};
  T oms[3] = {T(0), T(0), T(0)};
  T rpy0[3] = {T(0), T(0), T(0)};                        // Euler angles the pose is built from
  int ref_off = 0;
  if constexpr (TASK == PDX_TASK_CIRCLE) ref_off = (int)w[L.ref_offset];

  if constexpr (TASK == PDX_TASK_TAKEOFF) {                        // takeoff.py:179-212
    if (c.reset_distribution) {
      T u[3];
      rng.template uniforms<3>(SITE_RESET, 0, u);
      pos[0] = f32q(pos[0] + unif(T(-0.25), T(0.25), u[0]));
      pos[1] = f32q(pos[1] + unif(T(-0.25), T(0.25), u[1]));
      rpy0[2] = unif(-pi, pi, u[2]);
      quat_from_euler(T(0), T(0), rpy0[2], q);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { la[k] = T(-1); ring[k] = T(-1); ring[4 + k] = T(-1); }
  } else if (c.reset_distribution) {
    T u[14];
    rng.template uniforms<14>(SITE_RESET, 0, u);
    T rpy[3];
    if constexpr (TASK == PDX_TASK_HOVER) {                        // hover.py:192-229
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        pos[k] = f32q(pos[k] + unif(T(-0.25), T(0.25), u[k]));
        rpy[k] = unif(-(pi / T(6)), pi / T(6), u[3 + k]);
        vel[k] = T(0) + unif(T(-0.1), T(0.1), u[7 + k]);
        const T lim = pi * T(200) / T(180);
        oms[k] = T(0) + unif(-lim, lim, u[10 + k]);
      }
      rpy[2] = unif(-(T(2) * pi), T(2) * pi, u[6]);
    } else {                                             // circle.py:213-256
      // np.random.randint(0, 300): the tape holds the integer itself; Philox: floor(300 u)
      const bool from_tape = RG::kTape && !rng.dumping();
      ref_off = from_tape ? (int)u[0] : min(299, (int)(u[0] * T(300)));
      rng.dump_value(0, (double)ref_off);
      T tp[3];
      m.ref_point(ref_off, tp);
      const T a0 = pi * T(20) / T(180), lim = pi * T(50) / T(180);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        pos[k] = tp[k] + unif(T(-0.05), T(0.05), u[1 + k]);
        rpy[k] = unif(-a0, a0, u[4 + k]);
        vel[k] = T(0) + unif(T(-0.1), T(0.1), u[8 + k]);
      }
      rpy[2] = unif(-(T(0.1) * pi), T(0.1) * pi, u[7]);
      oms[0] = unif(-lim, lim, u[11]);
      oms[1] = unif(-lim, lim, u[12]);
    }
    const T yl = pi * T(20) / T(180);
    oms[2] = unif(-yl, yl, u[13]);
    quat_from_euler(rpy[0], rpy[1], rpy[2], q);
#pragma unroll
    for (int k = 0; k < 3; ++k) rpy0[k] = rpy[k];
    T z[8];
    rng.template normals<8>(SITE_RESET + 4, 14, z);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      x[k] = c.hover_x + T(0.02) * z[k];
      ring[k] = clampT(c.hover_action + T(0.02) * z[4 + k], T(-1), T(1));
    }
    if (c.buf_size > 1) {
      rng.template normals<4>(SITE_RESET + 6, 22, z);
#pragma unroll
      for (int k = 0; k < 4; ++k) ring[4 + k] = clampT(c.hover_action + T(0.02) * z[k], T(-1), T(1));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) la[k] = c.buf_size > 1 ? ring[4 + k] : ring[k];
  }

  // pose / velocity hand-over through "PyBullet": hover.py:232-243, agents.py:434-453
  T R[9];
  rot_from_quat(q, R);
  T ww[3], ob[3];
  ww[0] = R[0] * oms[0] + R[3] * oms[1] + R[6] * oms[2];   // R^T w written as WORLD rate
  ww[1] = R[1] * oms[0] + R[4] * oms[1] + R[7] * oms[2];
  ww[2] = R[2] * oms[0] + R[5] * oms[1] + R[8] * oms[2];
  ob[0] = R[0] * ww[0] + R[3] * ww[1] + R[6] * ww[2];      // ... and R^T again (quirk A.6-4)
  ob[1] = R[1] * ww[0] + R[4] * ww[1] + R[7] * ww[2];
  ob[2] = R[2] * ww[0] + R[5] * ww[1] + R[8] * ww[2];

  // apply_domain_randomization: base.py:239-296, agents.py:208-224
  int slot = c.slots_reset_task;
  if (c.dr_on) {
    auto draw = [&](T v, T uu) { const T b = c.dr * v; return unif(v - b, v + b, uu); };
    T u[15];
    if (Mo::BULLET && c.use_motor_dynamics) rng.template uniforms<15>(SITE_DR, slot, u);
    else rng.template uniforms<7>(SITE_DR, slot, u);
    w[L.dt] = draw(c.time_step, u[0]);
    w[L.mass] = draw(c.mass, u[1]);
    w[L.inertia] = draw(c.inertia[0], u[2]);
    w[L.inertia + 1] = draw(c.inertia[1], u[3]);
    w[L.inertia + 2] = draw(c.inertia[2], u[4]);           // u[5]: ftf0 (cancels, unused)
    w[L.ftf1] = draw(c.ftf1, u[6]);
    if constexpr (Mo::BULLET) if (c.use_motor_dynamics) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const T Tm = M<T>::fmax(draw(c.motor_tc, u[7 + k]), w[L.dt]);
        w[L.motor_b + k] = w[L.dt] / Tm;
        w[L.motor_k + k] = c.k_mass_dr * c.gravity * draw(c.thrust2weight, u[11 + k]) / T(4);  // quirk A.6-7
      }
    }
  }
  slot += c.slots_reset_dr;

  // commit kinematics
#pragma unroll
  for (int k = 0; k < 3; ++k) { w[L.xyz + k] = pos[k]; w[L.vel + k] = vel[k]; }
  if constexpr (Mo::BULLET) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { w[L.quat + k] = q[k]; w[L.motor_x + k] = x[k]; w[L.ring + k] = ring[k]; w[L.ring + 4 + k] = ring[4 + k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) w[L.omega_world + k] = ww[k];
    w[L.ring_idx] = T(0);
  } else {
    // agents.py:446: rpy = getEulerFromQuaternion(q).  With |roll|, |pitch| < pi/2 (all reset
    // distributions) that round trip is the identity up to wrapping yaw into (-pi, pi]; the
    // float32 kernels use that, the float64 kernels do the reference's round trip.
    T e[3];
    if constexpr (sizeof(T) == 4) {
      e[0] = rpy0[0]; e[1] = rpy0[1];
      e[2] = rpy0[2] - T(6.283185307179586) * rintf(rpy0[2] * T(0.15915494309189535));
    } else {
      euler_from_quat(q, e);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { w[L.rpy + k] = e[k]; w[L.omega + k] = ob[k]; }
  }
  if constexpr (NOISE) {
#pragma unroll
    for (int k = 0; k < 3; ++k) w[L.gyro_lpf + k] = stale[k];
  }
  if constexpr (TASK == PDX_TASK_CIRCLE) w[L.ref_offset] = (T)ref_off;
#pragma unroll
  for (int k = 0; k < 4; ++k) w[L.last_action + k] = la[k];
  w[L.ep_return] = T(0);
  w[L.ep_length] = T(0);
  if constexpr (PID) {                                    // control.reset(): control.py:182-191,282-287
#pragma unroll
    for (int k = 0; k < 12; ++k) w[L.pid + k] = T(0);
  }

  // two observation calls (base.py:420,429)
  T target[3] = {c.target[0], c.target[1], c.target[2]};
  if constexpr (TASK == PDX_TASK_CIRCLE) m.ref_point(ref_off % 300, target);
  if constexpr (TASK == PDX_TASK_TAKEOFF) m.ref_point(0, target);
  m.observe(rng, SITE_RESET_OBS1, slot, target, la, q, o1);
  m.observe(rng, SITE_RESET_OBS2, slot + c.slots_obs_full, target, la, q, o2);
}

// history slots of the state <- H-1 entries, `entry(s, idx)` gives word idx of slot s
template <class T, int E, int QH, class F>
__device__ __forceinline__ void store_history(T* state, int64_t n, int64_t i, int first_quad, int H, F entry) {
  for (int s = 0; s < H - 1; ++s) {
#pragma unroll
    for (int qd = 0; qd < QH; ++qd) {
      T v[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const int idx = qd * 4 + l;
        v[l] = idx < E ? entry(s, idx < E ? idx : 0) : T(0);
      }
      store_quad(state, n, i, first_quad + s * QH + qd, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
//  constructor: zero state, nominal parameters, base.py:143's compute_observation()
// ---------------------------------------------------------------------------------------------
template <class T, int TASK, int PHYS, bool NOISE, int RNG, bool PID>
__global__ void __launch_bounds__(kMaxBlock) k_init(const KArgs<T> a) {
  typedef Model<T, TASK, PHYS, NOISE, RNG, PID> Mo;
  constexpr Layout L = Mo::L;
  const int64_t n = a.b.n_envs;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Mo m(a.c);
  const DevCfg<T>& c = a.c;
#pragma unroll
  for (int k = 0; k < Mo::NW; ++k) m.w[k] = T(0);
  m.w[L.xyz + 2] = T(1);                                  // agents.py:32
  m.w[L.dt] = c.time_step;
  m.w[L.mass] = c.mass;
#pragma unroll
  for (int k = 0; k < 3; ++k) m.w[L.inertia + k] = c.inertia[k];
  m.w[L.ftf1] = c.ftf1;
  if constexpr (Mo::BULLET) {
    m.w[L.quat + 3] = T(1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      m.w[L.motor_b + k] = T(1) * c.time_step / c.motor_tc;   // agents.py:203-204
      m.w[L.motor_k + k] = c.max_thrust;                      // agents.py:200
    }
  }
  if constexpr (NOISE) {
    const Rng<T, RNG> rng = make_rng<T, RNG>(a, a.counter, i, a.b.tape_init, a.dump_init);
    T z[3];
    rng.template normals<3>(SITE_INIT, 12, z);
#pragma unroll
    for (int k = 0; k < 3; ++k) m.w[L.gyro_bias + k] = c.gyro_sigma_b * z[k];
  }
  T* state = reinterpret_cast<T*>(a.b.state);
  m.store(state, n, i, true);
  const T zero[4] = {T(0), T(0), T(0), T(0)};
  for (int qd = 0; qd < (c.history - 1) * Mo::QH; ++qd) store_quad(state, n, i, L.n_quads + qd, zero);
}

// ---------------------------------------------------------------------------------------------
//  explicit reset of all / masked environments (dense: one thread per env)
// ---------------------------------------------------------------------------------------------
template <class T, int TASK, int PHYS, bool NOISE, int RNG, bool PID>
__global__ void __launch_bounds__(kMaxBlock) k_reset(const KArgs<T> a) {
  typedef Model<T, TASK, PHYS, NOISE, RNG, PID> Mo;
  constexpr Layout L = Mo::L;
  constexpr int C = Mo::C, E = Mo::E, QH = Mo::QH;
  const int64_t n = a.b.n_envs;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (a.mask && !a.mask[i]) return;
  Mo m(a.c);
  T* state = reinterpret_cast<T*>(a.b.state);
  m.load(state, n, i);
  const Rng<T, RNG> rng = make_rng<T, RNG>(a, a.counter, i, a.b.tape_reset, a.dump_reset);
  T stale[3], o1[C], o2[C];
  m.body_rates(stale);
  reset_env(m, rng, stale, o1, o2);
  m.store(state, n, i, true);
  const int H = a.c.history;
  const T* la = &m.w[L.last_action];
  T* row = reinterpret_cast<T*>(a.b.obs) + i * a.c.obs_dim;
  for (int j = 0; j < H; ++j) {
    const bool newest = j == H - 1;
#pragma unroll
    for (int k = 0; k < C; ++k) row[j * E + k] = newest ? o2[k] : o1[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) row[j * E + C + k] = la[k];
  }
  store_history<T, E, QH>(state, n, i, L.n_quads, H, [&](int s, int idx) {
    return idx < C ? (s == H - 2 ? o2[idx < C ? idx : 0] : o1[idx < C ? idx : 0]) : la[idx - C < 4 && idx >= C ? idx - C : 0];
  });
}

// CAS-based min/max on doubles for the episode statistics
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *p;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long assumed = old;
    old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *p;
  while (__longlong_as_double((long long)old) < v) {
    const unsigned long long assumed = old;
    old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

// ---- bulk asynchronous copy shared -> global (TMA engine), bulk-group completion ------------
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Shared-memory plan of k_rollout (dynamic shared memory):
//   T tile0, tile1 [B][D]                     observation rows of the block, layout == global slice
//                                             (tile1 only when n_tiles == 2)
//   T rtab [B/32][kResetRows*4][kResetChunk]  per warp: reset draws of up to kResetChunk finished envs
//   double acc_sum [4][B], T acc_ext [4][B]   per-thread episode statistics (n, sum ret, sum ret^2,
//                                             sum len; min/max ret, min/max len), reduced once
//   int tile_free                             last step whose bulk copy is known to have left its tile
//   unsigned char fin_lane [B]                per warp: lanes that finished this step, in lane order
template <class T>
__host__ __device__ inline size_t rollout_smem_bytes(int block, int D, int n_tiles) {
  size_t bytes = ((size_t)n_tiles * block * D + (size_t)(block / 32) * kResetRows * 4 * kResetChunk + (size_t)4 * block) * sizeof(T);
  bytes = (bytes + 15) & ~(size_t)15;
  return bytes + sizeof(double) * 4 * block + 16 + kMaxBlock;
}

// ---------------------------------------------------------------------------------------------
//  fused multi-step env.step (+ auto-reset)
// ---------------------------------------------------------------------------------------------
template <class T, int TASK, int PHYS, bool NOISE, int RNG, bool PID>
__global__ void __launch_bounds__(kMaxBlock, 1) k_rollout(const KArgs<T> a) {
  typedef Model<T, TASK, PHYS, NOISE, RNG, PID> Mo;
  constexpr Layout L = Mo::L;
  constexpr int C = Mo::C, E = Mo::E, QH = Mo::QH;
  const DevCfg<T>& c = a.c;
  const int64_t n = a.b.n_envs;
  const int B = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t block_base = (int64_t)blockIdx.x * B;
  const int64_t i = block_base + tid;
  const bool valid = i < n;
  const int rows = (int)min((int64_t)B, n - block_base);
  const int D = c.obs_dim, H = c.history;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile0 = reinterpret_cast<T*>(smem_raw);
  const int NT = a.n_tiles;
  T* tile1 = NT == 2 ? tile0 + (size_t)B * D : tile0;
  T* rtab = tile0 + (size_t)NT * B * D + (size_t)warp * (kResetRows * 4 * kResetChunk);
  T* acc_ext = tile0 + (size_t)NT * B * D + (size_t)(B >> 5) * (kResetRows * 4 * kResetChunk);
  const size_t off = (((size_t)NT * B * D + (size_t)(B >> 5) * kResetRows * 4 * kResetChunk + (size_t)4 * B) * sizeof(T) + 15) & ~(size_t)15;
  double* acc_sum = reinterpret_cast<double*>(smem_raw + off);
  int* s_tile_free = reinterpret_cast<int*>(smem_raw + off + sizeof(double) * 4 * B);   // accessed with atomics only
  unsigned char* s_fin_lane = smem_raw + off + sizeof(double) * 4 * B + 16 + (warp << 5);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc_sum[k * B + tid] = 0.0;
    acc_ext[k * B + tid] = (k & 1) ? T(-1e30) : T(1e30);
  }
  if (tid == 0) *s_tile_free = -1;
  __syncthreads();

  Mo m(c);
  T* state = reinterpret_cast<T*>(a.b.state);
  T* my_row0 = tile0 + (size_t)tid * D;
  T* my_row1 = tile1 + (size_t)tid * D;
  // Programmatic dependent launch: this grid may have started while the kernel before it on the stream is
  // still draining.  Nothing that kernel could have written is touched before griddepcontrol.wait (which
  // returns once it has completed and its writes are visible).  With PDX_BUF_STATE_STABLE the caller
  // guarantees that kernel does not write `state` (collector: it is the policy kernel), so the state
  // load below runs under its tail and only the actions wait.
  const bool state_stable = (a.b.flags & PDX_BUF_STATE_STABLE) != 0;
  if (!state_stable) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  if (valid) {
    m.load(state, n, i);
    // history slots -> entries 1..H-1 of the "previous row" (tile1 plays the old tile at t = 0)
    for (int s = 0; s < H - 1; ++s) {
#pragma unroll
      for (int qd = 0; qd < QH; ++qd) {
        T v[4];
        load_quad(state, n, i, L.n_quads + s * QH + qd, v);
#pragma unroll
        for (int l = 0; l < 4; ++l) if (qd * 4 + l < E) my_row1[(s + 1) * E + qd * 4 + l] = v[l];
      }
    }
  }
  bool any_fin = false;
  const bool latency = Mo::BULLET && c.use_latency;
  T* w = m.w;
  // the action of step t+1 is requested while step t computes (an L2/HBM miss otherwise sits at
  // the head of every step's dependency chain)
  float4 a_next = make_float4(0.f, 0.f, 0.f, 0.f);
  if (state_stable) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  if (valid) a_next = reinterpret_cast<const float4*>(a.actions)[i];

  for (int t = 0; t < a.n_steps; ++t) {
    T* tn = (t & 1) ? my_row1 : my_row0;             // this step's row, previous step's row
    const T* to = (t & 1) ? my_row0 : my_row1;
    const int64_t tn_off = (int64_t)t * n;           // offset of step t in the [n_steps][n] outputs
    bool fin = false;
    T ep_ret_out = T(0);
    int ep_len_out = 0;
    T core[C], a_new[4], actT[4];
    int n_ep = 0;

    if (valid) {
      const float4 a4 = a_next;
      if (t + 1 < a.n_steps) a_next = reinterpret_cast<const float4*>(a.actions)[tn_off + n + i];
      const float act[4] = {a4.x, a4.y, a4.z, a4.w};
      const Rng<T, RNG> rng = make_rng<T, RNG>(a, a.counter + (uint64_t)t, i,
                                               a.b.tape_step ? a.b.tape_step + (int64_t)t * c.slots_step * n : nullptr,
                                               a.dump_step ? a.dump_step + (int64_t)t * c.slots_step * n : nullptr);
      n_ep = (int)w[L.ep_length] + 1;                // 1-based step index in the episode
      T la_prev[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { la_prev[k] = w[L.last_action + k]; actT[k] = (T)act[k]; }

      // ---- physics sub-steps; the observation call after each one is discarded by the
      // reference (base.py:464) but advances the gyro bias / low-pass state (quirk A.6-1)
      int slot = 0;
      for (int s = 0; s < c.agg; ++s) {
        const bool full = (s % c.obs_rate) == 0;
        m.substep(rng, act, s, slot, full);
        slot += 4 + (NOISE ? (full ? c.slots_obs_full : c.slots_obs_gyro) : 0);
      }

      // ---- the observation that is returned (base.py:468 -> compute_history)
      T target[3] = {c.target[0], c.target[1], c.target[2]};
      if constexpr (TASK == PDX_TASK_CIRCLE) m.ref_point((n_ep + (int)w[L.ref_offset]) % 300, target);   // circle.py:130
      if constexpr (TASK == PDX_TASK_TAKEOFF) m.ref_point(min(n_ep * c.agg, 299), target);                // takeoff.py:108
      T q_true[4] = {T(0), T(0), T(0), T(1)};
      if constexpr (!NOISE) {
        if constexpr (Mo::BULLET) { for (int k = 0; k < 4; ++k) q_true[k] = w[L.quat + k]; }
        else quat_from_euler(w[L.rpy], w[L.rpy + 1], w[L.rpy + 2], q_true);
      }
      m.observe(rng, SITE_FINAL_OBS, slot, target, actT, q_true, core);

      // a(k-1) paired with o(k).  quirk (Bullet agent): after a reset the action deque holds H
      // references to ring[-1] (agents.py:386, base.py:426-427), which the latency ring
      // overwrites in place with the current action -> those entries read as the *current*
      // action for as long as they stay in the deque.
      {
        const bool alias = latency && (H - 1 + n_ep <= H);
#pragma unroll
        for (int k = 0; k < 4; ++k) a_new[k] = alias ? actT[k] : la_prev[k];
      }

      // ---- reward / cost / done
      T e[3], om[3];
      m.euler(e);
      m.body_rates(om);
      const bool dn = m.done(e, om, target);
      // action penalties are float32 arithmetic in the reference (float32 action array)
      float nca2 = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float v = 0.5f * (fminf(fmaxf(act[k], -1.0f), 1.0f) + 1.0f);
        nca2 += v * v;
      }
      const T pa = (T)((float)c.pen_action * sqrtf(nca2));
      T par = T(0);
      if constexpr (TASK == PDX_TASK_CIRCLE) {                        // circle.py:186 (hover/takeoff: == 0, A.6-10)
        T d2 = T(0);
        if (!(latency && n_ep == 1)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { const T d = actT[k] - la_prev[k]; d2 += d * d; }
        }
        par = c.arp * M<T>::sqrt(d2);
      }
      const T prpy = c.pen_angle * norm3(e[0], e[1], e[2]);
      const T pspin = c.pen_spin * norm3(om[0], om[1], om[2]);
      const T pterm = dn ? c.pen_terminal : T(0);
      const T cvel = TASK == PDX_TASK_TAKEOFF ? c.pen_action : c.pen_velocity;   // takeoff.py:165
      const T pvel = cvel * norm3(w[L.vel], w[L.vel + 1], w[L.vel + 2]);
      const T penalties = ((((prpy + par) + pspin) + pvel) + pa) + pterm;
      const T dist = norm3(w[L.xyz] - target[0], w[L.xyz + 1] - target[1], w[L.xyz + 2] - target[2]);
      T r = -dist - penalties;
      if (TASK == PDX_TASK_TAKEOFF && w[L.xyz + 2] < T(0.08)) r -= T(1);
      const T cst = m.cost(e, om, act);

      // ---- episode accounting, TimeLimit (__init__.py:11)
      w[L.ep_return] += r;
      w[L.ep_length] = (T)n_ep;
      bool trunc = n_ep >= c.max_episode_steps;
      if (c.reset_on_nonfinite) {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 6; ++k) ok = ok && M<T>::finite(w[L.xyz + k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) ok = ok && M<T>::finite(om[k]) && M<T>::finite(e[k]);
        trunc = trunc || !ok;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) w[L.last_action + k] = actT[k];
      reinterpret_cast<T*>(a.b.reward)[tn_off + i] = r;
      reinterpret_cast<T*>(a.b.cost)[tn_off + i] = cst;
      a.b.terminated[tn_off + i] = dn ? 1 : 0;
      a.b.truncated[tn_off + i] = trunc ? 1 : 0;
      fin = dn || trunc;
      ep_ret_out = w[L.ep_return];
      ep_len_out = n_ep;
      if (a.b.episode_return) reinterpret_cast<T*>(a.b.episode_return)[tn_off + i] = fin ? ep_ret_out : T(0);
      if (a.b.episode_length) a.b.episode_length[tn_off + i] = fin ? ep_len_out : 0;
    }

    // ---- the tile of this step was last read by the bulk copy of step t - n_tiles: thread 0
    // publishes the newest step whose copy has drained (it waits right after issuing each copy;
    // with a single tile the rows are shifted in place, ascending, once the previous copy left)
    if (t >= NT) { while (atomicAdd(s_tile_free, 0) < t - NT) {} }

    // history emission: [o(k-H+1), a(k-H), ..., o(k), a(k-1)]  (base.py:303-319).
    // The tile has the row-major layout of the output, so lane l's row starts at bank l*D mod 32:
    // when D is a multiple of 16 a warp walking its rows in step hits few banks (D = 160: ONE
    // bank).  Each lane then walks its row rotated by its lane id (bank l*(D+1) + k: conflict
    // free); the new entry goes through a warp-private column-major staging buffer because
    // registers cannot be indexed by the rotated position.
    if ((D & 15) == 0) {                                   // >= 16-way conflicts in natural order
      const int rot = lane;
      if (valid) {
        for (int j = 0; j < H - 1; ++j) {
          const bool alias = latency && (j + n_ep <= H);
#pragma unroll
          for (int k = 0; k < E; ++k) {
            int kk = k + rot;
            kk = kk >= E ? kk - E : kk;
            kk = kk >= E ? kk - E : kk;
            T v = to[(j + 1) * E + kk];
            if (alias && kk >= C) v = kk == C ? actT[0] : kk == C + 1 ? actT[1] : kk == C + 2 ? actT[2] : actT[3];
            tn[j * E + kk] = v;
          }
        }
      }
      static_assert(E * 32 <= kResetRows * 4 * kResetChunk, "staging buffer must fit the reset table");
      T* stg = rtab;                                       // the reset table is idle at this point
#pragma unroll
      for (int k = 0; k < C; ++k) stg[k * 32 + lane] = core[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) stg[(C + k) * 32 + lane] = a_new[k];
      __syncwarp();
      if (valid) {
#pragma unroll
        for (int k = 0; k < E; ++k) {
          int kk = k + rot;
          kk = kk >= E ? kk - E : kk;
          kk = kk >= E ? kk - E : kk;
          tn[(H - 1) * E + kk] = stg[kk * 32 + lane];
        }
      }
      __syncwarp();
    } else if (valid) {                                    // other strides: at most 8-way (measured faster)
      for (int j = 0; j < H - 1; ++j) {
        const bool alias = latency && (j + n_ep <= H);
#pragma unroll
        for (int k = 0; k < E; ++k) {
          T v = to[(j + 1) * E + k];
          if (k >= C && alias) v = actT[k >= C ? k - C : 0];
          tn[j * E + k] = v;
        }
      }
#pragma unroll
      for (int k = 0; k < C; ++k) tn[(H - 1) * E + k] = core[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) tn[(H - 1) * E + C + k] = a_new[k];
    }

    // ---- auto-reset of finished episodes.  The reset path is as long as a step, mostly random
    // number generation, and only a few lanes of a warp need it: the warp generates the draws of
    // its finished environments TOGETHER (one Philox call per lane and pass, into a shared table)
    // and only the remaining arithmetic runs divergent on the owner lanes.  No block barrier.
    if (fin) {
      any_fin = true;
      acc_sum[0 * B + tid] += 1.0;
      acc_sum[1 * B + tid] += (double)ep_ret_out;
      acc_sum[2 * B + tid] += (double)ep_ret_out * (double)ep_ret_out;
      acc_sum[3 * B + tid] += (double)ep_len_out;
      acc_ext[0 * B + tid] = M<T>::fmin(acc_ext[0 * B + tid], ep_ret_out);
      acc_ext[1 * B + tid] = M<T>::fmax(acc_ext[1 * B + tid], ep_ret_out);
      acc_ext[2 * B + tid] = M<T>::fmin(acc_ext[2 * B + tid], (T)ep_len_out);
      acc_ext[3 * B + tid] = M<T>::fmax(acc_ext[3 * B + tid], (T)ep_len_out);
    }
    const bool do_reset = fin && c.auto_reset;
    const unsigned ballot = __ballot_sync(0xffffffffu, do_reset);
    if (ballot) {
      if (do_reset && a.b.final_obs) {                     // last observation of the episode
        T* fo = reinterpret_cast<T*>(a.b.final_obs) + (tn_off + i) * D;
        for (int k = 0; k < D; ++k) fo[k] = tn[k];
      }
      const uint64_t ctr = a.counter + (uint64_t)t;
      T o1[C], o2[C], stale[3];
      if (do_reset) {
        m.body_rates(stale);
        if (c.reset_on_nonfinite) {                        // extension: do not seed the next episode's
#pragma unroll
          for (int k = 0; k < 3; ++k) if (!M<T>::finite(stale[k])) stale[k] = T(0);   // low pass with inf / NaN
        }
      }
      if constexpr (RNG == PDX_RNG_PHILOX) {
        const int my_rank = __popc(ballot & ((1u << lane) - 1u));
        const int total = __popc(ballot);
        if (do_reset) s_fin_lane[my_rank] = (unsigned char)lane;
        __syncwarp();
        for (int chunk = 0; chunk < total; chunk += kResetChunk) {
          const int cnt = min(kResetChunk, total - chunk);
          // work item = (site row, finished env): consecutive lanes take consecutive envs.
          // The rows a reset draws form four runs (task, DR, first and second observation call).
          constexpr int n_task = TASK == PDX_TASK_TAKEOFF ? 1 : (Mo::BULLET ? 7 : 6);
          constexpr int n_dr = Mo::BULLET ? 4 : 2, n_obs = NOISE ? 6 : 0;
          constexpr int n_rows = n_task + n_dr + 2 * n_obs;
          const unsigned rcp = (65536u + (unsigned)cnt - 1u) / (unsigned)cnt;
          for (int item = lane; item < cnt * n_rows; item += 32) {
            const int ri = (int)(((unsigned)item * rcp) >> 16), slot = item - ri * cnt;
            const int row = ri < n_task ? ri
                          : ri < n_task + n_dr ? (int)(SITE_DR - SITE_RESET) + ri - n_task
                          : ri < n_task + n_dr + n_obs ? (int)(SITE_RESET_OBS1 - SITE_RESET) + ri - n_task - n_dr
                          : (int)(SITE_RESET_OBS2 - SITE_RESET) + ri - n_task - n_dr - n_obs;
            const uint32_t site = SITE_RESET + (uint32_t)row;
            const int owner = (int)s_fin_lane[chunk + slot];
            Rng<T, RNG> rr = make_rng<T, RNG>(a, ctr, block_base + (warp << 5) + owner, nullptr, nullptr);
            const uint4 r4 = rr.raw(site);
            T v[4];
            if (reset_site_is_normal(site)) {
              M<T>::box_muller(r4.x, r4.y, &v[0], &v[1]);
              M<T>::box_muller(r4.z, r4.w, &v[2], &v[3]);
            } else {
              v[0] = M<T>::unit(r4.x); v[1] = M<T>::unit(r4.y); v[2] = M<T>::unit(r4.z); v[3] = M<T>::unit(r4.w);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) rtab[(row * 4 + k) * kResetChunk + slot] = v[k];
          }
          __syncwarp();
          if (do_reset && my_rank >= chunk && my_rank < chunk + kResetChunk) {
            TableRng<T> tr;
            tr.tab = rtab + (my_rank - chunk);
            reset_env(m, tr, stale, o1, o2);
          }
          __syncwarp();
        }
      } else {
        if (do_reset) {
          const Rng<T, RNG> rr = make_rng<T, RNG>(a, ctr, i,
                                                  a.b.tape_reset ? a.b.tape_reset + (int64_t)t * c.slots_reset * n : nullptr,
                                                  a.dump_reset ? a.dump_reset + (int64_t)t * c.slots_reset * n : nullptr);
          reset_env(m, rr, stale, o1, o2);
        }
      }
      if (do_reset) {                                      // first observation of the new episode
        for (int j = 0; j < H; ++j) {
#pragma unroll
          for (int k = 0; k < C; ++k) tn[j * E + k] = (j == H - 1) ? o2[k] : o1[k];
#pragma unroll
          for (int k = 0; k < 4; ++k) tn[j * E + C + k] = w[L.last_action + k];
        }
      }
    }

    // ---- observation rows of the block leave as one bulk copy (or a flat coalesced copy when
    // the destination does not meet the 16-byte rules of the bulk engine)
    T* gdst = reinterpret_cast<T*>(a.b.obs) + (tn_off + block_base) * D;
    const T* tile = (t & 1) ? tile1 : tile0;
    const uint32_t bytes = (uint32_t)rows * (uint32_t)D * (uint32_t)sizeof(T);
    const bool bulk = ((reinterpret_cast<uintptr_t>(gdst) | bytes) & 15u) == 0;
    if (bulk) {
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        bulk_store(gdst, tile, bytes);
        if (NT == 2) { bulk_wait_read<1>(); atomicExch(s_tile_free, t - 1); }   // the copy issued one step ago has drained
        else { bulk_wait_read<0>(); atomicExch(s_tile_free, t); }
      }
    } else {
      __syncthreads();
      for (int e = tid; e < rows * D; e += B) gdst[e] = tile[e];
      __syncthreads();
      if (tid == 0) atomicExch(s_tile_free, t);
    }
  }

  // ---- epilogue: state and history back to HBM, statistics, drain the bulk engine
  if (valid) {
    m.store(state, n, i, true);
    const T* last = ((a.n_steps - 1) & 1) ? my_row1 : my_row0;
    store_history<T, E, QH>(state, n, i, L.n_quads, H, [&](int s, int idx) { return last[(s + 1) * E + idx]; });
  }
  if (a.b.episode_stats && __syncthreads_or(any_fin ? 1 : 0)) {
    if (warp == 0) {
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (k == 4 || k == 6) ? 1e300 : (k == 5 || k == 7) ? -1e300 : 0.0;
      for (int col = lane; col < B; col += 32) {
        if (acc_sum[col] > 0.0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] += acc_sum[k * B + col];
          v[4] = fmin(v[4], (double)acc_ext[0 * B + col]); v[5] = fmax(v[5], (double)acc_ext[1 * B + col]);
          v[6] = fmin(v[6], (double)acc_ext[2 * B + col]); v[7] = fmax(v[7], (double)acc_ext[3 * B + col]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        v[4] = fmin(v[4], __shfl_xor_sync(0xffffffffu, v[4], o));
        v[5] = fmax(v[5], __shfl_xor_sync(0xffffffffu, v[5], o));
        v[6] = fmin(v[6], __shfl_xor_sync(0xffffffffu, v[6], o));
        v[7] = fmax(v[7], __shfl_xor_sync(0xffffffffu, v[7], o));
      }
      if (lane == 0) {
        double* gs = a.b.episode_stats;
        atomicAdd(&gs[0], v[0]); atomicAdd(&gs[1], v[1]); atomicAdd(&gs[2], v[2]); atomicAdd(&gs[3], v[3]);
        atomic_min_double(&gs[4], v[4]); atomic_max_double(&gs[5], v[5]);
        atomic_min_double(&gs[6], v[6]); atomic_max_double(&gs[7], v[7]);
      }
    }
  }
  if (tid == 0) bulk_wait_all();
}

}  // namespace pdx
