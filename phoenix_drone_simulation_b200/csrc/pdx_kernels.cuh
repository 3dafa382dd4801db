// Kernels: init (constructor), reset (masked) and the fused step with in-kernel auto-reset.
// One thread per environment; state quads are loaded once, kept in registers across all
// physics sub-steps of the env.step and stored once.
#pragma once
#include "pdx_model.cuh"

namespace pdx {

constexpr int kBlock = 128;

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
};

template <class T, int RNG>
__device__ __forceinline__ Rng<T, RNG> make_rng(const KArgs<T>& a, int64_t i, const double* tape,
                                                double* dump) {
  Rng<T, RNG> r;
  const uint64_t env = (uint64_t)(a.b.env_offset + i);
  r.key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  r.env_lo = (uint32_t)env;
  r.env_hi = (uint32_t)(env >> 32) ^ ((uint32_t)(a.counter >> 32) << 8);
  r.ctr_lo = (uint32_t)a.counter;
  r.tape = tape ? tape + i : nullptr;
  r.stride = a.b.n_envs;
  r.dump = dump ? dump + i : nullptr;
  return r;
}

// float32 quantisation of the initial position (quirk A.6-6: `pos` is a float32 array)
template <class T> __device__ __forceinline__ T f32q(T x) { return (T)(float)x; }
template <class T> __device__ __forceinline__ T unif(T lo, T hi, T u) { return lo + (hi - lo) * u; }

// ---------------------------------------------------------------------------------------------
//  DroneBaseEnv.reset for one env.  Keeps OU state and gyro bias (never reset, quirk A.6-8).
// ---------------------------------------------------------------------------------------------
template <class T, int TASK, int PHYS, bool NOISE, int RNG>
__device__ __forceinline__ void reset_env(Model<T, TASK, PHYS, NOISE, RNG>& m, const Rng<T, RNG>& rng,
                                       T* state, int64_t n, int64_t i, T* obs_row) {
  typedef Model<T, TASK, PHYS, NOISE, RNG> Mo;
  constexpr Layout L = Mo::L;
  constexpr int C = Mo::C, E = Mo::E, QH = Mo::QH;
  const DevCfg<T>& c = m.c;
  T* w = m.w;
  const T pi = T(3.14159265358979323846);

  T stale[3];
  m.body_rates(stale);                                   // base.py:411 (quirk A.6-5)

  T la[4] = {T(0), T(0), T(0), T(0)};                    // drone.last_action = ring[-1]
  T ring[8] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
  T x[4] = {T(0), T(0), T(0), T(0)};
  T pos[3] = {c.init_xyz[0], c.init_xyz[1], c.init_xyz[2]};
  T q[4] = {T(0), T(0), T(0), T(1)};
  T vel[3] = {T(0), T(0), T(0)};
  T oms[3] = {T(0), T(0), T(0)};
  int ref_off = 0;
  if constexpr (TASK == PDX_TASK_CIRCLE) ref_off = (int)w[L.ref_offset];

  if constexpr (TASK == PDX_TASK_TAKEOFF) {                        // takeoff.py:179-212
    if (c.reset_distribution) {
      T u[3];
      rng.template uniforms<3>(SITE_RESET, 0, u);
      pos[0] = f32q(pos[0] + unif(T(-0.25), T(0.25), u[0]));
      pos[1] = f32q(pos[1] + unif(T(-0.25), T(0.25), u[1]));
      quat_from_euler(T(0), T(0), unif(-pi, pi, u[2]), q);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { la[k] = T(-1); ring[k] = T(-1); ring[4 + k] = T(-1); }
  } else if (c.reset_distribution) {
    T u[14];
    rng.template uniforms<4>(SITE_RESET + 0, 0, &u[0]);
    rng.template uniforms<4>(SITE_RESET + 1, 4, &u[4]);
    rng.template uniforms<4>(SITE_RESET + 2, 8, &u[8]);
    rng.template uniforms<2>(SITE_RESET + 3, 12, &u[12]);
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
      const bool from_tape = RNG == PDX_RNG_TAPE && !rng.dump;
      ref_off = from_tape ? (int)u[0] : min(299, (int)(u[0] * T(300)));
      if (RNG == PDX_RNG_TAPE && rng.dump) rng.dump[0] = (double)ref_off;
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
    T z[4];
    rng.template normals<4>(SITE_RESET + 4, 14, z);
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = c.hover_x + T(0.02) * z[k];
    rng.template normals<4>(SITE_RESET + 5, 18, z);
#pragma unroll
    for (int k = 0; k < 4; ++k) ring[k] = clampT(c.hover_action + T(0.02) * z[k], T(-1), T(1));
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
    T u[4];
    auto draw = [&](T v, T uu) { const T b = c.dr * v; return unif(v - b, v + b, uu); };
    rng.template uniforms<4>(SITE_DR + 0, slot, u);
    w[L.dt] = draw(c.time_step, u[0]);
    w[L.mass] = draw(c.mass, u[1]);
    w[L.inertia] = draw(c.inertia[0], u[2]);
    w[L.inertia + 1] = draw(c.inertia[1], u[3]);
    rng.template uniforms<3>(SITE_DR + 1, slot + 4, u);    // u[1]: ftf0 (cancels, unused)
    w[L.inertia + 2] = draw(c.inertia[2], u[0]);
    w[L.ftf1] = draw(c.ftf1, u[2]);
    if constexpr (Mo::BULLET) if (c.use_motor_dynamics) {
      T t2[4];
      rng.template uniforms<4>(SITE_DR + 2, slot + 7, u);
      rng.template uniforms<4>(SITE_DR + 3, slot + 11, t2);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const T Tm = M<T>::fmax(draw(c.motor_tc, u[k]), w[L.dt]);
        w[L.motor_b + k] = w[L.dt] / Tm;
        w[L.motor_k + k] = c.k_mass_dr * c.gravity * draw(c.thrust2weight, t2[k]) / T(4);  // quirk A.6-7
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
    T e[3];
    euler_from_quat(q, e);
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

  // two observation calls (base.py:420,429) and the history fill (base.py:424-427)
  T target[3] = {c.target[0], c.target[1], c.target[2]};
  if constexpr (TASK == PDX_TASK_CIRCLE) m.ref_point(ref_off % 300, target);
  if constexpr (TASK == PDX_TASK_TAKEOFF) m.ref_point(0, target);
  T o1[C], o2[C];
  m.observe(rng, SITE_RESET_OBS1, slot, target, la, q, o1);
  m.observe(rng, SITE_RESET_OBS2, slot + c.slots_obs_full, target, la, q, o2);
  const int H = c.history;
  const int g = (int)w[L.hist_phase];
  for (int j = 0; j < H; ++j) {
    const bool newest = j == H - 1;
#pragma unroll
    for (int k = 0; k < C; ++k) obs_row[j * E + k] = newest ? o2[k] : o1[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) obs_row[j * E + C + k] = la[k];
  }
  if (H > 1) {
    const int newest_pos = (g + H - 2) % (H - 1);
    for (int s = 0; s < H - 1; ++s) {
      const bool nw = s == newest_pos;
#pragma unroll
      for (int qd = 0; qd < QH; ++qd) {
        T v[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const int idx = qd * 4 + l;
          v[l] = idx < C ? (nw ? o2[idx < C ? idx : 0] : o1[idx < C ? idx : 0])
                         : (idx < E ? la[idx - C < 4 ? idx - C : 0] : T(0));
        }
        store_quad(state, n, i, L.n_quads + s * QH + qd, v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
//  constructor: zero state, nominal parameters, base.py:143's compute_observation()
// ---------------------------------------------------------------------------------------------
template <class T, int TASK, int PHYS, bool NOISE, int RNG>
__global__ void __launch_bounds__(kBlock) k_init(const KArgs<T> a) {
  typedef Model<T, TASK, PHYS, NOISE, RNG> Mo;
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
    const Rng<T, RNG> rng = make_rng<T, RNG>(a, i, a.b.tape_init, a.dump_init);
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

template <class T, int TASK, int PHYS, bool NOISE, int RNG>
__global__ void __launch_bounds__(kBlock) k_reset(const KArgs<T> a) {
  typedef Model<T, TASK, PHYS, NOISE, RNG> Mo;
  const int64_t n = a.b.n_envs;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (a.mask && !a.mask[i]) return;
  Mo m(a.c);
  T* state = reinterpret_cast<T*>(a.b.state);
  m.load(state, n, i);
  const Rng<T, RNG> rng = make_rng<T, RNG>(a, i, a.b.tape_reset, a.dump_reset);
  reset_env(m, rng, state, n, i, reinterpret_cast<T*>(a.b.obs) + i * a.c.obs_dim);
  m.store(state, n, i, true);
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

// ---------------------------------------------------------------------------------------------
//  fused env.step (+ auto-reset)
// ---------------------------------------------------------------------------------------------
template <class T, int TASK, int PHYS, bool NOISE, int RNG>
__global__ void __launch_bounds__(kBlock) k_step(const KArgs<T> a) {
  typedef Model<T, TASK, PHYS, NOISE, RNG> Mo;
  constexpr Layout L = Mo::L;
  constexpr int C = Mo::C, E = Mo::E, QH = Mo::QH;
  const DevCfg<T>& c = a.c;
  const int64_t n = a.b.n_envs;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;

  __shared__ double s_stats[8];
  __shared__ int s_any;
  if (a.b.episode_stats) {
    if (threadIdx.x == 0) s_any = 0;
    if (threadIdx.x < 8)
      s_stats[threadIdx.x] = (threadIdx.x == 4 || threadIdx.x == 6) ? 1e300 : (threadIdx.x == 5 || threadIdx.x == 7) ? -1e300 : 0.0;
    __syncthreads();
  }

  bool fin = false;
  T ep_ret_out = T(0);
  int ep_len_out = 0;
  Mo m(c);
  T* state = reinterpret_cast<T*>(a.b.state);
  T* obs_row = nullptr;

  if (valid) {
    T* w = m.w;
    m.load(state, n, i);
    const float4 a4 = reinterpret_cast<const float4*>(a.actions)[i];
    const float act[4] = {a4.x, a4.y, a4.z, a4.w};
    const Rng<T, RNG> rng = make_rng<T, RNG>(a, i, a.b.tape_step, a.dump_step);
    const int n_ep = (int)w[L.ep_length] + 1;             // 1-based step index in the episode
    T la_prev[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) la_prev[k] = w[L.last_action + k];

    // ---- physics sub-steps; the observation call after each one is discarded by the
    // reference (base.py:464) but advances the gyro bias / low-pass state (quirk A.6-1)
    int slot = 0;
    for (int s = 0; s < c.agg; ++s) {
      if constexpr (Mo::BULLET) m.physics_bullet(rng, act, s, slot); else m.physics_simple(rng, act, s, slot);
      slot += 4;
      if constexpr (NOISE) {
        T om[3];
        m.body_rates(om);
        const bool full = (s % c.obs_rate) == 0;
        m.gyro_update(rng, SITE_SUBSTEP + 4 * s + 1, slot + (full ? 12 : 0), om);
        slot += full ? c.slots_obs_full : c.slots_obs_gyro;
      }
    }

    // ---- the observation that is returned (base.py:468 -> compute_history)
    T target[3] = {c.target[0], c.target[1], c.target[2]};
    if constexpr (TASK == PDX_TASK_CIRCLE) m.ref_point((n_ep + (int)w[L.ref_offset]) % 300, target);   // circle.py:130
    if constexpr (TASK == PDX_TASK_TAKEOFF) m.ref_point(min(n_ep * c.agg, 299), target);                // takeoff.py:108
    T actT[4] = {(T)act[0], (T)act[1], (T)act[2], (T)act[3]};
    T q_true[4] = {T(0), T(0), T(0), T(1)};
    if constexpr (!NOISE) {
      if constexpr (Mo::BULLET) { for (int k = 0; k < 4; ++k) q_true[k] = w[L.quat + k]; }
      else quat_from_euler(w[L.rpy], w[L.rpy + 1], w[L.rpy + 2], q_true);
    }
    T core[C];
    m.observe(rng, SITE_FINAL_OBS, slot, target, actT, q_true, core);

    // ---- history emission: [o(k-H+1), a(k-H), ..., o(k), a(k-1)]  (base.py:303-319)
    const int H = c.history;
    const int g = (int)w[L.hist_phase];
    obs_row = reinterpret_cast<T*>(a.b.obs) + i * c.obs_dim;
    // quirk (Bullet agent): after a reset the action deque holds H references to ring[-1]
    // (agents.py:386, base.py:426-427), which the latency ring overwrites in place with the
    // current action -> those entries read as the *current* action.
    const bool latency = Mo::BULLET && c.use_latency;
    for (int j = 0; j < H - 1; ++j) {
      const int pos = (g + j) % (H - 1);
      const bool alias = latency && (j + n_ep <= H);
      for (int qd = 0; qd < QH; ++qd) {
        T v[4];
        load_quad(state, n, i, L.n_quads + pos * QH + qd, v);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const int idx = qd * 4 + l;
          if (idx < E) {
            T val = v[l];
            if (alias && idx >= C) val = actT[idx - C];
            obs_row[j * E + idx] = val;
          }
        }
      }
    }
    T a_new[4];                                          // a(k-1) paired with o(k)
    {
      const bool alias = latency && (H - 1 + n_ep <= H);
#pragma unroll
      for (int k = 0; k < 4; ++k) a_new[k] = alias ? actT[k] : la_prev[k];
#pragma unroll
      for (int k = 0; k < C; ++k) obs_row[(H - 1) * E + k] = core[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) obs_row[(H - 1) * E + C + k] = a_new[k];
    }
    if (H > 1) {                                         // newest entry replaces the oldest
      const int pos0 = g % (H - 1);
#pragma unroll
      for (int qd = 0; qd < QH; ++qd) {
        T v[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const int idx = qd * 4 + l;
          v[l] = idx < C ? core[idx < C ? idx : 0] : (idx < E ? la_prev[idx - C < 4 ? idx - C : 0] : T(0));
        }
        store_quad(state, n, i, L.n_quads + pos0 * QH + qd, v);
      }
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
    w[L.hist_phase] = (T)((g + 1) % (H > 1 ? 4 * (H - 1) : 1));
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
    reinterpret_cast<T*>(a.b.reward)[i] = r;
    reinterpret_cast<T*>(a.b.cost)[i] = cst;
    a.b.terminated[i] = dn ? 1 : 0;
    a.b.truncated[i] = trunc ? 1 : 0;
    fin = dn || trunc;
    ep_ret_out = w[L.ep_return];
    ep_len_out = n_ep;
    if (a.b.episode_return) reinterpret_cast<T*>(a.b.episode_return)[i] = fin ? ep_ret_out : T(0);
    if (a.b.episode_length) a.b.episode_length[i] = fin ? ep_len_out : 0;
  }

  // ---- per-block episode statistics: warp shuffles, then one set of atomics per block
  if (a.b.episode_stats) {
    const unsigned any = __ballot_sync(0xffffffffu, fin);
    if (any) {
      double cnt = fin ? 1.0 : 0.0, sr = fin ? (double)ep_ret_out : 0.0, sl = fin ? (double)ep_len_out : 0.0;
      double sr2 = sr * sr;
      double mn = fin ? (double)ep_ret_out : 1e300, mx = fin ? (double)ep_ret_out : -1e300;
      double ln = fin ? (double)ep_len_out : 1e300, lx = fin ? (double)ep_len_out : -1e300;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        sr2 += __shfl_xor_sync(0xffffffffu, sr2, o);
        sl += __shfl_xor_sync(0xffffffffu, sl, o);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        ln = fmin(ln, __shfl_xor_sync(0xffffffffu, ln, o));
        lx = fmax(lx, __shfl_xor_sync(0xffffffffu, lx, o));
      }
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_stats[0], cnt); atomicAdd(&s_stats[1], sr); atomicAdd(&s_stats[2], sr2);
        atomicAdd(&s_stats[3], sl);
        atomic_min_double(&s_stats[4], mn); atomic_max_double(&s_stats[5], mx);
        atomic_min_double(&s_stats[6], ln); atomic_max_double(&s_stats[7], lx);
        s_any = 1;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_any) {
      double* gs = a.b.episode_stats;
      atomicAdd(&gs[0], s_stats[0]); atomicAdd(&gs[1], s_stats[1]); atomicAdd(&gs[2], s_stats[2]);
      atomicAdd(&gs[3], s_stats[3]);
      atomic_min_double(&gs[4], s_stats[4]); atomic_max_double(&gs[5], s_stats[5]);
      atomic_min_double(&gs[6], s_stats[6]); atomic_max_double(&gs[7], s_stats[7]);
    }
  }

  if (valid) {
    const bool do_reset = fin && c.auto_reset;
    if (do_reset) {                                      // auto-reset in the same launch
      if (a.b.final_obs) {
        T* fo = reinterpret_cast<T*>(a.b.final_obs) + i * c.obs_dim;
        for (int k = 0; k < c.obs_dim; ++k) fo[k] = obs_row[k];
      }
      const Rng<T, RNG> rr = make_rng<T, RNG>(a, i, a.b.tape_reset, a.dump_reset);
      reset_env(m, rr, state, n, i, obs_row);
    }
    m.store(state, n, i, do_reset);
  }
}

}  // namespace pdx
