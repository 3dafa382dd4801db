// The fused Crazyflie env.step / env.reset model: one thread owns one environment, its
// whole state sits in registers from the first load to the last store of a launch.
//
// What each piece restates (reference paths relative to phoenix_drone_simulation/):
//   motor()           envs/agents.py:259-298, envs/control.py:94-100, envs/utils.py:104-108
//   physics_simple()  envs/physics.py:130-200
//   physics_bullet()  envs/physics.py:91-124 + single-rigid-body stand-in for Bullet's
//                     stepSimulation (parity with real Bullet unpinned)
//   observe_*()       envs/{hover.py:131-163,circle.py:128-177,takeoff.py:107-149},
//                     envs/sensors.py:75-134, envs/utils.py:76-79
//   emit history      envs/base.py:303-319
//   reward/done/cost  envs/hover.py:89-129,169-187, circle.py:116-126,183-204,
//                     takeoff.py:96-105,155-174
//   reset_env()       envs/base.py:382-431, 239-296, agents.py:208-224,377-386 and the three
//                     task_specific_reset() (hover.py:192-243, circle.py:213-277,
//                     takeoff.py:179-212)
// Quirks that change numbers are kept and marked "quirk" (SURVEY.md A.6).
#pragma once
#include "pdx_layout.h"
#include "pdx_math.cuh"

namespace pdx {

// Device-side copy of PdxConfig in the arithmetic type (built once per launch on the host).
template <class T>
struct DevCfg {
  int control_mode, history, agg, obs_rate, use_latency, buf_size, use_motor_dynamics, reset_distribution,
      ground_effect, max_episode_steps, core_dim, obs_dim, dr_on, reset_on_nonfinite, auto_reset;
  int slots_obs_full, slots_obs_gyro, slots_reset_task, slots_reset_dr, slots_step, slots_reset;
  T dr, time_step, mass, inertia[3], arm, gravity, thrust2weight, max_thrust,
      k_mass_dr, ftf1, hover_x, hover_action, motor_tc, ou_theta, ou_sigma, lpf_ratio,
      pos_std, pos_unif, vel_std, quat_std, quat_unif, gyro_pi, gyro_sigma_b, gyro_rw, gyro_to, gyro_white,
      pen_action, pen_angle, pen_spin, pen_terminal, pen_velocity, arp, target[3], init_xyz[3],
      drag[3], prop_xy[4][2], prop_z, gec, prop_r, ge_hclip, lin_damp, ang_damp, ground_z;
};

template <class T> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<double> { typedef double4 type; };   // two 128-bit accesses

// quad index of (plane q, environment i): n < 2^31 (checked at the ABI), so the product is ONE 32 x 32 -> 64 bit
// multiply-add instead of the three instructions of a 64-bit product
__device__ __forceinline__ int64_t quad_index(int q, int64_t n, int64_t i) { return (int64_t)q * (int64_t)(int)n + i; }

template <class T>
__device__ __forceinline__ void load_quad(const T* base, int64_t n, int64_t i, int q, T* out) {
  if (sizeof(T) == 4) {
    const float4 v = reinterpret_cast<const float4*>(base)[quad_index(q, n, i)];
    out[0] = (T)v.x; out[1] = (T)v.y; out[2] = (T)v.z; out[3] = (T)v.w;
  } else {
    const double2* p = reinterpret_cast<const double2*>(base) + quad_index(q, n, i) * 2;
    const double2 a = p[0], b = p[1];
    out[0] = (T)a.x; out[1] = (T)a.y; out[2] = (T)b.x; out[3] = (T)b.y;
  }
}

template <class T>
__device__ __forceinline__ void store_quad(T* base, int64_t n, int64_t i, int q, const T* in) {
  if (sizeof(T) == 4) {
    reinterpret_cast<float4*>(base)[quad_index(q, n, i)] =
        make_float4((float)in[0], (float)in[1], (float)in[2], (float)in[3]);
  } else {
    double2* p = reinterpret_cast<double2*>(base) + quad_index(q, n, i) * 2;
    p[0] = make_double2((double)in[0], (double)in[1]);
    p[1] = make_double2((double)in[2], (double)in[3]);
  }
}

template <class T> __device__ __forceinline__ T clampT(T x, T lo, T hi) {
  return M<T>::fmin(M<T>::fmax(x, lo), hi);
}
template <class T> __device__ __forceinline__ T norm3(T a, T b, T c) {
  return M<T>::sqrt(a * a + b * b + c * c);
}

template <class T, int TASK, int PHYS, bool NOISE, int RNG, bool PID = false>
struct Model {
  static constexpr Layout L = make_layout(TASK, PHYS, NOISE, PID);
  static constexpr int NW = L.n_quads * 4;
  static constexpr int C = L.core_dim;
  static constexpr int E = C + 4;              // one history entry: observation + action
  static constexpr int QH = L.hist_quads;
  static constexpr bool BULLET = PHYS == PDX_PHYSICS_BULLET;

  const DevCfg<T>& c;
  T w[NW];                                     // persistent state words (registers)

  __device__ __forceinline__ explicit Model(const DevCfg<T>& cfg) : c(cfg) {}

  // ------------------------------------------------------------------------------------------
  __device__ __forceinline__ void load(const T* state, int64_t n, int64_t i) {
#pragma unroll
    for (int q = 0; q < L.n_quads; ++q) load_quad(state, n, i, q, &w[4 * q]);
  }
  // `all` = false: only the quads holding per-step words (constants change on reset only)
  __device__ __forceinline__ void store(T* state, int64_t n, int64_t i, bool all) const {
#pragma unroll
    for (int q = 0; q < L.n_quads; ++q)
      if (q < L.n_store_quads || all) store_quad(state, n, i, q, &w[4 * q]);
  }

  // reference trajectory point t (circle.py:46-56, takeoff.py:44-48)
  __device__ __forceinline__ void ref_point(int t, T ref[3]) const {
    if constexpr (TASK == PDX_TASK_CIRCLE) {
      T s, co;
      const T ang = T(2) * T(3.14159265358979323846) * T(t) / T(300);
      M<T>::sincos(ang, &s, &co);
      ref[0] = T(0.25) * (T(1) - co);
      ref[1] = T(0.25) * s;
      ref[2] = T(1);
    } else {
      ref[0] = T(0); ref[1] = T(0); ref[2] = T(t) / T(300);
    }
  }

  // body rates (`drone.rpy_dot`) and Euler angles of the current state
  __device__ __forceinline__ void body_rates(T om[3]) const {
    if constexpr (!BULLET) {
      om[0] = w[L.omega]; om[1] = w[L.omega + 1]; om[2] = w[L.omega + 2];
    } else {                                    // agents.py:452-453: R^T w_world
      T R[9];
      rot_from_quat(&w[L.quat], R);
      const T* ww = &w[L.omega_world];
      om[0] = R[0] * ww[0] + R[3] * ww[1] + R[6] * ww[2];
      om[1] = R[1] * ww[0] + R[4] * ww[1] + R[7] * ww[2];
      om[2] = R[2] * ww[0] + R[5] * ww[1] + R[8] * ww[2];
    }
  }
  __device__ __forceinline__ void euler(T e[3]) const {
    if constexpr (!BULLET) { e[0] = w[L.rpy]; e[1] = w[L.rpy + 1]; e[2] = w[L.rpy + 2]; }
    else euler_from_quat(&w[L.quat], e);        // agents.py:446
  }

  // ------------------------------------------------------------------------------------------
  //  motor model -> forces[4], z torque.   `substep` picks the OU draw site / tape slot.
  // ------------------------------------------------------------------------------------------
  __device__ __forceinline__ void motor(const T z[4], const float a[4], T f[4], T* tz) {
    T u[4];
    bool delayed = false;
    T d[4];                                      // action the controller sees, in T
    if constexpr (BULLET) {
      if (c.use_latency) {
        // agents.py:267-276: delayed action out of the ring, current action in.  The ring is
        // float64 in the reference, so the control stage runs in T here (quirk A.6-3).
        delayed = true;
        const int idx = (int)w[L.ring_idx];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          d[k] = idx == 0 ? w[L.ring + k] : w[L.ring + 4 + k];
          if (idx == 0) w[L.ring + k] = (T)a[k]; else w[L.ring + 4 + k] = (T)a[k];
        }
        w[L.ring_idx] = (T)((idx + 1) % c.buf_size);
      }
    }
    if constexpr (!PID) {
      if (delayed) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const T pwm = T(30000) + clampT(d[k], T(-1), T(1)) * T(30000);
          u[k] = pwm / T(60000);
        }
      } else {
        // control.py:98-99 evaluated in float32 because the policy's action is float32.
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float pwm = 30000.0f + fminf(fmaxf(a[k], -1.0f), 1.0f) * 30000.0f;
          u[k] = (T)(pwm / 60000.0f);
        }
      }
    } else {
      // control.py:120-287.  Leading arithmetic (clip, thrust, targets) in the dtype of the action
      // the controller receives: T out of the latency ring, float32 straight from the policy.
      const T pi = T(3.14159265358979323846);
      const bool attitude = c.control_mode == PDX_CTRL_ATTITUDE;
      T thrust, tgt[3];
      if (delayed) {
        const T c0 = clampT(d[0], T(-1), T(1));
        thrust = attitude ? T(45000) + c0 * T(10000) : T(30000) + c0 * T(30000);
#pragma unroll
        for (int k = 0; k < 3; ++k) tgt[k] = (clampT(d[1 + k], T(-1), T(1)) * pi) / (attitude ? T(18) : T(3));
      } else {
        const float c0 = fminf(fmaxf(a[0], -1.0f), 1.0f);
        thrust = (T)(attitude ? 45000.0f + c0 * 10000.0f : 30000.0f + c0 * 30000.0f);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          tgt[k] = (T)((fminf(fmaxf(a[1 + k], -1.0f), 1.0f) * 3.14159265358979323846f) / (attitude ? 18.0f : 3.0f));
      }
      const T dt = c.time_step;                  // controllers keep the nominal step (quirk A.6-13)
      if (attitude) {                            // outer loop: control.py:264-280
        const T kp[3] = {T(6), T(6), T(6)}, ki[3] = {T(3), T(3), T(1)}, kd[3] = {T(0), T(0), T(0.35)};
        const T lim[3] = {T(20), T(20), T(360)};
        T e[3];
        euler(e);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const T err = ((tgt[k] - e[k]) * T(180)) / pi;
          const T der = (err - w[L.pid + 9 + k]) / dt;
          w[L.pid + 9 + k] = err;
          w[L.pid + 6 + k] = clampT(w[L.pid + 6 + k] + err * dt, -lim[k], lim[k]);
          const T offs = (kp[k] * err + ki[k] * w[L.pid + 6 + k]) + kd[k] * der;
          tgt[k] = (offs / T(180)) * pi;
        }
      }
      // rate loop: control.py:160-180, firmware gains control.py:13-26
      const T kp[3] = {T(250), T(250), T(120)}, ki[3] = {T(500), T(500), T(16.7)}, kd[3] = {T(2.5), T(2.5), T(0)};
      const T lim[3] = {T(33.3), T(33.3), T(166.7)};
      T om[3], f[3];
      body_rates(om);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const T err = ((tgt[k] - om[k]) * T(180)) / pi;
        const T der = (err - w[L.pid + 3 + k]) / dt;
        w[L.pid + 3 + k] = err;
        w[L.pid + k] = clampT(w[L.pid + k] + err * dt, -lim[k], lim[k]);
        f[k] = (kp[k] * err + ki[k] * w[L.pid + k]) + kd[k] * der;
      }
      // mixer: control.py:34-50 (QUAD_FORMATION_X)
      const T r = f[0] / T(2), p = f[1] / T(2), y = f[2];
      const T pw[4] = {((thrust - r) - p) - y, ((thrust - r) + p) + y, ((thrust + r) + p) - y, ((thrust + r) - p) + y};
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = clampT(pw[k], T(0), T(60000)) / T(60000);
    }
    T tq[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      T& ou = w[L.ou + k];                     // envs/utils.py:104-108 (never reset)
      ou = ou + (c.ou_theta * (T(0) - ou) + c.ou_sigma * z[k]);
      T noisy = (T(1) + ou) * u[k];
      T K = c.max_thrust;
      if constexpr (BULLET) {
        if (c.use_motor_dynamics) {            // agents.py:284-288
          T& x = w[L.motor_x + k];
          const T B = w[L.motor_b + k];
          x = (T(1) - B) * x + B * M<T>::sqrt(u[k]);
          noisy = (T(1) + ou) * (x * x);
        }
        K = w[L.motor_k + k];
      }
      f[k] = K * clampT(noisy, T(0), T(1));
      tq[k] = w[L.ftf1] * f[k];                 // ftf0 cancels in the alternating sum
    }
    *tz = ((-tq[0] + tq[1]) - tq[2]) + tq[3];
  }

  // physics.py:27-58 applied per propeller (extension: never enabled by the reference)
  __device__ __forceinline__ void ground_effect(const T R[9], T f[4]) const {
    T e[3];
    euler(e);
    const T half_pi = T(1.5707963267948966);
    if (!(M<T>::fabs(e[0]) < half_pi && M<T>::fabs(e[1]) < half_pi)) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      T pz = w[L.xyz + 2] + R[6] * c.prop_xy[k][0] + R[7] * c.prop_xy[k][1] + R[8] * c.prop_z;
      pz = M<T>::fmax(pz, c.ge_hclip);
      const T r = c.prop_r / (T(4) * pz);
      f[k] = f[k] + f[k] * c.gec * (r * r);
    }
  }

  __device__ __forceinline__ void physics_simple(const T z_ou[4], const float a[4]) {
    T f[4], tz;
    motor(z_ou, a, f, &tz);
    T q[4], R[9];
    quat_from_euler(w[L.rpy], w[L.rpy + 1], w[L.rpy + 2], q);
    rot_from_quat(q, R);
    if (c.ground_effect) ground_effect(R, f);
    const T thrust = ((f[0] + f[1]) + f[2]) + f[3];
    const T m = w[L.mass], dt = w[L.dt];
    const T Fx = R[2] * thrust, Fy = R[5] * thrust, Fz = R[8] * thrust - c.gravity * m;
    const T sqrt2 = T(1.4142135623730951);
    const T tx = ((((-f[0] - f[1]) + f[2]) + f[3]) * c.arm) / sqrt2;
    const T ty = ((((-f[0] + f[1]) + f[2]) - f[3]) * c.arm) / sqrt2;
    T* om = &w[L.omega];
    const T* J = &w[L.inertia];
    const T Jw0 = J[0] * om[0], Jw1 = J[1] * om[1], Jw2 = J[2] * om[2];
    const T t0 = tx - (om[1] * Jw2 - om[2] * Jw1);
    const T t1 = ty - (om[2] * Jw0 - om[0] * Jw2);
    const T t2 = tz - (om[0] * Jw1 - om[1] * Jw0);
    T* v = &w[L.vel];
    v[0] += dt * (Fx / m); v[1] += dt * (Fy / m); v[2] += dt * (Fz / m);
    om[0] += dt * ((T(1) / J[0]) * t0);
    om[1] += dt * ((T(1) / J[1]) * t1);
    om[2] += dt * ((T(1) / J[2]) * t2);
    T* p = &w[L.xyz];
    p[0] += dt * v[0]; p[1] += dt * v[1]; p[2] += dt * v[2];
    w[L.rpy] += dt * om[0]; w[L.rpy + 1] += dt * om[1]; w[L.rpy + 2] += dt * om[2];
    p[2] = M<T>::fmax(p[2], T(0));             // physics.py:182
  }

  __device__ __forceinline__ void physics_bullet(const T z_ou[4], const float a[4]) {
    T f[4], tz;
    motor(z_ou, a, f, &tz);
    T R[9];
    rot_from_quat(&w[L.quat], R);
    T* v = &w[L.vel];
    T* ww = &w[L.omega_world];
    T* p = &w[L.xyz];
    const T* J = &w[L.inertia];
    const T m = w[L.mass], dt = w[L.dt];
    // drag: physics.py:106-115 (quirk: rotated by R here and once more by LINK_FRAME)
    T S = T(0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const T x = w[L.motor_x + k];
      S += T(2) * T(3.14159265358979323846) * (x * x * T(25000)) / T(60);
    }
    const T kv0 = (T(-1) * c.drag[0] * S) * v[0], kv1 = (T(-1) * c.drag[1] * S) * v[1],
            kv2 = (T(-1) * c.drag[2] * S) * v[2];
    if (c.ground_effect) ground_effect(R, f);
    T fb[3], tb[3];
    fb[0] = R[0] * kv0 + R[1] * kv1 + R[2] * kv2;
    fb[1] = R[3] * kv0 + R[4] * kv1 + R[5] * kv2;
    fb[2] = R[6] * kv0 + R[7] * kv1 + R[8] * kv2 + (((f[0] + f[1]) + f[2]) + f[3]);
    tb[0] = c.prop_xy[0][1] * f[0] + c.prop_xy[1][1] * f[1] + c.prop_xy[2][1] * f[2] + c.prop_xy[3][1] * f[3];
    tb[1] = -(c.prop_xy[0][0] * f[0] + c.prop_xy[1][0] * f[1] + c.prop_xy[2][0] * f[2] + c.prop_xy[3][0] * f[3]);
    tb[2] = tz;
    T vb[3], wb[3];
    vb[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    vb[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    vb[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    wb[0] = R[0] * ww[0] + R[3] * ww[1] + R[6] * ww[2];
    wb[1] = R[1] * ww[0] + R[4] * ww[1] + R[7] * ww[2];
    wb[2] = R[2] * ww[0] + R[5] * ww[1] + R[8] * ww[2];
    const T gz = -c.gravity * m;                // gravity in the body frame: R^T (0,0,-g m)
    fb[0] += R[6] * gz; fb[1] += R[7] * gz; fb[2] += R[8] * gz;
    const T Jw0 = J[0] * wb[0], Jw1 = J[1] * wb[1], Jw2 = J[2] * wb[2];
    const T ld = c.lin_damp + c.lin_damp * norm3(vb[0], vb[1], vb[2]);
    const T ad = c.ang_damp + c.ang_damp * norm3(wb[0], wb[1], wb[2]);
    fb[0] -= m * vb[0] * ld; fb[1] -= m * vb[1] * ld; fb[2] -= m * vb[2] * ld;
    tb[0] -= Jw0 * ad; tb[1] -= Jw1 * ad; tb[2] -= Jw2 * ad;
    tb[0] -= wb[1] * Jw2 - wb[2] * Jw1;
    tb[1] -= wb[2] * Jw0 - wb[0] * Jw2;
    tb[2] -= wb[0] * Jw1 - wb[1] * Jw0;
    const T ab0 = fb[0] / m, ab1 = fb[1] / m, ab2 = fb[2] / m;
    const T al0 = tb[0] / J[0], al1 = tb[1] / J[1], al2 = tb[2] / J[2];
    v[0] += dt * (R[0] * ab0 + R[1] * ab1 + R[2] * ab2);
    v[1] += dt * (R[3] * ab0 + R[4] * ab1 + R[5] * ab2);
    v[2] += dt * (R[6] * ab0 + R[7] * ab1 + R[8] * ab2);
    ww[0] += dt * (R[0] * al0 + R[1] * al1 + R[2] * al2);
    ww[1] += dt * (R[3] * al0 + R[4] * al1 + R[5] * al2);
    ww[2] += dt * (R[6] * al0 + R[7] * al1 + R[8] * al2);
    p[0] += dt * v[0]; p[1] += dt * v[1]; p[2] += dt * v[2];
    const T wn = norm3(ww[0], ww[1], ww[2]);
    if (wn * dt > T(1e-12)) {                   // exponential map of the world rate
      T s, co;
      M<T>::sincos(T(0.5) * wn * dt, &s, &co);
      const T ax = ww[0] / wn * s, ay = ww[1] / wn * s, az = ww[2] / wn * s, aw = co;
      T* q = &w[L.quat];
      const T bx = q[0], by = q[1], bz = q[2], bw = q[3];
      const T nx = aw * bx + ax * bw + ay * bz - az * by;
      const T ny = aw * by - ax * bz + ay * bw + az * bx;
      const T nz = aw * bz + ax * by - ay * bx + az * bw;
      const T nw = aw * bw - ax * bx - ay * by - az * bz;
      const T nn = M<T>::sqrt(nx * nx + ny * ny + nz * nz + nw * nw);
      q[0] = nx / nn; q[1] = ny / nn; q[2] = nz / nn; q[3] = nw / nn;
    }
    if (p[2] < c.ground_z) {                    // crude ground plane (not Bullet's contact solver)
      p[2] = c.ground_z;
      v[2] = M<T>::fmax(v[2], T(0));
    }
  }

  // ------------------------------------------------------------------------------------------
  //  observation
  // ------------------------------------------------------------------------------------------
  // sensors.py:121-134 + envs/utils.py:76-79: gyro bias random walk, white noise, low pass.
  // n[0..2] bias, n[3..5] random walk, n[6..8] turn-on draws.
  __device__ __forceinline__ void gyro_update(const T n[9], const T om[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      T& b = w[L.gyro_bias + k];
      b = c.gyro_pi * b + c.gyro_sigma_b * n[k];
      const T noisy = ((om[k] + b) + c.gyro_rw * n[3 + k]) + c.gyro_to * n[6 + k];
      T& lp = w[L.gyro_lpf + k];
      lp = (T(1) - c.lpf_ratio) * lp + (T(1) * c.lpf_ratio) * noisy;
    }
  }

  // Production draws: the two independent white-noise terms of the gyro model
  // (0.0105 N + 5deg N, sensors.py:130-131) are one normal with the combined standard deviation
  // gyro_white = sqrt(rw^2 + to^2) -- the same distribution with one draw instead of two.
  __device__ __forceinline__ void gyro_update_merged(const T nb[3], const T nw[3], const T om[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      T& b = w[L.gyro_bias + k];
      b = c.gyro_pi * b + c.gyro_sigma_b * nb[k];
      const T noisy = (om[k] + b) + c.gyro_white * nw[k];
      T& lp = w[L.gyro_lpf + k];
      lp = (T(1) - c.lpf_ratio) * lp + (T(1) * c.lpf_ratio) * noisy;
    }
  }
  // ... and what the oracle has to be told when it replays a launch (pdx_dump_draws): two
  // reference draws N2, N3 with rw N2 + to N3 == gyro_white * nw.
  template <class RG>
  __device__ __forceinline__ void dump_gyro(const RG& rng, int slot, const T nb[3], const T nw[3]) const {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      rng.dump_value(slot + k, (double)nb[k]);
      rng.dump_value(slot + 3 + k, (double)nw[k] * (double)c.gyro_rw / (double)c.gyro_white);
      rng.dump_value(slot + 6 + k, (double)nw[k] * (double)c.gyro_to / (double)c.gyro_white);
    }
  }

  // One physics sub-step plus the (discarded) observation call that follows it in the reference
  // (base.py:461-464): the OU draw and the gyro draws of that call come from one draw block.
  // `slot`: first tape slot of the sub-step; `full`: the observation call is a full one.
  __device__ __forceinline__ void substep(const Rng<T, RNG>& rng, const float a[4], int s, int slot, bool full) {
    if constexpr (NOISE) {
      const int g0 = 4 + (full ? 12 : 0);
      T om[3];
      if (rng.exact_draws()) {
        const int rel[13] = {0, 1, 2, 3, g0, g0 + 1, g0 + 2, g0 + 3, g0 + 4, g0 + 5, g0 + 6, g0 + 7, g0 + 8};
        T z[13];
        rng.template normals_at<13>(SITE_SUBSTEP + 8 * s, slot, rel, z);
        if constexpr (BULLET) physics_bullet(z, a); else physics_simple(z, a);
        body_rates(om);
        gyro_update(&z[4], om);
      } else {
        T z[10];
        rng.template gen_normals<10>(SITE_SUBSTEP + 8 * s, z);
#pragma unroll
        for (int k = 0; k < 4; ++k) rng.dump_value(slot + k, (double)z[k]);
        dump_gyro(rng, slot + g0, &z[4], &z[7]);
        if constexpr (BULLET) physics_bullet(z, a); else physics_simple(z, a);
        body_rates(om);
        gyro_update_merged(&z[4], &z[7], om);
      }
    } else {
      T z[4];
      rng.template normals<4>(SITE_SUBSTEP + 8 * s, slot, z);
      if constexpr (BULLET) physics_bullet(z, a); else physics_simple(z, a);
    }
  }

  // One full compute_observation() -> core[C].  `site0`/`slot`: first draw site / tape slot.
  // `q_true`: quaternion to report when noise is off (drone.quaternion).
  template <class R>
  __device__ __forceinline__ void observe(const R& rng, uint32_t site0, int slot,
                                          const T target[3], const T act[4], const T q_true[4],
                                          T core[C]) {
    const T* p = &w[L.xyz];
    const T* v = &w[L.vel];
    T om[3];
    body_rates(om);
    T px[3];
    if constexpr (NOISE) {
      // tape slots of one call (sensors.py:84-118): 0-2 pos n, 3-5 pos u, 6-8 vel n, 9-11 vel u
      // (zero range, unused), 12-20 gyro n, 21-23 theta n, 24-26 theta u, 27-32 accelerometer.
      T pn[3], vn[3], tn[3], zu[6], e[3], q[4];
      if (rng.exact_draws()) {
        const int reln[18] = {0, 1, 2, 6, 7, 8, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23};
        const int relu[6] = {3, 4, 5, 24, 25, 26};
        T zn[18];
        rng.template normals_at<18>(site0, slot, reln, zn);
        rng.template uniforms_at<6>(site0 + 4, slot, relu, zu);
        gyro_update(&zn[6], om);
#pragma unroll
        for (int k = 0; k < 3; ++k) { pn[k] = zn[k]; vn[k] = zn[3 + k]; tn[k] = zn[15 + k]; }
      } else {
        T zn[15];
        rng.template gen_normals<15>(site0, zn);
        rng.gen_uniforms6(site0 + 4, zu);
        gyro_update_merged(&zn[6], &zn[9], om);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          pn[k] = zn[k]; vn[k] = zn[3 + k]; tn[k] = zn[12 + k];
          rng.dump_value(slot + k, (double)pn[k]); rng.dump_value(slot + 3 + k, (double)zu[k]);
          rng.dump_value(slot + 6 + k, (double)vn[k]); rng.dump_value(slot + 21 + k, (double)tn[k]);
          rng.dump_value(slot + 24 + k, (double)zu[3 + k]);
        }
        dump_gyro(rng, slot + 12, &zn[6], &zn[9]);
      }
      euler(e);
      const T pi = T(3.14159265358979323846);
      const T lo[3] = {-pi, -pi / T(2), -pi}, hi[3] = {pi, pi / T(2), pi};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        px[k] = p[k] + (c.pos_std * pn[k] + (-c.pos_unif + (c.pos_unif - (-c.pos_unif)) * zu[k]));
        core[7 + k] = v[k] + c.vel_std * vn[k];
        const T th = c.quat_std * tn[k] + (-c.quat_unif + (c.quat_unif - (-c.quat_unif)) * zu[3 + k]);
        e[k] = clampT(e[k] + th, lo[k], hi[k]);
        core[10 + k] = w[L.gyro_lpf + k];
      }
      quat_from_euler(e[0], e[1], e[2], q);
#pragma unroll
      for (int k = 0; k < 4; ++k) core[3 + k] = q[k];
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) { px[k] = p[k]; core[7 + k] = v[k]; core[10 + k] = om[k]; }
#pragma unroll
      for (int k = 0; k < 4; ++k) core[3 + k] = q_true[k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) core[k] = px[k];
    int o = 13;
    if constexpr (TASK == PDX_TASK_TAKEOFF || (TASK == PDX_TASK_HOVER && !NOISE)) {
#pragma unroll
      for (int k = 0; k < 4; ++k) core[o + k] = act[k];
      o += 4;
    }
    if constexpr (TASK != PDX_TASK_HOVER) {
#pragma unroll
      for (int k = 0; k < 3; ++k) core[o + k] = target[k] - px[k];
    }
  }

  // ------------------------------------------------------------------------------------------
  //  done / cost
  // ------------------------------------------------------------------------------------------
  __device__ __forceinline__ bool done(const T e[3], const T om[3], const T target[3]) const {
    if constexpr (TASK == PDX_TASK_HOVER) {               // hover.py:89-101
      const T pi = T(3.14159265358979323846);
      const T d = pi * T(60) / T(180);
      bool r = w[L.xyz + 2] < T(0.2);
      r = r || M<T>::fabs(e[0]) > d || M<T>::fabs(e[1]) > d;
#pragma unroll
      for (int k = 0; k < 3; ++k) r = r || (T(180) * M<T>::fabs(om[k]) / pi > T(300));
      return r;
    }
    if constexpr (TASK == PDX_TASK_CIRCLE) {              // circle.py:116-120
      return norm3(w[L.xyz] - target[0], w[L.xyz + 1] - target[1], w[L.xyz + 2] - target[2]) > T(0.25);
    }
    return false;                               // takeoff.py:100 (quirk A.6-2)
  }

  __device__ __forceinline__ T cost(const T e[3], const T om[3], const float a[4]) const {
    if constexpr (TASK != PDX_TASK_HOVER) return T(0);    // circle.py:122-126, takeoff.py:102-105
    const T pi = T(3.14159265358979323846);
    bool cst = M<T>::fabs(w[L.xyz]) > T(0.10) || M<T>::fabs(w[L.xyz + 1]) > T(0.10) || w[L.xyz + 2] > T(1.20);
    const T rp = pi * T(10) / T(180);
    cst = cst || M<T>::fabs(e[0]) > rp || M<T>::fabs(e[1]) > rp;
    // quirk (hover.py:118-124): state[10:13] are the body rates (not the velocity) and
    // state[13:16] the first three action components (not the rates).
    const T rl = pi * T(200) / T(180);
#pragma unroll
    for (int k = 0; k < 3; ++k) cst = cst || M<T>::fabs(om[k]) > T(0.25) || M<T>::fabs((T)a[k]) > rl;
    return cst ? T(1) : T(0);
  }
};

}  // namespace pdx
