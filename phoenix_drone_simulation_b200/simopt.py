"""Simulation optimisation objective on the batched engine.

The reference fits simulation parameters (thrust-to-weight ratio, motor time constant, latency)
to real-flight logs by replaying mini-trajectories one at a time and one candidate at a time
(paths relative to phoenix_drone_simulation/):
  simopt/pybullet.py:72-128   ObjectiveFunctionPyBullet.evaluate (loop over mini-trajectories)
  simopt/pybullet.py:130-183  evaluate_once: pre-steps for the motor state, set the logged initial
                              state, replay T-1 logged actions, discounted mean of the loss
  simopt/pybullet.py:195-225  loss_function: L1 + L2 norm of (rpy error, 100 xyz error,
                              10 velocity error, body-rate error)
  simopt/pybullet.py:232-247  set_parameters -> agents.py:208-224 update_motor_dynamics
  simopt/core.py:47-80        mini-trajectory slicing (T = 35, 5 pre-steps)
Here every (candidate, mini-trajectory) pair is one environment of a VecEnv: K candidates x M
mini-trajectories are evaluated with one reset, `pre_steps` single-step launches and ONE fused
launch.  Per-environment parameters are the motor words of the state (`motor_b` = Ts / T,
`motor_k` = 0.028 g t2w / 4, agents.py:222-224).

Latency as a parameter (third candidate column; simopt/pybullet.py:248 -> agents.py:388-404 `set_latency`): the
reference's ring holds L = int(latency / TIME_STEP) sub-steps (none below one sub-step).  The objective replays
KNOWN action sequences, so a ring of L sub-steps is a ring of r in {1, 2} sub-steps (what the engine's registers hold)
behind a sequence delayed by m whole steps, L = 2 m + r (two sub-steps per step): the controller of sub-step s sees
the action of step (s - L) // 2 either way, and zeros -- the cleared ring -- before that.  Candidates are grouped by r
(0 = no latency); each group runs on its own VecEnv with its own, per-environment delayed action tensor.
"""
import math

import torch

from .vec_env import VecEnv


def euler_from_quat(q):
    """pybullet.getEulerFromQuaternion, batched ([..., 4] (x, y, z, w) -> [..., 3])."""
    x, y, z, w = q.unbind(-1)
    sarg = -2.0 * (x * z - w * y)
    roll = torch.atan2(2.0 * (y * z + w * x), w * w - x * x - y * y + z * z)
    pitch = torch.asin(sarg.clamp(-1.0, 1.0))
    yaw = torch.atan2(2.0 * (x * y + w * z), w * w + x * x - y * y - z * z)
    lo, hi = sarg <= -0.99999, sarg >= 0.99999
    roll = torch.where(lo | hi, torch.zeros_like(roll), roll)
    pitch = torch.where(lo, torch.full_like(pitch, -0.5 * math.pi), torch.where(hi, torch.full_like(pitch, 0.5 * math.pi), pitch))
    yaw = torch.where(lo, 2.0 * torch.atan2(x, -y), torch.where(hi, 2.0 * torch.atan2(-x, y), yaw))
    return torch.stack([roll, pitch, yaw], dim=-1)


def quat_from_euler(rpy):
    """pybullet.getQuaternionFromEuler, batched."""
    h = rpy * 0.5
    sr, cr, sp, cp, sy, cy = torch.sin(h[..., 0]), torch.cos(h[..., 0]), torch.sin(h[..., 1]), torch.cos(h[..., 1]), \
        torch.sin(h[..., 2]), torch.cos(h[..., 2])
    q = torch.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy], dim=-1)
    return q / q.norm(dim=-1, keepdim=True)


def rot_from_quat(q):
    x, y, z, w = q.unbind(-1)
    s = 2.0 / (q * q).sum(-1)
    R = torch.stack([1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y),
                     s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x),
                     s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)], dim=-1)
    return R.reshape(*q.shape[:-1], 3, 3)


class TrajectoryObjective:
    """observations [M, T, 12] = (xyz, xyz_dot, rpy, body rates) logs, actions [M, T, 4] in [-1, 1],
    pre_inputs [M, P, 4]; `evaluate(candidates [K, 2] = (thrust-to-weight, motor time constant))` or
    `[K, 3] = (..., latency in seconds)` returns the K objective values of simopt/pybullet.py:72-128;
    `evaluate(..., per_trajectory=True)` the [K, M] losses of evaluate_once (pybullet.py:130-183) themselves.
    Without a latency column the latency of the environment's configuration applies to every candidate."""

    def __init__(self, observations, actions, pre_inputs, env_id='DroneHoverBulletEnv-v0', device='cuda',
                 dtype=torch.float64, gamma=0.95, seed=0, **env_kwargs):
        self.obs = torch.as_tensor(observations, dtype=torch.float64, device=device)
        self.act = torch.as_tensor(actions, dtype=torch.float32, device=device)
        self.pre = torch.as_tensor(pre_inputs, dtype=torch.float32, device=device)
        self.M, self.T = self.obs.shape[:2]
        # pybullet.py:263-265; evaluate_once switches the reset distribution off for good (pybullet.py:157), so
        # the pre-steps start from a drone at rest with motor state 0
        kw = dict(domain_randomization=-1, observation_noise=-1, auto_reset=False, enable_reset_distribution=False)
        kw.update(env_kwargs)
        self.env_id, self.device, self.dtype, self.gamma, self.seed, self.kw = env_id, device, dtype, gamma, seed, kw
        self._envs = {}

    def _make(self, n, ring=None):
        """VecEnv of n environments; ring = None: the configured latency, 0 / 1 / 2: latency ring of that many sub-steps."""
        env = self._envs.get(ring)
        if env is None or env.num_envs != n:
            kw = dict(self.kw)
            if ring is not None:
                kw['latency'] = {0: 0.0, 1: 0.005, 2: 0.0125}[ring]      # agents.py:165,180 with TIME_STEP = 1 / 200
            env = self._envs[ring] = VecEnv(self.env_id, n, device=self.device, dtype=self.dtype, seed=self.seed, **kw)
            assert env.cfg.physics == 'PyBulletPhysics', 'the objective fits the motor model of the Bullet ids'
            if ring is not None:
                assert env.pdx.agg == 2 and abs(env.pdx.time_step - 0.005) < 1e-12, 'latency candidates: 200 Hz sub-steps, two per step'
                assert (env.pdx.use_latency, env.pdx.buf_size if ring else 1) == (int(ring > 0), max(ring, 1))
        return env

    @staticmethod
    def ring_length(latency, time_step=0.005):
        """Sub-steps of delay of `set_latency` (agents.py:388-404): none below one sub-step, else int(latency / TIME_STEP)
        -- true division, unlike the constructor's floor division (0.015 s: 3 here, 2 there)."""
        return 0 if latency < time_step else int(latency / time_step)

    @torch.no_grad()
    def evaluate(self, candidates, per_trajectory=False):
        cand = torch.as_tensor(candidates, dtype=torch.float64, device=self.device)
        cand = cand.reshape(-1, cand.shape[-1])
        if cand.shape[1] == 2:
            return self._evaluate_group(cand, None, None, per_trajectory)
        assert cand.shape[1] == 3, 'candidates: (thrust-to-weight, motor time constant[, latency])'
        L = [self.ring_length(float(x)) for x in cand[:, 2].tolist()]
        res = torch.empty((cand.shape[0], self.M), dtype=torch.float64, device=self.device)
        for ring in (0, 1, 2):                                   # L = 2 m + r
            ks = [k for k, l in enumerate(L) if (l == 0 and ring == 0) or (l > 0 and 2 - l % 2 == ring)]
            if ks:
                shift = torch.tensor([(L[k] - ring) // 2 for k in ks], device=self.device)
                res[ks] = self._evaluate_group(cand[ks, :2], ring, shift, True)
        return res if per_trajectory else res.mean(1)

    def _delayed(self, seq, shift):
        """seq [M, S, 4] -> [K, M, S, 4]: candidate k sees seq delayed by shift[k] steps, zeros (the cleared ring) first."""
        K, S = shift.shape[0], seq.shape[1]
        t = torch.arange(S, device=self.device)[None, :] - shift[:, None]                  # [K, S]
        out = seq[:, t.clamp_min(0)]                                                       # [M, K, S, 4]
        return (out * (t >= 0)[None, :, :, None]).transpose(0, 1)

    def _evaluate_group(self, cand, ring, shift, per_trajectory):
        K, M, T = cand.shape[0], self.M, self.T
        env = self._make(K * M, ring)
        rep = lambda x: x.unsqueeze(0).expand(K, *x.shape).reshape(K * M, *x.shape[1:])
        if shift is None:
            pre_seq, act_seq = rep(self.pre), rep(self.act[:, :T - 1])
        else:
            pre_seq = self._delayed(self.pre, shift).reshape(K * M, *self.pre.shape[1:])
            act_seq = self._delayed(self.act[:, :T - 1], shift).reshape(K * M, T - 1, 4)
        # 1) reset, candidate parameters (update_motor_dynamics), pre-steps for the motor state
        env.reset()
        ts = env.pdx.time_step
        t2w, tau = cand[:, 0].clamp_min(0), cand[:, 1].clamp_min(ts)
        env.set_state('motor_b', (ts / tau).repeat_interleave(M)[:, None].expand(-1, 4))
        env.set_state('motor_k', (0.028 * 9.81 * t2w / 4).repeat_interleave(M)[:, None].expand(-1, 4))
        for j in range(self.pre.shape[1]):
            env.step(pre_seq[:, j].contiguous())
        # 2) logged initial state: pose, velocities, cleared latency ring.  quirk kept (verified against the reference,
        # tests/golden_collector/simopt_hover.npz): evaluate_once stores R w_logged as init_rpy_dot
        # (pybullet.py:153-154) and task_specific_reset hands R^T init_rpy_dot to Bullet as the WORLD rate
        # (hover.py:242) -- the logged body rates end up as the world angular velocity, unrotated
        x0 = rep(self.obs[:, 0])
        q = quat_from_euler(x0[:, 6:9])
        env.set_state('xyz', x0[:, 0:3])
        env.set_state('vel', x0[:, 3:6])
        env.set_state('quat', q)
        env.set_state('omega_world', x0[:, 9:12])
        for name, width in (('ring', 8), ('ring_idx', 1), ('last_action', 4), ('ep_length', 1), ('ep_return', 1)):
            env.set_state(name, torch.zeros((K * M, width), dtype=torch.float64, device=self.device))
        # 3) replay the logged actions in one fused launch
        n, d = K * M, env.obs_dim
        acts = act_seq.transpose(0, 1).contiguous()                                        # [T-1, n, 4]
        out = {'obs': torch.empty((T - 1, n, d), dtype=self.dtype, device=self.device),
               'reward': torch.empty((T - 1, n), dtype=self.dtype, device=self.device),
               'cost': torch.empty((T - 1, n), dtype=self.dtype, device=self.device),
               'terminated': torch.empty((T - 1, n), dtype=torch.uint8, device=self.device),
               'truncated': torch.empty((T - 1, n), dtype=torch.uint8, device=self.device)}
        env.step_many(acts, out)
        e0 = (env.cfg.observation_history_size - 1) * (env.core_dim + 4)
        sim = out['obs'][:, :, e0:e0 + 13].double()                                        # newest entry: xyz, quat, vel, rates
        real = rep(self.obs[:, 1:]).transpose(0, 1)                                        # [T-1, n, 12]
        # 4) loss_function, discounted mean over the mini-trajectory, mean over mini-trajectories
        err = torch.cat([euler_from_quat(sim[..., 3:7]) - real[..., 6:9], 100.0 * (sim[..., 0:3] - real[..., 0:3]),
                         10.0 * (sim[..., 7:10] - real[..., 3:6]), sim[..., 10:13] - real[..., 9:12]], dim=-1)
        L = err.abs().sum(-1) + err.norm(dim=-1)
        w = self.gamma ** torch.arange(T - 1, dtype=torch.float64, device=self.device)
        per_traj = (L * w[:, None]).mean(0)                                                # np.mean(errs)
        per_traj = per_traj.reshape(K, M)
        return per_traj if per_trajectory else per_traj.mean(1)
