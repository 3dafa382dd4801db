"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the rollout-collector arithmetic.

Restates, for the parity tests of pdx_gae / pdx_moments / rollout.OnlineMeanStd / EpisodeStats
(paths relative to /root/reference/phoenix_drone_simulation/):
  discount_cumsum           algs/core.py:105-119
  finish_path               algs/core.py:497-534 (+ calculate_adv_and_value_targets :458-479,
                            reward scaling through OnlineMeanStd.forward, online_mean_std.py:32-48)
  online_mean_std_update    utils/online_mean_std.py:50-95 with P ranks emulated in one process
                            (mpi_avg_torch_tensor, mpi_tools.py:199-214 = mean over ranks)
  statistics_scalar         utils/mpi_tools.py:217-240
Pinned against the unmodified reference by oracle/gen_golden_collector.py ->
tests/golden_collector/*.npz -> tests/test_collector_oracle.py.  Only tests/ may import this file.
"""
import numpy as np


def discount_cumsum(x, discount):
    """y[t] = x[t] + discount * y[t+1] (core.py:105-119 does it with lfilter on the reversed
    vector; the recursion is the same)."""
    x = np.asarray(x)
    y = np.zeros_like(x)
    acc = x.dtype.type(0)
    d = x.dtype.type(discount)
    for t in range(len(x) - 1, -1, -1):
        acc = x[t] + d * acc
        y[t] = acc
    return y


def finish_path(rew, val, last_val, gamma=0.99, lam=0.95, ret_std=None, eps=1e-5, bound=10.0):
    """One trajectory (core.py:497-534).  Returns (adv, target_v, discounted_ret).
    ret_std: std of ret_oms when reward scaling is on (the bootstrap value appended to `rews` is
    scaled as well, but only rews[:-1] enters the deltas)."""
    rews = np.append(np.asarray(rew, dtype=np.float32), np.float32(last_val))
    vals = np.append(np.asarray(val, dtype=np.float32), np.float32(last_val))
    disc_ret = discount_cumsum(rews, gamma)[:-1]
    if ret_std is not None:
        rews = np.clip(rews / (np.float32(ret_std) + np.float32(eps)), -bound, bound)
    deltas = rews[:-1] + np.float32(gamma) * vals[1:] - vals[:-1]
    adv = discount_cumsum(deltas, gamma * lam)
    return adv, adv + vals[:-1], disc_ret


def rollout_gae(rew, val, done, boot_val, last_val, gamma=0.99, lam=0.95, ret_std=None):
    """[T, N] lock-step rollout cut into trajectories per env column, each fed to finish_path
    exactly as roll_out does (iwpg.py:371-385): done==1 -> last_val 0, done==2 -> boot_val[t],
    end of the rollout -> last_val[i]."""
    T, N = rew.shape
    adv = np.zeros((T, N), np.float32)
    tv = np.zeros((T, N), np.float32)
    dr = np.zeros((T, N), np.float32)
    for i in range(N):
        start = 0
        for t in range(T):
            end_of_path = done[t, i] != 0 or t == T - 1
            if not end_of_path:
                continue
            if done[t, i] == 1:
                lv = 0.0
            elif done[t, i] == 2:
                lv = boot_val[t, i]
            else:
                lv = last_val[i]
            a, v, d = finish_path(rew[start:t + 1, i], val[start:t + 1, i], lv, gamma, lam, ret_std)
            adv[start:t + 1, i], tv[start:t + 1, i], dr[start:t + 1, i] = a, v, d
            start = t + 1
    return adv, tv, dr


class OnlineMeanStdOracle:
    """online_mean_std.py:6-95 with `per_rank_batches` = the list of batches the P ranks hold."""

    def __init__(self, dim, eps=1e-5):
        self.mean = np.zeros(dim, np.float32)
        self.std = np.ones(dim, np.float32)
        self.count = np.zeros(1, np.float32)
        self.eps = eps

    def update(self, per_rank_batches):
        P = len(per_rank_batches)
        xs = [np.asarray(x, np.float32).reshape(-1, self.mean.shape[0]) for x in per_rank_batches]
        n_B = np.float32(xs[0].shape[0] * P)
        n_A = self.count.copy()
        n_AB = self.count + n_B
        batch_mean = np.mean([x.mean(axis=0, dtype=np.float32) for x in xs], axis=0, dtype=np.float32)
        delta = batch_mean - self.mean
        mean_new = self.mean + delta * n_B / n_AB
        batch_var = np.mean([((x - mean_new) ** 2).mean(axis=0, dtype=np.float32) for x in xs], axis=0,
                            dtype=np.float32)
        M2 = n_A * self.std ** 2 + n_B * batch_var + delta ** 2 * (n_A * n_B / n_AB)
        self.mean, self.count = mean_new.astype(np.float32), n_AB
        self.std = np.sqrt(M2 / n_AB).astype(np.float32)

    def forward(self, x, subtract_mean=True, clip=False):
        y = (x - self.mean) / (self.std + self.eps) if subtract_mean else x / (self.std + self.eps)
        return np.clip(y, -10, 10) if clip else y


def statistics_scalar(per_rank_values):
    """mpi_statistics_scalar(with_min_and_max=True) over P ranks."""
    xs = [np.asarray(x, np.float32) for x in per_rank_values]
    n = sum(len(x) for x in xs)
    mean = sum(float(np.sum(x)) for x in xs) / n
    std = np.sqrt(sum(float(np.sum((x - mean) ** 2)) for x in xs) / n)
    mn = min(float(np.min(x)) if len(x) else np.inf for x in xs)
    mx = max(float(np.max(x)) if len(x) else -np.inf for x in xs)
    return mean, std, mn, mx


# ---------------------------------------------------------------------------------------------
#  ActorCritic.step (algs/core.py:370-393), deterministic part, float64
# ---------------------------------------------------------------------------------------------
def mlp_forward(x, layers, activation):
    """build_mlp_network (utils/utils.py:56-110) / core.build_mlp_network: Linear layers with `activation`
    between them and identity after the last.  layers: [(W [out][in], b [out]), ...]."""
    act = {'relu': lambda z: np.maximum(z, 0.0), 'tanh': np.tanh}[activation]
    h = np.asarray(x, dtype=np.float64)
    for k, (w, b) in enumerate(layers):
        h = h @ np.asarray(w, dtype=np.float64).T + np.asarray(b, dtype=np.float64)
        if k + 1 < len(layers):
            h = act(h)
    return h


def actor_critic_step(obs, mean, std, pi_layers, v_layers, log_std, eps_draw, norm_eps=1e-5):
    """ActorCritic.step for a batch: standardise (utils/online_mean_std.py:42-48), v = critic(o) (tanh MLP,
    core.py:297-310), mu = actor(o) (relu MLP, core.py:227-289), a = mu + exp(log_std) * eps (core.py:282-289
    with the N(0,1) draw supplied), log p = sum log N(a; mu, std) (core.py:254-262).
    Returns (action, value, logp, mu)."""
    o = (np.asarray(obs, dtype=np.float64) - np.asarray(mean, dtype=np.float64)) / (np.asarray(std, dtype=np.float64) + norm_eps)
    mu = mlp_forward(o, pi_layers, 'relu')
    v = mlp_forward(o, v_layers, 'tanh')[:, 0] if v_layers is not None else None
    sd = np.exp(np.asarray(log_std, dtype=np.float64))
    a = mu + sd * eps_draw
    logp = (-0.5 * ((a - mu) / sd) ** 2 - np.log(sd) - 0.5 * np.log(2.0 * np.pi)).sum(-1)
    return a, v, logp, mu


def load_policy_json_layers(path):
    """The JSON policy format (utils/utils.py:362-430 dump_network_json): scaling parameters and dense layers."""
    import json
    with open(path) as f:
        data = json.load(f)
    layers = []
    while str(len(layers)) in data:
        e = data[str(len(layers))]
        layers.append((np.asarray(e['weights'], dtype=np.float64), np.asarray(e['biases'], dtype=np.float64).reshape(-1)))
    sp = np.asarray(data['scaling_parameters'], dtype=np.float64)
    return sp[0], sp[1], layers, data.get('activation', 'relu')
