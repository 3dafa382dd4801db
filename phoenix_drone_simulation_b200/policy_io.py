"""Policy JSON import / export and batched evaluation.

Format and semantics follow the reference (paths relative to phoenix_drone_simulation/):
  utils/utils.py:362-430   dump_network_json: {"scaling_parameters": [mean[D], std[D]], "activation",
                           "0"/"1"/...: {"type": "standard", "weights": [out][in], "biases": [out]}}
  utils/utils.py:56-110    build_mlp_network(data): the loader used by load_network_json (:309-340)
  utils/evaluation.py:52-107  EnvironmentEvaluator: deterministic policy, one episode per evaluation,
                           returns / episode lengths / costs
Only dense ("standard") layers are supported; the CSR product layers of the firmware export are not.
"""
import json

import torch

from .rollout import ActorCritic
from .vec_env import VecEnv


def export_policy_json(ac, path, activation='relu'):
    """Actor network + observation scaling of an ActorCritic -> JSON file (dump_network_json)."""
    lins = [m for m in ac.pi if isinstance(m, torch.nn.Linear)]
    d = lins[0].in_features
    mean = ac.obs_oms.mean if ac.obs_oms is not None else torch.zeros(d)
    std = ac.obs_oms.std if ac.obs_oms is not None else torch.ones(d)
    data = {'scaling_parameters': [mean.detach().cpu().double().tolist(), std.detach().cpu().double().tolist()],
            'activation': activation}
    for i, lin in enumerate(lins):
        data[str(i)] = {'type': 'standard', 'weights': lin.weight.detach().cpu().double().tolist(),
                        'biases': lin.bias.detach().cpu().double().tolist()}
    with open(path, 'w') as f:
        json.dump(data, f)
    return data


def load_policy_json(path, device='cuda', standardize=True, **ac_kwargs):
    """JSON file -> ActorCritic whose actor and observation scaling come from the file (the critic is
    freshly initialised: the export does not contain it).  Any two hidden layers of up to 64 units
    (experiments/05_impact_of_hidden_neurons trains 10 ... 64).  standardize=False drops the scaling:
    the network then sees raw observations, which is how the file's `check_sum` known answer is defined
    (utils/export.py:47-53: sum of net(ones)); `verify_check_sum` evaluates it."""
    with open(path) as f:
        data = json.load(f)
    layers = []
    while str(len(layers)) in data:
        entry = data[str(len(layers))]
        if entry.get('type', 'standard') != 'standard':
            raise NotImplementedError('only dense layers (type "standard") are supported')
        layers.append((torch.tensor(entry['weights'], dtype=torch.float32), torch.tensor(entry['biases'], dtype=torch.float32).reshape(-1)))
    if data.get('activation', 'relu') != 'relu' or len(layers) != 3:
        raise NotImplementedError('expected a relu actor with two hidden layers (ppo/defaults.py:6-19)')
    obs_dim, act_dim = layers[0][0].shape[1], layers[-1][0].shape[0]
    ac = ActorCritic(obs_dim, act_dim=act_dim, pi_hidden=(layers[0][0].shape[0], layers[1][0].shape[0]), device=device,
                     use_standardized_obs=standardize, **ac_kwargs)
    lins = [m for m in ac.pi if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for lin, (w, b) in zip(lins, layers):
            lin.weight.copy_(w)
            lin.bias.copy_(b)
        if standardize:
            sp = torch.tensor(data['scaling_parameters'], dtype=torch.float32)
            ac.obs_oms.mean.copy_(sp[0])
            ac.obs_oms.std.copy_(sp[1])
    ac.check_sum = float(data['check_sum']) if 'check_sum' in data else None
    return ac


@torch.no_grad()
def verify_check_sum(path, device='cuda', policy_kernel='tc', rtol=1e-5, atol=1e-5):
    """The reference's known-answer test for an exported policy (utils/export.py:47-53 writes it,
    utils/utils.py:324-330 is meant to check it): the sum of the network's outputs for an all-ones input
    must equal the file's `check_sum`.  Evaluated with the fused policy kernel selected by
    `policy_kernel` on CUDA devices (the torch modules on CPU).  Returns (got, expected)."""
    ac = load_policy_json(path, device=device, standardize=False, policy_kernel=policy_kernel)
    if ac.check_sum is None:
        raise ValueError(f'{path} carries no check_sum')
    d = ac.pi.net[0].in_features
    ones = torch.ones((128, d), dtype=torch.float32, device=device)       # one full tensor-core tile of identical rows
    if torch.device(device).type == 'cuda':
        n = ones.shape[0]
        act = torch.empty((n, 4), device=device); val = torch.empty(n, device=device)
        logp = torch.empty(n, device=device); mu = torch.empty((n, 4), device=device)
        ac.step_into(ones, act, val, logp, mu)
        assert bool((mu == mu[0]).all()), 'identical rows must give identical outputs'
        got = float(mu[0, :ac.log_std.shape[0]].double().sum())
    else:
        got = float(ac.pi(ones[:1]).double().sum())
    if abs(got - ac.check_sum) > atol + rtol * abs(ac.check_sum):
        raise AssertionError(f'check_sum mismatch: network gives {got}, file says {ac.check_sum}')
    return got, ac.check_sum


@torch.no_grad()
def evaluate(env_id, ac, num_evaluations=128, seed=0, device='cuda', **env_kwargs):
    """EnvironmentEvaluator.eval for `num_evaluations` episodes run in lock-step, one per environment,
    with the deterministic policy (action = mean, core.py:282-289).  Returns (returns, lengths, costs)."""
    env = VecEnv(env_id, num_evaluations, device=device, seed=seed, auto_reset=False, **env_kwargs)
    obs = env.reset()
    n = num_evaluations
    ret = torch.zeros(n, dtype=torch.float64, device=env.device)
    cost = torch.zeros(n, dtype=torch.float64, device=env.device)
    length = torch.zeros(n, dtype=torch.int64, device=env.device)
    alive = torch.ones(n, dtype=torch.bool, device=env.device)
    act = torch.empty((n, 4), dtype=torch.float32, device=env.device)
    val = torch.empty(n, dtype=torch.float32, device=env.device)
    logp = torch.empty(n, dtype=torch.float32, device=env.device)
    mu = torch.empty((n, 4), dtype=torch.float32, device=env.device)
    for _ in range(env.max_episode_steps):
        ac.step_into(obs.float().contiguous(), act, val, logp, mu)
        obs, r, term, trunc, info = env.step(mu)
        ret += r.double() * alive
        cost += info['cost'].double() * alive
        length += alive
        alive &= ~(term | trunc)
        if not bool(alive.any()):
            break
    return ret.cpu().numpy(), length.cpu().numpy(), cost.cpu().numpy()
