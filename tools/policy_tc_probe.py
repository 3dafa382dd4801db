"""GPU probe for the tcgen05 policy kernel: accuracy against a float64 torch evaluation of the same
networks and launch time, next to the CUDA-core kernel.  python tools/policy_tc_probe.py [N] [D]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phoenix_drone_simulation_b200.rollout import ActorCritic


def reference(ac, obs):
    o = obs.double()
    if ac.obs_oms:
        o = (o - ac.obs_oms.mean.double()) / (ac.obs_oms.std.double() + ac.obs_oms.eps)
    pi = ac.pi.double()
    v = ac.v.double()
    with torch.no_grad():
        mu, val = pi(o), v(o).squeeze(-1)
    ac.pi.float(); ac.v.float()
    return mu, val


def run(kernel, n, d, seed=3, reps=20, pi_hidden=(50, 50)):
    torch.manual_seed(seed)
    ac = ActorCritic(d, 4, pi_hidden=pi_hidden, device='cuda', policy_kernel=kernel, seed=11)
    for net in (ac.pi, ac.v):                     # non-trivial biases / scales
        for p in net.parameters():
            p.data.mul_(1.7).add_(0.05 * torch.randn_like(p))
    ac.obs_oms.mean.copy_(0.3 * torch.randn(d, device='cuda'))
    ac.obs_oms.std.copy_(0.5 + torch.rand(d, device='cuda'))
    obs = (ac.obs_oms.mean + 2.0 * torch.randn((n, d), device='cuda')).contiguous()
    mu_ref, v_ref = reference(ac, obs)
    act = torch.zeros((n, 4), device='cuda'); val = torch.zeros(n, device='cuda'); logp = torch.zeros(n, device='cuda')
    mu = torch.zeros((n, 4), device='cuda')
    ac.step_into(obs, act, val, logp, mu)
    torch.cuda.synchronize()
    e_mu = (mu.double() - mu_ref).abs().max().item()
    e_v = (val.double() - v_ref).abs().max().item()
    eps = (act - mu) / torch.exp(ac.log_std)
    lp = (-0.5 * eps ** 2 - ac.log_std - 0.9189385332046727).sum(-1)
    e_lp = (lp - logp).abs().max().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import ctypes as C
    h = ac.prepare_step_into(obs, act, val, logp)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        ac.step_prepared(h, st)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ac.step_prepared(h, st)
    e1.record()
    torch.cuda.synchronize()
    print(f'{kernel:8s} flags={os.environ.get("PDX_TC_FLAGS", "0")} n={n} d={d} pi={pi_hidden}: max|mu-ref|={e_mu:.3e} max|v-ref|={e_v:.3e} '
          f'logp self-consistency={e_lp:.2e} eps mean/std={eps.mean().item():+.3f}/{eps.std().item():.3f} '
          f'mu scale={mu_ref.abs().mean().item():.3f} time={1e3 * e0.elapsed_time(e1) / reps:.1f} us', flush=True)
    return e_mu, e_v


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 34
    kernels = sys.argv[3].split(',') if len(sys.argv) > 3 else ['cuda', 'tc_tf32', 'tc']
    for kernel in kernels:
        run(kernel, 1000, d)
        run(kernel, n, d, reps=50)
        run(kernel, 8 * n, d, reps=20)
    if 'tc' in kernels:
        run('tc', n, 40, pi_hidden=(64, 64))
        run('tc', 77, 48, pi_hidden=(32, 17))
