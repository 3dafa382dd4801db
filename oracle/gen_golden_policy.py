"""TEST INFRASTRUCTURE ONLY -- golden vector for the policy JSON format: a JSON written by this
repo's exporter is read by the UNMODIFIED reference loader (utils/utils.py:309-340,
load_network_json -> build_mlp_network) and evaluated on fixed observations.

    python oracle/gen_golden_policy.py      # writes tests/golden_collector/policy_json.{json,npz}
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get('PHOENIX_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, 'shim'))
sys.path.insert(0, REFERENCE)

from phoenix_drone_simulation.utils.utils import load_network_json            # noqa: E402  (reference)
from phoenix_drone_simulation_b200.policy_io import export_policy_json         # noqa: E402
from phoenix_drone_simulation_b200.rollout import ActorCritic                   # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden_collector')


def main():
    torch.manual_seed(7)
    ac = ActorCritic(34, device='cpu', fused=False)
    ac.obs_oms.mean.copy_(torch.randn(34) * 0.2)
    ac.obs_oms.std.copy_(torch.rand(34) + 0.5)
    path = os.path.join(OUT, 'policy_json.json')
    export_policy_json(ac, path)
    net = load_network_json(path)                       # the reference's loader
    obs = torch.randn(64, 34) * 1.5
    with torch.no_grad():
        scaled = (obs - ac.obs_oms.mean) / (ac.obs_oms.std + 1e-5)     # online_mean_std.py:42-48
        out = net(scaled).numpy()
    np.savez_compressed(os.path.join(OUT, 'policy_json.npz'), obs=obs.numpy(), mu=out)
    print('wrote policy_json.json / .npz; reference net:', net)


if __name__ == '__main__':
    main()
