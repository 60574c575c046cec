"""Where does an e2e step go?  GPU time (CUDA events) and host time per phase of the public-API loop.
python profiles/tools/e2e_probe.py [handle|dense]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    api = sys.argv[1] if len(sys.argv) > 1 else "handle"
    from ic3net_b200 import data
    from ic3net_b200.action_utils import parse_action_args, select_action
    from ic3net_b200.comm import CommNetMLP
    from ic3net_b200.trainer import Trainer
    a = bench.make_args("pp_hard_ic3net", 0, "dense")
    a.policy_impl, a.obs_chunk_mb, a.obs_api = None, 0.0, api
    env = data.init(a.env_name, a)
    a.num_inputs = env.observation_dim
    a.num_actions = [env.num_actions, 2]
    a.dim_actions = 2
    parse_action_args(a)
    torch.manual_seed(0)
    net = CommNetMLP(a, a.num_inputs)
    tr = Trainer(a, net, env)
    B, N = a.nenvs, a.nagents
    obs = env.reset(0)
    hc = net.init_hidden(B)
    info = {"comm_action": torch.zeros(B, N, dtype=torch.uint8).pin_memory()}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tot = dict(fwd_gpu=0.0, sel_gpu=0.0, env_gpu=0.0, fwd_host=0.0, sel_host=0.0, env_host=0.0)
    steps = 60
    for t in range(steps + 5):
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        torch.cuda.synchronize()
        h0 = time.perf_counter()
        e0.record()
        action_out, value, hc = net([obs, hc], info)
        e1.record()
        h1 = time.perf_counter()
        action = select_action(a, action_out)
        e2.record()
        h2 = time.perf_counter()
        torch.cuda.synchronize()
        h2b = time.perf_counter()
        obs, reward, done, info_env = env.step([action[..., 0]])
        e3.record()
        h3 = time.perf_counter()
        torch.cuda.synchronize()
        info = {"comm_action": action[..., -1].to(torch.uint8)}
        if t >= 5:
            tot["fwd_gpu"] += e0.elapsed_time(e1); tot["sel_gpu"] += e1.elapsed_time(e2); tot["env_gpu"] += e2.elapsed_time(e3)
            tot["fwd_host"] += 1e3 * (h1 - h0); tot["sel_host"] += 1e3 * (h2 - h1); tot["env_host"] += 1e3 * (h3 - h2b)
    print(api, {k: round(v / steps, 4) for k, v in tot.items()})


if __name__ == "__main__":
    main()
