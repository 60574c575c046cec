"""cProfile of the host side of the public-API loop (observation handles).  python profiles/tools/e2e_hostprof.py"""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from ic3net_b200 import data
    from ic3net_b200.action_utils import parse_action_args, select_action
    from ic3net_b200.comm import CommNetMLP
    from ic3net_b200.trainer import Trainer
    a = bench.make_args("pp_hard_ic3net", 0, "dense")
    a.policy_impl, a.obs_chunk_mb, a.obs_api = None, 0.0, "handle"
    env = data.init(a.env_name, a)
    a.num_inputs = env.observation_dim
    a.num_actions = [env.num_actions, 2]
    a.dim_actions = 2
    parse_action_args(a)
    torch.manual_seed(0)
    net = CommNetMLP(a, a.num_inputs)
    Trainer(a, net, env)
    B, N = a.nenvs, a.nagents
    env.env.strict = False
    pin = lambda *s, dtype: torch.empty(*s, dtype=dtype).pin_memory()
    act_h, rew_h, done_h = pin(B, N, 2, dtype=torch.int32), pin(B, N, dtype=torch.float32), pin(B, dtype=torch.bool)
    comm_h, env_act_h = pin(B, N, dtype=torch.uint8), pin(B, N, dtype=torch.int32)
    state = dict(obs=env.reset(0), hc=net.init_hidden(B), info={"comm_action": torch.zeros(B, N, dtype=torch.uint8).pin_memory()})

    def step():
        action_out, value, hc = net([state["obs"], state["hc"]], state["info"])
        action = select_action(a, action_out)
        act_h.copy_(action, non_blocking=True)
        torch.cuda.synchronize()
        env_act_h.copy_(act_h[..., 0])
        obs, reward, done, info_env = env.step([env_act_h])
        rew_h.copy_(reward, non_blocking=True)
        done_h.copy_(done, non_blocking=True)
        comm_h.copy_(act_h[..., -1])
        torch.cuda.synchronize()
        state.update(obs=obs, hc=hc, info={"comm_action": comm_h})

    for _ in range(20):
        step()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        step()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
    main()
