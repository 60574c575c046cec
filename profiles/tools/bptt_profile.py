"""One small training update at bench scale (for ncu launch lists of the backward kernels).
python profiles/tools/bptt_profile.py [workload] [batch_size]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "pp_hard_ic3net"
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    from ic3net_b200 import data
    from ic3net_b200.action_utils import parse_action_args
    from ic3net_b200.comm import CommNetMLP
    from ic3net_b200.trainer import Trainer
    a = bench.make_args(wl, 0, "index")
    a.policy_impl = None
    for k, v in dict(record_for_grad=True, batch_size=bs, grad_impl="auto", batch_boundary="cut", value_coeff=0.01,
                     entr=0.0, gamma=1.0, normalize_rewards=False, detach_gap=10000).items():
        setattr(a, k, v)
    a.max_steps = bs          # T = bs lock-steps
    env = data.init(a.env_name, a)
    a.num_inputs = env.observation_dim
    a.num_actions = [env.num_actions] + ([2] if a.hard_attn else [])
    a.dim_actions = len(a.num_actions)
    parse_action_args(a)
    torch.manual_seed(0)
    net = CommNetMLP(a, a.num_inputs)
    tr = Trainer(a, net, env)
    for u in range(2):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        batch, stat = tr.run_batch(u)
        e1.record()
        tr.optimizer.zero_grad(set_to_none=False)
        s = tr.compute_grad(batch)
        e2.record()
        torch.cuda.synchronize()
        print("update %d: T=%d rollout %.3f ms, compute_grad %.3f ms (%.3f ms/step), losses %s" % (
            u, tr.steps_per_batch(), e0.elapsed_time(e1), e1.elapsed_time(e2), e1.elapsed_time(e2) / tr.steps_per_batch(), s))


if __name__ == "__main__":
    main()
