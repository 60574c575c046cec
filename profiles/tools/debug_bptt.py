"""Development aid: per-parameter gradient comparison of the BPTT kernels against the autograd recompute path
on the same recorded rollout (GPU).  python profiles/tools/debug_bptt.py [fixture] [B]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import finish_args, load_golden, ns  # noqa: E402
from oracle.gen_golden import make_weights  # noqa: E402


def build(meta, B, grad_impl, **over):
    from ic3net_b200 import data
    from ic3net_b200.comm import CommNetMLP
    from ic3net_b200.trainer import Trainer
    args = ns(meta["args"], nenvs=B, seed=808, env_id0=30, obs_mode="index", use_graph=False, policy_impl="tc",
              record_for_grad=True, grad_window=16, grad_impl=grad_impl, **over)
    env = data.init(args.env_name, args)
    finish_args(args, env)
    net = CommNetMLP(args, args.num_inputs)
    sd = make_weights(meta["weights_seed"], args.num_inputs, args.hid_size, args.naction_heads, args.comm_init)
    net.load_state_dict({k: torch.from_numpy(v).float() for k, v in sd.items()})
    return args, net, Trainer(args, net, env)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "grad_pp_easy_ic3net"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    meta, z = load_golden(name)
    res = {}
    for impl in ("autograd", "kernels"):
        args, net, tr = build(meta, B, impl)
        batch, stat = tr.run_batch(0)
        tr.optimizer.zero_grad(set_to_none=False)
        s = tr.compute_grad(batch)
        torch.cuda.synchronize()
        res[impl] = ({k: p.grad.detach().double().cpu().numpy().copy() for k, p in net.named_parameters()}, s, stat,
                     int(tr._buf["err"].item()))
    ga, sa, sta, ea = res["autograd"]
    gk, sk, stk, ek = res["kernels"]
    print("steps", sta["num_steps"], stk["num_steps"], "err flags", ea, ek)
    print("losses autograd", sa)
    print("losses kernels ", sk)
    for k in ga:
        den = max(np.abs(ga[k]).max(), 1e-30)
        print("%-28s max|ref| %.3e  rel err %.3e  (nan %d)" % (k, np.abs(ga[k]).max(), np.abs(gk[k] - ga[k]).max() / den,
                                                              int(np.isnan(gk[k]).sum())))


if __name__ == "__main__":
    main()
