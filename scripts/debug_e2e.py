import sys, time, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch, numpy as np
import bench
from ic3net_b200 import data, _lib
from ic3net_b200.action_utils import parse_action_args, select_action
from ic3net_b200.comm import CommNetMLP
from ic3net_b200.trainer import Trainer

def build(obs_mode):
    a = bench.make_args("pp_hard_ic3net", 0, obs_mode)
    env = data.init(a.env_name, a)
    a.num_inputs = env.observation_dim; a.num_actions = [env.num_actions, 2]; a.dim_actions = 2
    parse_action_args(a)
    torch.manual_seed(0)
    net = CommNetMLP(a, a.num_inputs)
    return a, env, net, Trainer(a, net, env)

def show(tag, r):
    print(tag, "ms/step %.3f" % r["ms_per_step"], "median", r["step_ms_median"], "tail", r["step_ms_sorted_tail"], r["phases_ms"], "mallocs", r["cuda_mallocs_in_loop"], r["cuda_frees_in_loop"])

which = sys.argv[1]
a, env, net, tr = build("dense")
if which == "fresh":
    show("fresh", bench.e2e_loop(a, env, net, 40, np, torch, select_action))
elif which == "after_rollout":
    tr._alloc(80); env.env.reset(want_obs=False); tr._enqueue(80); torch.cuda.synchronize()
    show("after_rollout", bench.e2e_loop(a, env, net, 40, np, torch, select_action))
elif which == "after_kernels":
    tr._alloc(80); env.env.reset(want_obs=False); tr._enqueue(80); torch.cuda.synchronize()
    bench.per_kernel_times(tr, a, env, net, 20, C, torch, _lib)
    show("after_kernels", bench.e2e_loop(a, env, net, 40, np, torch, select_action))
elif which == "after_alt":
    tr._alloc(80); env.env.reset(want_obs=False); tr._enqueue(80); torch.cuda.synchronize()
    a2, env2, net2, tr2 = build("index"); tr2._alloc(80); env2.env.reset(want_obs=False); tr2._enqueue(80); torch.cuda.synchronize()
    del tr2, net2, env2
    show("after_alt", bench.e2e_loop(a, env, net, 40, np, torch, select_action))
    show("after_alt_again", bench.e2e_loop(a, env, net, 40, np, torch, select_action))
