import sys, os, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.set_num_threads(1)
import bench
from ic3net_b200 import data
from ic3net_b200.action_utils import parse_action_args, select_action
from ic3net_b200.comm import CommNetMLP
a = bench.make_args("pp_hard_ic3net", 0, "dense")
env = data.init(a.env_name, a)
a.num_inputs = env.observation_dim; a.num_actions = [env.num_actions, 2]; a.dim_actions = 2
parse_action_args(a)
net = CommNetMLP(a, a.num_inputs)
bench.e2e_loop(a, env, net, 20, np, torch, select_action)
pr = cProfile.Profile(); pr.enable()
r = bench.e2e_loop(a, env, net, 200, np, torch, select_action)
pr.disable()
print(r["ms_per_step"], r["phases_ms"])
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
