"""Shared test helpers (fixture loading, argument namespaces)."""
import argparse
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, z


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def unpack_routes(ln, cells):
    return [[cells[g, k, :ln[g, k]] for k in range(ln.shape[1])] for g in range(ln.shape[0])]


def tj_tables(z):
    return {"grid": z["grid"], "routes": unpack_routes(z["route_len"], z["route_cells"])}


def ns(meta_args, **over):
    """argparse.Namespace from a fixture's recorded args (the main.py flag names)."""
    d = dict(meta_args)
    d.update(over)
    a = argparse.Namespace(**d)
    if not hasattr(a, "nfriendly"):
        a.nfriendly = a.nagents
    return a


def finish_args(args, env):
    """main.py:134-155 (derived fields)."""
    from ic3net_b200.action_utils import parse_action_args
    args.num_inputs = env.observation_dim
    na = env.num_actions
    args.num_actions = [na] if not isinstance(na, (list, tuple)) else list(na)
    args.dim_actions = env.dim_actions
    if args.hard_attn and args.commnet:
        args.num_actions = [*args.num_actions, 2]
        args.dim_actions = env.dim_actions + 1
    if args.commnet and (args.recurrent or args.rnn_type == "LSTM"):
        args.recurrent = True
        args.rnn_type = "LSTM"
    parse_action_args(args)
    return args


def make_oracle_env(args, tables=None):
    from oracle import pp_env, tj_env
    if args.env_name == "predator_prey":
        return pp_env.PredatorPreyOracle(args.nfriendly, args.dim, args.vision, args.mode, args.nenemies, args.no_stay,
                                         getattr(args, "enemy_comm", False))
    return tj_env.TrafficJunctionOracle(args.nagents, args.dim, args.vision, args.difficulty, tables,
                                        args.add_rate_min, args.add_rate_max, args.curr_start, args.curr_end)
