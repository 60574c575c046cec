"""GPU edge cases: single agent, 32 agents (lane limit), one env, ragged 128-row tile tails, TJ curriculum through
reset(epoch), hid_size 64 rollouts on the SIMT path, and the reference-flag CLI end to end."""
import argparse

import numpy as np
import pytest
import torch

from helpers import finish_args, load_golden, make_oracle_env, ns, tj_tables
from oracle import policy as opolicy
from oracle.gen_golden import make_weights
from oracle.rollout import run_episode

pytestmark = pytest.mark.gpu


def cpu(t):
    return t.detach().cpu().numpy()


def close(a, b, tol=1e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b)))


def pp_args(N, dim, vision, B, hid=128, **kw):
    d = dict(env_name="predator_prey", nagents=N, nfriendly=N, dim=dim, vision=vision, mode="mixed", nenemies=1,
             no_stay=False, moving_prey=False, enemy_comm=False, nenvs=B, seed=11, env_id0=3, hid_size=hid,
             recurrent=True, rnn_type="LSTM", commnet=True, hard_attn=True, comm_action_one=False, comm_mode="avg",
             comm_passes=1, comm_mask_zero=False, comm_init="uniform", share_weights=False, max_steps=12,
             batch_size=24, lrate=1e-3, obs_mode="index", use_graph=False, continuous=False, detach_gap=10000,
             gamma=1.0, mean_ratio=0.0, value_coeff=0.01, entr=0.0, normalize_rewards=False)
    d.update(kw)
    return argparse.Namespace(**d)


def rollout_vs_oracle(args, T, tables=None, slots=None):
    from ic3net_b200 import data
    from ic3net_b200.comm import CommNetMLP
    from ic3net_b200.trainer import Trainer
    env = data.init(args.env_name, args)
    finish_args(args, env)
    net = CommNetMLP(args, args.num_inputs)
    sd = make_weights(5, args.num_inputs, args.hid_size, args.naction_heads)
    net.load_state_dict({k: torch.from_numpy(v).float() for k, v in sd.items()})
    tr = Trainer(args, net, env)
    batch = tr.rollout(T, 0)
    stat = tr.collect_stat()
    p = opolicy.params_to_f64(sd)
    act, rew, val = cpu(batch.action), cpu(batch.reward), cpu(batch.value)
    for b in (range(args.nenvs) if slots is None else slots):
        orc = make_oracle_env(args, tables)
        t0, k = 0, 0
        while t0 < T:
            ep = run_episode(orc, p, args, args.seed, args.env_id0 + b, tick0=t0, episode=k,
                             forced_actions=act[t0:, b], max_steps=min(args.max_steps, T - t0))
            L = ep["num_steps"]
            assert np.array_equal(rew[t0:t0 + L, b], ep["reward"].astype(np.float32)), (b, k)
            assert close(val[t0:t0 + L, b], ep["value"]), (b, k)
            t0, k = t0 + L, k + 1
    return stat


def test_single_agent_env():
    stat = rollout_vs_oracle(pp_args(1, 3, 1, 5), 30)          # N = 1: nobody to talk to, S = 0
    assert stat["num_steps"] == 150


def test_thirty_one_predators_and_one_env():
    rollout_vs_oracle(pp_args(31, 8, 1, 1), 14)                # lane 31 holds the prey; B = 1


def test_thirty_one_predators_and_a_communicating_prey():
    # --enemy_comm at the lane limit: 31 predators + the prey = 32 agent rows of the policy, prey on lane 31
    stat = rollout_vs_oracle(pp_args(32, 8, 1, 2, nfriendly=31, enemy_comm=True), 14)
    assert len(stat["reward"]) == 31 and len(stat["enemy_reward"]) == 1 and len(stat["enemy_comm"]) == 1


def test_thirty_two_cars():
    meta, z = load_golden("env_tj_hard")
    a = ns(meta["args"], nagents=32, nfriendly=32, nenvs=3, seed=2, env_id0=0, hid_size=128, recurrent=True,
           rnn_type="LSTM", commnet=True, ic3net=True, hard_attn=True, comm_action_one=True, comm_mode="avg",
           comm_passes=1, comm_mask_zero=False, comm_init="uniform", share_weights=False, max_steps=25, batch_size=25,
           lrate=1e-3, obs_mode="dense", use_graph=False, continuous=False)
    rollout_vs_oracle(a, 25, tj_tables(z))


def test_ragged_tile_tail_and_straddling_envs():
    # R = 13 * 7 = 91 rows and R = 37 * 7 = 259 rows: partial last tile, envs straddling the 128-row boundary
    for B in (13, 37):
        rollout_vs_oracle(pp_args(7, 5, 1, B), 20, slots=[0, B // 2, B - 1, min(B - 1, 18)])


def test_hid64_rollout_simt():
    rollout_vs_oracle(pp_args(4, 4, 0, 6, hid=64), 30)


def test_tj_curriculum_through_reset():
    """add_rate schedule (traffic_junction_env.py:196-200,620-626) driven by reset(epoch) changes the spawn threshold."""
    from ic3net_b200 import data
    meta, z = load_golden("env_tj_medium")
    a = ns(meta["args"], nenvs=4, seed=1, env_id0=0, add_rate_min=0.1, add_rate_max=0.3, curr_start=0, curr_end=10)
    w = data.init(a.env_name, a)
    orc = make_oracle_env(a, tj_tables(z))
    for epoch in (0, 1, 2, 2, 5, 11, 12):
        w.reset(epoch)
        orc.reset(epoch)
        assert w.env.add_rate == orc.add_rate
        assert w.env.cfg.spawn_thr == orc.spawn_threshold()
    assert w.env.add_rate > 0.1


def test_cli_runs_reference_flags(capsys):
    from ic3net_b200 import main as cli
    rc = cli.main(["--env_name", "predator_prey", "--nagents", "3", "--dim", "5", "--vision", "0", "--max_steps", "20",
                   "--hid_size", "128", "--ic3net", "--recurrent", "--nenvs", "64", "--num_epochs", "2", "--epoch_size",
                   "2", "--batch_size", "40", "--seed", "4", "--detach_gap", "10", "--lrate", "0.001"])
    out = capsys.readouterr().out
    assert rc == 0 and "Epoch 2" in out and "steps/s" in out
    # the independent-controller baselines of models.py (main.py:162-169): MLP, and RNN with --recurrent (IC / IRIC)
    small = ["--env_name", "predator_prey", "--nagents", "3", "--dim", "5", "--max_steps", "10", "--nenvs", "4",
             "--hid_size", "64", "--num_epochs", "1", "--epoch_size", "1", "--batch_size", "10", "--seed", "2"]
    assert cli.main(small) == 0
    assert cli.main(small + ["--recurrent", "--mean_ratio", "0"]) == 0
    # --enemy_comm (main.py:124-131): the prey joins the policy; its reward / gate statistics print on their own lines
    capsys.readouterr()
    rc = cli.main(["--env_name", "predator_prey", "--nagents", "3", "--dim", "5", "--vision", "1", "--max_steps", "10",
                   "--hid_size", "128", "--ic3net", "--recurrent", "--enemy_comm", "--nenvs", "16", "--num_epochs", "1",
                   "--epoch_size", "2", "--batch_size", "10", "--seed", "3"])
    out = capsys.readouterr().out
    assert rc == 0 and "Enemy-Reward: [" in out and "Enemy-Comm: [" in out
    rc = cli.main(["--env_name", "traffic_junction", "--nagents", "5", "--dim", "6", "--vision", "0", "--max_steps",
                   "20", "--hid_size", "128", "--ic3net", "--recurrent", "--nenvs", "32", "--num_epochs", "1",
                   "--epoch_size", "1", "--batch_size", "20", "--seed", "4", "--difficulty", "easy", "--add_rate_min",
                   "0.3", "--add_rate_max", "0.3", "--rollout_only"])
    assert rc == 0


def test_fp16_operand_range_is_flagged_not_silent():
    """The tensor-core path splits operands into fp16 halves (|activation| < 4094, |folded weight| < 255).  Weights or
    activations outside that range must raise the device flag (IC3_ERR_FP16_RANGE) -- surfaced as an exception by
    Trainer.collect_stat -- instead of silently producing inf; the fp32 SIMT path accepts the same model."""
    from ic3net_b200 import _lib, data
    from ic3net_b200.comm import CommNetMLP
    from ic3net_b200.trainer import Trainer

    def run(impl, scale_w=1.0, scale_enc=1.0):
        args = pp_args(3, 5, 0, 4, policy_impl=impl)
        env = data.init(args.env_name, args)
        finish_args(args, env)
        torch.manual_seed(0)
        net = CommNetMLP(args, args.num_inputs)
        with torch.no_grad():
            net.f_module.weight_hh.mul_(scale_w)
            net.encoder.weight.mul_(scale_enc)
        tr = Trainer(args, net, env)
        tr.rollout(4, 0)
        return tr

    run("tc").collect_stat()                                             # in range: no flag
    with pytest.raises(RuntimeError, match="0x200"):                     # |w| * 256 >= 65504
        run("tc", scale_w=4000.0).collect_stat()
    with pytest.raises(RuntimeError, match="0x200"):                     # |x| * 16 >= 65504
        run("tc", scale_enc=1.0e5).collect_stat()
    tr = run("simt", scale_w=4000.0)                                     # fp32 kernels: same weights are fine
    tr.collect_stat()
    assert _lib.ERR_FP16_RANGE == 0x200
