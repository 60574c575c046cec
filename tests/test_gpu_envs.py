"""GPU parity tests of the environment kernels (csrc/pp_env.cu, csrc/tj_env.cu) through the
reference-shaped classes.  Bar: bit-exact integer state / masks / observations, rewards
equal to float32(reference float64 reward)."""
import numpy as np
import pytest
import torch

from helpers import golden_names, load_golden, make_oracle_env, ns, tj_tables

pytestmark = pytest.mark.gpu


def make_env(args, **over):
    from ic3net_b200 import data
    for k, v in over.items():
        setattr(args, k, v)
    return data.init(args.env_name, args)


def cpu(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name", golden_names("env_"))
def test_env_matches_reference_golden(name):
    """B = 1, same Philox streams as the fixture (which the unmodified reference produced)."""
    meta, z = load_golden(name)
    args = ns(meta["args"], nenvs=1, seed=meta["seed"], env_id0=meta["env_id"])
    w = make_env(args)
    env = w.env
    is_tj = args.env_name == "traffic_junction"
    assert w.observation_dim == meta["obs_dim"]
    obs = w.reset(0)
    loc = cpu(env.car_loc if is_tj else env.loc)[0]
    assert np.array_equal(loc, z["loc0"])
    assert np.array_equal(cpu(obs)[0], z["obs0"].astype(np.float32))
    for t in range(len(z["act"])):
        obs, r, done, info = w.step([z["act"][t][None]])
        if is_tj:
            loc = cpu(env.car_loc)[0]
            aux = np.stack([cpu(env.wait)[0], cpu(env.route_id)[0], cpu(env.car_last_act)[0],
                            cpu(env.car_route_loc)[0]], 1)
            assert np.array_equal(cpu(info["alive_mask"])[0], z["alive"][t])
            assert np.array_equal(cpu(info["is_completed"])[0], z["completed"][t])
        else:
            loc = cpu(env.loc)[0]
            aux = cpu(env.reached_prey)[0][:, None]
        assert np.array_equal(loc, z["loc"][t]), (name, t)
        assert np.array_equal(aux, z["aux"][t]), (name, t)
        assert np.array_equal(cpu(r)[0], z["reward"][t].astype(np.float32)), (name, t)
        assert int(cpu(done)[0]) == z["done"][t]
        if "obs" in z:
            assert np.array_equal(cpu(obs)[0], z["obs"][t]), (name, t)
    assert w.get_stat().get("success", -1) == meta["success"]


@pytest.mark.parametrize("name", ["env_pp_easy", "env_pp_v1", "env_pp_hard", "env_tj_medium", "env_tj_hard_v1",
                                  "env_tj_easy", "env_pp_enemy", "env_pp_enemy_coop"])
def test_batched_envs_match_oracle(name):
    """B = 37 envs with distinct Philox streams vs 37 oracle instances, random actions."""
    meta, z = load_golden(name)
    B, T, seed, id0 = 37, 25, 1234, 1000
    args = ns(meta["args"], nenvs=B, seed=seed, env_id0=id0)
    w = make_env(args)
    env = w.env
    is_tj = args.env_name == "traffic_junction"
    tables = tj_tables(z) if is_tj else None
    orcs = [make_oracle_env(args, tables) for _ in range(B)]
    obs = cpu(w.reset(0))
    for b, o in enumerate(orcs):
        oo = o.reset(0) if is_tj else o.flat_obs(o.reset(seed=seed, env_id=id0 + b, episode=0))
        assert np.array_equal(obs[b], oo.astype(np.float32))
    rs = np.random.RandomState(5)
    env.strict = False
    for t in range(T):
        act = rs.randint(0, env.naction, size=(B, args.nagents))
        done_before = np.array([o.episode_over for o in orcs])
        obs, r, done, info = w.step([act])
        obs, r, done = cpu(obs), cpu(r), cpu(done)
        for b, o in enumerate(orcs):
            if done_before[b]:
                continue          # stepping a finished env is an error on both sides (checked elsewhere)
            if is_tj:
                oo, orr, od, _ = o.step(act[b], seed=seed, env_id=id0 + b)
            else:
                oo, orr, od, _ = o.step(act[b])
                oo = o.flat_obs(oo)
            assert np.array_equal(obs[b], oo.astype(np.float32)), (name, t, b)
            assert np.array_equal(r[b], orr.astype(np.float32)), (name, t, b)
            assert bool(done[b]) == bool(od)
    env.err.zero_()


def test_tj_explicit_draws_tape():
    """Spawn decisions from an explicit [B,G,3] tape instead of the Philox stream."""
    meta, z = load_golden("env_tj_medium")
    B = 9
    args = ns(meta["args"], nenvs=B, seed=0, env_id0=0)
    w = make_env(args)
    env = w.env
    orcs = [make_oracle_env(args, tj_tables(z)) for _ in range(B)]
    w.reset(0)
    [o.reset(0) for o in orcs]
    rs = np.random.RandomState(9)
    G = env.cfg.G
    for t in range(30):
        act = rs.randint(0, 2, size=(B, args.nagents))
        draws = rs.randint(0, 1 << 24, size=(B, G, 3))
        draws[:, :, 0] = np.where(rs.rand(B, G) < 0.5, 0, draws[:, :, 0])   # force some spawns
        obs, r, done, info = env.step(act, draws=draws)
        for b, o in enumerate(orcs):
            oo, orr, _, oi = o.step(act[b], draws=draws[b])
            assert np.array_equal(cpu(obs)[b], oo.astype(np.float32))
            assert np.array_equal(cpu(r)[b], orr.astype(np.float32))
            assert np.array_equal(cpu(info["alive_mask"])[b], oi["alive_mask"])


def test_pp_set_state_and_kat():
    """SURVEY KAT-PP2 through the CUDA path (injected spawn positions)."""
    import argparse
    a = argparse.Namespace(env_name="predator_prey", nagents=2, nfriendly=2, dim=4, vision=1, mode="mixed",
                           nenemies=1, no_stay=False, moving_prey=False, enemy_comm=False, nenvs=1, seed=0)
    w = make_env(a)
    env = w.env
    env.set_state([[[0, 0], [3, 3]]], [[[1, 1]]])
    want = [[[0, 0], [3, 3]], [[0, 0], [3, 3]], [[1, 0], [2, 3]], [[1, 1], [2, 2]]]
    for t, act in enumerate([[0, 2], [3, 1], [2, 0], [1, 3]]):
        obs, r, done, info = env.step([act])
        assert cpu(info["predator_locs"])[0].tolist() == want[t]
    assert np.allclose(cpu(r)[0], [0, -.05])
    win = cpu(obs)[0, 0]
    classes = [[sorted(np.flatnonzero(win[y, x]).tolist()) for x in range(3)] for y in range(3)]
    assert classes == [[[0], [1], [2]], [[4], [5, 18, 19], [6]], [[8], [9], [10, 19]]]


def test_pp_episode_done_raises():
    import argparse
    a = argparse.Namespace(env_name="predator_prey", nagents=1, nfriendly=1, dim=2, vision=0, mode="mixed",
                           nenemies=1, no_stay=False, moving_prey=False, enemy_comm=False, nenvs=1, seed=0)
    env = make_env(a).env
    env.set_state([[[0, 0]]], [[[0, 1]]])
    obs, r, done, _ = env.step([[1]])
    assert bool(done[0])
    with pytest.raises(RuntimeError, match="Episode is done"):
        env.step([[0]])
    a.mode = "bogus"
    with pytest.raises(RuntimeError, match="Incorrect mode"):
        make_env(a)
    a.mode, a.moving_prey = "mixed", True
    with pytest.raises(NotImplementedError):
        make_env(a)


def test_pp_hard_full_size_properties():
    """BASELINE config c2 (8192 envs, 10 agents, dim 20, vision 1): size-independent
    properties of the observation tensor and of the spawn."""
    import argparse
    B, N, D = 8192, 10, 20
    a = argparse.Namespace(env_name="predator_prey", nagents=N, nfriendly=N, dim=D, vision=1, mode="mixed",
                           nenemies=1, no_stay=False, moving_prey=False, enemy_comm=False, nenvs=B, seed=7)
    w = make_env(a)
    env = w.env
    obs = w.reset(0)
    assert obs.shape == (B, N, 3636)
    loc = env.loc.long()
    cell = loc[..., 0] * D + loc[..., 1]
    assert int((cell.sort(1).values.diff(dim=1) == 0).sum()) == 0        # N+1 distinct cells per env
    assert int(cell.min()) >= 0 and int(cell.max()) < D * D
    o = obs.view(B, N, 9, 404)
    assert torch.equal(o[..., :402].sum(-1), torch.ones(B, N, 9, device=obs.device))   # one cell-id/OUTSIDE hot
    centre = o[:, :, 4]
    own = (loc[:, :N, 0] * D + loc[:, :N, 1])
    assert torch.equal(centre.gather(-1, own.unsqueeze(-1)).squeeze(-1), torch.ones(B, N, device=obs.device))
    assert float(centre[..., 403].min()) >= 1.0                              # own cell counts the agent itself
    assert float(o[..., 402].sum()) == float((((loc[:, :N] - loc[:, N:]).abs().max(-1).values) <= 1).sum())
    # moves are clamped and actions 4 = stay leave everything unchanged
    before = env.loc.clone()
    w.step([torch.full((B, N), 4, dtype=torch.int32, device=obs.device)])
    assert torch.equal(before, env.loc)
    for _ in range(25):
        w.step([torch.zeros(B, N, dtype=torch.int32, device=obs.device)])   # everybody UP
    assert int(env.loc[:, :N, 0].max()) == 0 or bool(env.reached_prey.any())
