"""CPU tests: the oracle restatement against the committed golden fixtures (which were
produced by the unmodified reference, oracle/gen_golden.py) and against the
known-answer vectors of SURVEY.md section 8(a)."""
import numpy as np
import pytest

from helpers import golden_names, load_golden, make_oracle_env, ns, tj_tables
from oracle import philox, policy
from oracle.gen_golden import make_weights
from oracle.pp_env import PredatorPreyOracle


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    def hx(c, k):
        return ["%08x" % v for v in philox.philox4x32(c, k)]
    assert hx([0, 0, 0, 0], (0, 0)) == ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]
    assert hx([0xffffffff] * 4, (0xffffffff, 0xffffffff)) == ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]
    assert hx([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0)) == \
        ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]


def test_kat_pp1():
    # SURVEY.md KAT-PP1 (harvested from the reference)
    e = PredatorPreyOracle(3, 5, 0)
    e.reset(locs=[[0, 0], [2, 2], [4, 4], [2, 3]])
    acts = [[0, 1, 2], [3, 4, 1], [1, 0, 3], [2, 2, 0], [1, 1, 0], [1, 3, 0]]
    locs = [[[0, 0], [2, 3], [4, 4]], [[0, 0], [2, 3], [4, 4]], [[0, 1], [2, 3], [4, 3]], [[1, 1], [2, 3], [3, 3]],
            [[1, 2], [2, 3], [2, 3]], [[1, 3], [2, 3], [2, 3]]]
    for t, a in enumerate(acts):
        obs, r, done, _ = e.step(a)
        assert e.predator_loc.tolist() == locs[t]
        assert np.allclose(r, [-.05, 0, -.05] if t < 4 else [-.05, 0, 0])
        assert not done
        if t == 0:
            assert set(np.flatnonzero(e.flat_obs(obs)[1])) == {13, 27, 28}
    assert e.reached.tolist() == [0, 1, 1]


def test_kat_pp2():
    e = PredatorPreyOracle(2, 4, 1)
    e.reset(locs=[[0, 0], [3, 3], [1, 1]])
    want = [[[0, 0], [3, 3]], [[0, 0], [3, 3]], [[1, 0], [2, 3]], [[1, 1], [2, 2]]]
    for t, a in enumerate([[0, 2], [3, 1], [2, 0], [1, 3]]):
        obs, r, done, _ = e.step(a)
        assert e.predator_loc.tolist() == want[t]
    assert np.allclose(r, [0, -.05])
    win = obs[0]
    classes = [[sorted(np.flatnonzero(win[y, x]).tolist()) for x in range(3)] for y in range(3)]
    assert classes == [[[0], [1], [2]], [[4], [5, 18, 19], [6]], [[8], [9], [10, 19]]]


def test_pp_errors():
    e = PredatorPreyOracle(1, 2, 0)
    e.reset(locs=[[0, 0], [0, 1]])
    e.step([1])
    assert e.episode_over
    with pytest.raises(RuntimeError):
        e.step([0])
    with pytest.raises(RuntimeError):
        PredatorPreyOracle(1, 2, 0, mode="bogus")


@pytest.mark.parametrize("name", golden_names("env_"))
def test_env_golden(name):
    meta, z = load_golden(name)
    args = ns(meta["args"])
    is_tj = args.env_name == "traffic_junction"
    env = make_oracle_env(args, tj_tables(z) if is_tj else None)
    seed, env_id = meta["seed"], meta["env_id"]
    if is_tj:
        obs = env.reset(0)
        loc = env.car_loc
    else:
        obs = env.flat_obs(env.reset(seed=seed, env_id=env_id, episode=0))
        loc = np.vstack([env.predator_loc, env.prey_loc])
    assert np.array_equal(obs, z["obs0"]) and np.array_equal(loc, z["loc0"])
    for t in range(len(z["act"])):
        if is_tj:
            obs, r, done, info = env.step(z["act"][t], seed=seed, env_id=env_id)
            loc = env.car_loc
            assert np.array_equal(info["alive_mask"], z["alive"][t])
            assert np.array_equal(info["is_completed"], z["completed"][t])
            aux = np.stack([env.wait, env.route_id, env.last_act, env.route_loc], 1)
        else:
            o, r, done, info = env.step(z["act"][t])
            obs = env.flat_obs(o)
            loc = np.vstack([env.predator_loc, env.prey_loc])
            aux = env.reached[:, None]
        assert np.array_equal(loc, z["loc"][t]), (name, t)
        assert np.array_equal(aux, z["aux"][t]), (name, t)
        assert np.array_equal(r, z["reward"][t]), (name, t)
        assert int(done) == z["done"][t]
        if "obs" in z:
            assert np.array_equal(obs.astype(np.float32), z["obs"][t]), (name, t)
    assert env.stat.get("success", -1) == meta["success"]


@pytest.mark.parametrize("name", golden_names("fwd_"))
def test_forward_golden(name):
    meta, z = load_golden(name)
    sd = make_weights(meta["weights_seed"], meta["obs_dim"], meta["hid_size"], meta["heads"], meta["comm_init"])
    p = policy.params_to_f64(sd)
    for k in range(len(z["obs"])):
        lo, v, h2, c2, x = policy.forward(p, z["obs"][k], z["h"][k], z["c"][k],
                                          z["comm"][k] if meta["hard_attn"] else None,
                                          z["alive"][k] if meta["use_alive"] else None, meta["hard_attn"],
                                          meta["comm_mode"], meta["comm_mask_zero"])
        assert np.allclose(v, z["value"][k], rtol=1e-12, atol=1e-13)
        assert np.allclose(h2, z["h2"][k], rtol=1e-12, atol=1e-13)
        assert np.allclose(c2, z["c2"][k], rtol=1e-12, atol=1e-13)
        for j in range(len(meta["heads"])):
            assert np.allclose(lo[j], z["logp%d" % j][k], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("name", golden_names("ep_"))
def test_episode_golden(name):
    """Replay trainer.get_episode with the oracle on the Philox streams."""
    from oracle.rollout import run_episode
    meta, z = load_golden(name)
    args = ns(meta["args"])
    is_tj = args.env_name == "traffic_junction"
    sd = make_weights(meta["weights_seed"], meta["obs_dim"], args.hid_size, meta["heads"], args.comm_init)
    p = policy.params_to_f64(sd)
    for i, env_id in enumerate(meta["env_ids"]):
        env = make_oracle_env(args, tj_tables(z) if is_tj else None)
        ep = run_episode(env, p, args, meta["seed"], env_id, epoch=meta["epoch"])
        g = lambda k: z["e%d_%s" % (i, k)]
        assert np.array_equal(ep["act"], g("act"))
        assert np.array_equal(ep["reward"], g("reward"))
        assert np.array_equal(ep["alive"], g("alive"))
        assert np.array_equal(ep["emask"], g("emask")) and np.array_equal(ep["mini"], g("mini"))
        assert np.allclose(ep["value"], g("value"), rtol=1e-12, atol=1e-13)
        for j, t in enumerate(g("h_steps")):
            assert np.allclose(ep["h"][t], g("h_sel")[j], rtol=1e-12, atol=1e-13)
            assert np.allclose(ep["c"][t], g("c_sel")[j], rtol=1e-12, atol=1e-13)
        assert ep["success"] == int(g("success"))


@pytest.mark.parametrize("name", golden_names("grad_"))
def test_gradient_golden(name):
    """oracle/grad.py against the gradients of the reference's own Trainer.compute_grad."""
    from oracle import grad as ograd
    from oracle.rollout import run_episode
    meta, z = load_golden(name)
    args = ns(meta["args"])
    is_tj = args.env_name == "traffic_junction"
    sd = make_weights(meta["weights_seed"], meta["obs_dim"], args.hid_size, meta["heads"], args.comm_init)
    p = policy.params_to_f64(sd)
    env = make_oracle_env(args, tj_tables(z) if is_tj else None)
    eps, tick, k = [], 0, 0
    while tick < meta["num_steps"]:
        ep = run_episode(env, p, args, meta["seed"], meta["env_id"], epoch=0, tick0=tick, episode=k)
        eps.append(ep)
        tick += ep["num_steps"]
        k += 1
    assert k == meta["num_episodes"]
    g, st, extra = ograd.compute_grad(p, eps, args)
    assert np.isclose(st["action_loss"], meta["action_loss"], rtol=1e-9, atol=1e-9)
    assert np.isclose(st["value_loss"], meta["value_loss"], rtol=1e-9, atol=1e-9)
    assert np.isclose(st["entropy"], meta["entropy"], rtol=1e-9, atol=1e-9)
    assert np.allclose(extra["returns"], z["returns"], rtol=1e-12, atol=1e-12)
    checked = 0
    for key in z.files:
        if key.startswith("g_"):
            assert np.allclose(g[key[2:]], z[key], rtol=1e-8, atol=1e-10), key
            checked += 1
        elif key.startswith("gsample_"):
            q = g[key[8:]]
            assert np.allclose(q.ravel()[::max(1, q.size // 2048)][:2048], z[key], rtol=1e-8, atol=1e-10), key
            ref = z["gsum_" + key[8:]]
            assert np.allclose([q.sum(), np.abs(q).sum(), (q ** 2).sum()], ref, rtol=1e-8)
            checked += 1
    assert checked >= 8


def test_rmsprop_golden():
    """oracle/optim.py against torch.optim.RMSprop driven like trainer.py:245-256 (fixture from gen_golden.py)."""
    from oracle import optim as ooptim
    meta, z = load_golden("rmsprop_ref")
    n = len(meta["shapes"])
    params = [z["p0_%d" % i].copy() for i in range(n)]
    sq = [np.zeros_like(p) for p in params]
    for u, ns_ in enumerate(meta["num_steps"]):
        grads = [z["g%d_%d" % (u, i)] if meta["live"][i] else None for i in range(n)]
        ooptim.rmsprop_step(params, grads, sq, meta["lr"], meta["alpha"], meta["eps"], grad_div=ns_)
        for i in range(n):
            assert np.allclose(params[i], z["p%d_%d" % (u + 1, i)], rtol=1e-12, atol=1e-14), (u, i)
    for i in range(n):
        if meta["live"][i]:
            assert np.allclose(sq[i], z["v_%d" % i], rtol=1e-12, atol=0)
        else:                                   # never touched: parameter unchanged, no state
            assert np.array_equal(params[i], z["p0_%d" % i])


@pytest.mark.parametrize("name", golden_names("grad_"))
def test_manual_bptt_matches_reference_gradients(name):
    """oracle/bptt.py (explicit per-step backward formulas = the arithmetic of hand-written BPTT kernels) against
    the gradients of the reference's own Trainer.compute_grad and against the autograd oracle."""
    from oracle import bptt
    from oracle import grad as ograd
    from oracle.rollout import run_episode
    meta, z = load_golden(name)
    args = ns(meta["args"])
    is_tj = args.env_name == "traffic_junction"
    sd = make_weights(meta["weights_seed"], meta["obs_dim"], args.hid_size, meta["heads"], args.comm_init)
    p = policy.params_to_f64(sd)
    env = make_oracle_env(args, tj_tables(z) if is_tj else None)
    eps, tick, k = [], 0, 0
    while tick < meta["num_steps"]:
        ep = run_episode(env, p, args, meta["seed"], meta["env_id"], epoch=0, tick0=tick, episode=k)
        eps.append(ep)
        tick += ep["num_steps"]
        k += 1
    g, st = bptt.compute_grad_manual(p, eps, args)
    assert np.isclose(st["action_loss"], meta["action_loss"], rtol=1e-9, atol=1e-9)
    assert np.isclose(st["value_loss"], meta["value_loss"], rtol=1e-9, atol=1e-9)
    assert np.isclose(st["entropy"], meta["entropy"], rtol=1e-9, atol=1e-9)
    checked = 0
    for key in z.files:
        if key.startswith("g_"):
            assert np.allclose(g[key[2:]], z[key], rtol=1e-8, atol=1e-10), key
            checked += 1
        elif key.startswith("gsample_"):
            q = g[key[8:]]
            assert np.allclose(q.ravel()[::max(1, q.size // 2048)][:2048], z[key], rtol=1e-8, atol=1e-10), key
            assert np.allclose([q.sum(), np.abs(q).sum(), (q ** 2).sum()], z["gsum_" + key[8:]], rtol=1e-8)
            checked += 1
    assert checked >= 8
    ga, _, _ = ograd.compute_grad(p, eps, args)
    for key, v in ga.items():
        if v is None:
            assert g[key] is None
        else:
            assert np.allclose(g[key], v, rtol=1e-9, atol=1e-11), key
