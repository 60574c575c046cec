"""Episode loop, CPU restatement of ``trainer.py:Trainer.get_episode`` (:26-126) for
ONE environment.  TEST INFRASTRUCTURE (oracle); also the body of the CPU baseline.

Per step (trainer.py:43-108):
  t == 0: comm_action = zeros (:45-46), (h, c) = zeros (:50-51), no alive mask (info = {})
  policy forward (:53-54) -> select_action (:65) -> env.step(action head 0) (:67)
  comm_action for t+1 = gate head, or ones with comm_action_one (:70-71)
  alive mask for t+1 = info['alive_mask'] of this step (comm.py:102-104)
  done = env done or t == max_steps-1 (:90); episode_mask / episode_mini_mask (:92-99)
  reward_terminal added to the last transition (:112-121)
Random draws: Philox action stream at tick = tick0 + t (env step counter), spawn
streams inside the env oracles.
"""
import numpy as np

from . import policy


def run_episode(env, params, args, seed, env_id, epoch=0, tick0=0, episode=0, forced_actions=None,
                max_steps=None):
    n, H = args.nagents, args.hid_size
    heads = policy.nheads(params)
    T = args.max_steps if max_steps is None else max_steps
    is_tj = args.env_name == "traffic_junction"
    hard = bool(args.hard_attn) and bool(args.commnet)
    if is_tj:
        env.tick = tick0
        obs = env.reset(epoch)
    else:
        obs = env.flat_obs(env.reset(seed=seed, env_id=env_id, episode=episode))
    h, c = np.zeros((n, H)), np.zeros((n, H))
    comm, alive = np.zeros(n, dtype=np.int64), None
    rec = dict(obs=[], act=[], reward=[], value=[], alive=[], mini=[], emask=[], margin=[], h=[], c=[], x=[],
               comm_in=[], alive_in=[], logp=[])
    for t in range(T):
        rec["obs"].append(obs)
        rec["comm_in"].append(comm.copy())
        rec["alive_in"].append(np.ones(n) if alive is None else alive.copy())
        lo, v, h, c, x = policy.forward(params, obs, h, c, comm if hard else None, alive, hard,
                                        getattr(args, "comm_mode", "avg"), bool(args.comm_mask_zero))
        a, margin = policy.sample_actions(lo, policy.action_draws(seed, env_id, tick0 + t, n, heads))
        if forced_actions is not None:
            a = np.asarray(forced_actions[t]).reshape(n, heads)
        if is_tj:
            obs, rew, done, info = env.step(a[:, 0], seed=seed, env_id=env_id)
            alive = info["alive_mask"]
        else:
            o, rew, done, info = env.step(a[:, 0])
            obs = env.flat_obs(o)
        if hard:
            comm = a[:, -1].copy() if not args.comm_action_one else np.ones(n, dtype=np.int64)
        done = bool(done) or t == T - 1
        emask = np.zeros(n) if done else np.ones(n)
        mini = np.ones(n)
        if not done and is_tj:
            mini = 1 - info["is_completed"]
        if done:
            rew = rew + env.reward_terminal()
        rec["act"].append(a); rec["reward"].append(rew); rec["value"].append(v)
        rec["alive"].append(np.ones(n) if alive is None else alive.copy())
        rec["mini"].append(mini); rec["emask"].append(emask); rec["margin"].append(margin)
        rec["h"].append(h); rec["c"].append(c); rec["x"].append(x)
        rec["logp"].append(np.concatenate(lo, axis=-1))
        if done:
            break
    out = {k: np.array(v) for k, v in rec.items()}
    out["success"] = int(env.stat.get("success", -1))
    out["num_steps"] = len(rec["act"])
    return out
