"""REINFORCE gradient, CPU restatement (torch float64 autograd) of ``trainer.py:Trainer.compute_grad``
(:128-225) for the batch of ONE environment slot (= one reference process).  TEST INFRASTRUCTURE.

  returns   coop[i]  = r[i] + gamma * coop[i+1]  * episode_mask[i]                        (:165-166)
            ncoop[i] = r[i] + gamma * ncoop[i+1] * episode_mask[i] * episode_mini_mask[i]
            R[i]     = mean_ratio * mean_agents(coop[i]) + (1 - mean_ratio) * ncoop[i]    (:171-172)
  advantage A = R - value.data; optional (A - mean) / std over the batch                  (:176-180)
  loss      = -sum A * logp(action) * alive + value_coeff * sum (value - R)^2 * alive
              - entr * entropy,   entropy = -sum logp * exp(logp)  (not alive-masked)      (:186-220)
  backward through the stored rollout graph: h, c chain, detached every detach_gap steps
  (trainer.py:56-60); each episode starts from zeros (trainer.py:50-51).

The policy forward is ``comm.py:134-244`` restated with differentiable torch ops (the numpy
version of oracle/policy.py is the value oracle; both are checked against each other in tests).
"""
import numpy as np
import torch


def forward_torch(p, obs, h, c, comm_action, alive, hard_attn, comm_mode="avg", comm_mask_zero=False):
    """p: dict of float64 tensors (state_dict keys).  One env: obs [N,O], h,c [N,H]."""
    n = h.shape[0]
    x = obs @ p["encoder.weight"].t() + p["encoder.bias"]
    alive_v = torch.ones(n, dtype=torch.float64) if alive is None else torch.as_tensor(alive, dtype=torch.float64)
    n_alive = float(alive_v.sum())
    g = alive_v.clone()
    if hard_attn:
        g = g * torch.as_tensor(comm_action, dtype=torch.float64)
    if comm_mask_zero:
        S = torch.zeros_like(h)
    else:
        scale = 1.0 / (n_alive - 1) if (comm_mode == "avg" and n_alive > 1) else 1.0
        mask = (1.0 - torch.eye(n, dtype=torch.float64)) * g[:, None] * g[None, :] * scale    # [src, dst]
        S = mask.t() @ h
    cvec = S @ p["C_modules.0.weight"].t() + p["C_modules.0.bias"]
    inp = x + cvec
    gates = (inp @ p["f_module.weight_ih"].t() + p["f_module.bias_ih"]
             + h @ p["f_module.weight_hh"].t() + p["f_module.bias_hh"])
    H = h.shape[1]
    gi, gf, gg, go = (gates[:, k * H:(k + 1) * H] for k in range(4))
    c2 = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
    h2 = torch.sigmoid(go) * torch.tanh(c2)
    value = (h2 @ p["value_head.weight"].t() + p["value_head.bias"])[:, 0]
    logps, k = [], 0
    while "heads.%d.weight" % k in p:
        logps.append(torch.log_softmax(h2 @ p["heads.%d.weight" % k].t() + p["heads.%d.bias" % k], dim=-1))
        k += 1
    return logps, value, h2, c2


def returns_np(reward, emask, mini, gamma, mean_ratio):
    """reward/emask/mini [T,N] -> returns [T,N] (trainer.py:160-173), float64."""
    T, n = reward.shape
    coop, ncoop, ret = np.zeros((T, n)), np.zeros((T, n)), np.zeros((T, n))
    pc, pn = np.zeros(n), np.zeros(n)
    for i in reversed(range(T)):
        coop[i] = reward[i] + gamma * pc * emask[i]
        ncoop[i] = reward[i] + gamma * pn * emask[i] * mini[i]
        pc, pn = coop[i].copy(), ncoop[i].copy()
        ret[i] = mean_ratio * coop[i].mean() + (1 - mean_ratio) * ncoop[i]
    return ret


def compute_grad(params_np, episodes, args):
    """episodes: list of dicts from oracle.rollout.run_episode (one env slot, in order).
    Returns (grads dict of float64 numpy, stat dict) for loss summed over the slot's batch."""
    p = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in params_np.items()}
    hard = bool(args.hard_attn) and bool(args.commnet)
    heads = [p["heads.%d.weight" % k].shape[0] for k in range(sum(1 for q in p if q.startswith("heads.") and q.endswith(".weight")))]
    vals, logps_all, acts, rews, emasks, minis, alives = [], [[] for _ in heads], [], [], [], [], []
    for ep in episodes:
        n, H = ep["h"].shape[1], ep["h"].shape[2]
        h = torch.zeros(n, H, dtype=torch.float64)
        c = torch.zeros(n, H, dtype=torch.float64)
        for t in range(ep["num_steps"]):
            comm = ep["comm_in"][t]
            alive = None if t == 0 else ep["alive_in"][t]
            lo, v, h, c = forward_torch(p, torch.tensor(ep["obs"][t], dtype=torch.float64), h, c,
                                        comm if hard else None, alive, hard, getattr(args, "comm_mode", "avg"),
                                        bool(args.comm_mask_zero))
            if (t + 1) % args.detach_gap == 0:                    # trainer.py:56-60
                h, c = h.detach(), c.detach()
            vals.append(v)
            for k in range(len(heads)):
                logps_all[k].append(lo[k])
        acts.append(ep["act"]); rews.append(ep["reward"]); emasks.append(ep["emask"]); minis.append(ep["mini"])
        alives.append(ep["alive"])
    values = torch.stack(vals)                                    # [T, N]
    reward, emask, mini = np.concatenate(rews), np.concatenate(emasks), np.concatenate(minis)
    action, alive = np.concatenate(acts), torch.tensor(np.concatenate(alives), dtype=torch.float64)
    ret = torch.tensor(returns_np(reward, emask, mini, args.gamma, args.mean_ratio))
    adv = ret - values.detach()
    if args.normalize_rewards:
        adv = (adv - adv.mean()) / adv.std()
    logp = [torch.stack(l) for l in logps_all]                    # per head [T, N, na]
    lp = sum(logp[k].gather(-1, torch.tensor(action[..., k:k + 1])).squeeze(-1) for k in range(len(heads)))
    action_loss = (-adv * lp * alive).sum()
    value_loss = ((values - ret).pow(2) * alive).sum()
    entropy = -sum((l * l.exp()).sum() for l in logp)
    loss = action_loss + args.value_coeff * value_loss
    if args.entr > 0:
        loss = loss - args.entr * entropy
    loss.backward()
    grads = {k: (v.grad.numpy().copy() if v.grad is not None else None) for k, v in p.items()}
    stat = dict(action_loss=float(action_loss), value_loss=float(value_loss), entropy=float(entropy),
                num_steps=int(values.shape[0]))
    return grads, stat, dict(returns=ret.numpy(), adv=adv.numpy(), values=values.detach().numpy())
