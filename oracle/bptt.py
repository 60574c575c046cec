"""Hand-derived backward pass (BPTT) of the rollout loss, numpy float64.  TEST INFRASTRUCTURE -- never imported
by the product.

``oracle/grad.py`` restates ``Trainer.compute_grad`` (trainer.py:128-225) with torch autograd; this module states
the SAME gradient as explicit per-step formulas, i.e. the arithmetic that hand-written backward kernels have to
perform ("needs either hand-written BPTT ... or recompute-with-autograd").  It is pinned by
``tests/test_oracle_golden.py::test_manual_bptt_matches_reference_gradients`` against the gradients of the
reference's own ``compute_grad`` (tests/golden/grad_*.npz).

Forward of one step of one environment (comm.py:134-244, N agents, H hidden units), with the non-differentiable
gate ``g = alive * comm_action`` (hard attention; ones for CommNet) and ``scale = 1/(n_alive-1)`` (comm_mode avg):

    x   = obs W_e^T + b_e                                    (comm.py:119)
    S_k = g_k * scale * sum_{j != k} g_j h_j                 (comm.py:181-205)
    u   = x + S W_c^T + b_c                                  (comm.py:206-215)
    a   = u W_ih^T + b_ih + h W_hh^T + b_hh                  (LSTMCell, gate order i, f, g, o)
    c'  = sig(a_f) c + sig(a_i) tanh(a_g);   h' = sig(a_o) tanh(c')
    v   = h' w_v + b_v;   logp^m = log_softmax(h' W_m^T + b_m) per head m

Loss of the batch (trainer.py:186-220), A = R - v treated as a constant:

    L = sum_t sum_k alive [ -A * sum_m logp^m[action_m] + value_coeff (v - R)^2 ] + entr * sum logp*exp(logp)

Backward of one step, given dL/dh' (``dh``) and dL/dc' (``dc``) arriving from step t+1 (zero at an episode end and
wherever the reference detaches, trainer.py:56-60):

    dv       = 2 value_coeff alive (v - R)
    dlogit^m = -A alive (onehot(action_m) - p^m) + entr * p^m * (logp^m + H^m),   H^m = -sum p^m logp^m
    dh      += dv w_v + sum_m dlogit^m W_m
    do = dh tanh(c') ;  dc += dh sig(a_o) (1 - tanh(c')^2)
    da_o = do sig(a_o)(1-sig(a_o)); da_f = dc c sig(a_f)(1-sig(a_f)); da_i = dc tanh(a_g) sig(a_i)(1-sig(a_i));
    da_g = dc sig(a_i)(1 - tanh(a_g)^2);      dc_prev = dc sig(a_f)
    du = da W_ih ;  dh_prev = da W_hh ;  dx = du ;  dS = du W_c
    dh_prev_j += g_j * scale * sum_{k != j} g_k dS_k          (the transpose of the gated mean)
    parameter gradients: outer products of (da, u), (da, h), (du, S), (dx, obs), (dlogit, h'), (dv, h') + bias sums
"""
import numpy as np

from .grad import returns_np


def _sig(z):
    return 1.0 / (1.0 + np.exp(-z))


def _forward_step(p, obs, h, c, g, scale, comm_mask_zero, nheads):
    x = obs @ p["encoder.weight"].T + p["encoder.bias"]
    n = h.shape[0]
    if comm_mask_zero:
        M = np.zeros((n, n))
    else:
        M = (1.0 - np.eye(n)) * g[:, None] * g[None, :] * scale            # [dst k, src j]
    S = M @ h
    u = x + S @ p["C_modules.0.weight"].T + p["C_modules.0.bias"]
    a = u @ p["f_module.weight_ih"].T + p["f_module.bias_ih"] + h @ p["f_module.weight_hh"].T + p["f_module.bias_hh"]
    H = h.shape[1]
    si, sf, tg, so = _sig(a[:, :H]), _sig(a[:, H:2 * H]), np.tanh(a[:, 2 * H:3 * H]), _sig(a[:, 3 * H:])
    c2 = sf * c + si * tg
    tc = np.tanh(c2)
    h2 = so * tc
    v = (h2 @ p["value_head.weight"].T + p["value_head.bias"])[:, 0]
    logps = []
    for m in range(nheads):
        z = h2 @ p["heads.%d.weight" % m].T + p["heads.%d.bias" % m]
        z = z - z.max(-1, keepdims=True)
        logps.append(z - np.log(np.exp(z).sum(-1, keepdims=True)))
    cache = dict(obs=obs, h=h, c=c, M=M, S=S, u=u, si=si, sf=sf, tg=tg, so=so, tc=tc, h2=h2)
    return logps, v, h2, c2, cache


def compute_grad_manual(params_np, episodes, args):
    """Same inputs / outputs as oracle.grad.compute_grad (grads dict of float64 arrays, stat dict)."""
    p = {k: np.asarray(v, dtype=np.float64) for k, v in params_np.items()}
    hard = bool(args.hard_attn) and bool(args.commnet)
    nheads = sum(1 for q in p if q.startswith("heads.") and q.endswith(".weight"))
    avg = getattr(args, "comm_mode", "avg") == "avg"
    # ---- forward over the batch, keeping what the backward pass needs --------------------------------
    steps = []                       # per global step: (cache, logps, v, episode index, t inside the episode)
    for e_idx, ep in enumerate(episodes):
        n, H = ep["h"].shape[1], ep["h"].shape[2]
        h, c = np.zeros((n, H)), np.zeros((n, H))
        for t in range(ep["num_steps"]):
            alive = np.ones(n) if t == 0 else np.asarray(ep["alive_in"][t], dtype=np.float64)
            g = alive * (np.asarray(ep["comm_in"][t], dtype=np.float64) if hard else 1.0)
            n_alive = alive.sum()
            scale = 1.0 / (n_alive - 1) if (avg and n_alive > 1) else 1.0
            logps, v, h, c, cache = _forward_step(p, np.asarray(ep["obs"][t], dtype=np.float64), h, c, g, scale,
                                                  bool(args.comm_mask_zero), nheads)
            steps.append((cache, logps, v, e_idx, t))
    values = np.stack([s[2] for s in steps])
    reward = np.concatenate([ep["reward"] for ep in episodes])
    emask = np.concatenate([ep["emask"] for ep in episodes])
    mini = np.concatenate([ep["mini"] for ep in episodes])
    action = np.concatenate([ep["act"] for ep in episodes])
    alive_m = np.concatenate([ep["alive"] for ep in episodes]).astype(np.float64)
    ret = returns_np(reward, emask, mini, args.gamma, args.mean_ratio)
    adv = ret - values
    if args.normalize_rewards:
        adv = (adv - adv.mean()) / adv.std(ddof=1)                                  # torch.std is unbiased
    # ---- losses (stat) -------------------------------------------------------------------------------
    lp = np.zeros_like(values)
    entropy = 0.0
    for i, (cache, logps, v, _, _) in enumerate(steps):
        for m in range(nheads):
            lp[i] += np.take_along_axis(logps[m], action[i][:, m:m + 1], axis=-1)[:, 0]
            entropy -= (logps[m] * np.exp(logps[m])).sum()
    action_loss = (-adv * lp * alive_m).sum()
    value_loss = (((values - ret) ** 2) * alive_m).sum()
    # ---- backward ------------------------------------------------------------------------------------
    G = {k: np.zeros_like(v) for k, v in p.items()}
    W_ih, W_hh, W_c = p["f_module.weight_ih"], p["f_module.weight_hh"], p["C_modules.0.weight"]
    dh_next = dc_next = None
    for i in reversed(range(len(steps))):
        cache, logps, v, e_idx, t = steps[i]
        n, H = cache["h"].shape
        last_of_episode = (i == len(steps) - 1) or steps[i + 1][3] != e_idx
        detached = (t + 1) % args.detach_gap == 0                                   # trainer.py:56-60
        if last_of_episode or detached or dh_next is None:
            dh, dc = np.zeros((n, H)), np.zeros((n, H))
        else:
            dh, dc = dh_next, dc_next
        h2 = cache["h2"]
        # heads and value
        dv = 2.0 * args.value_coeff * alive_m[i] * (v - ret[i])
        G["value_head.weight"] += (dv[:, None] * h2).sum(0, keepdims=True)
        G["value_head.bias"] += dv.sum(keepdims=True)
        dh = dh + dv[:, None] * p["value_head.weight"]
        for m in range(nheads):
            pm = np.exp(logps[m])
            onehot = np.zeros_like(pm)
            np.put_along_axis(onehot, action[i][:, m:m + 1], 1.0, axis=-1)
            dlogit = (-adv[i] * alive_m[i])[:, None] * (onehot - pm)
            if args.entr > 0:
                Hm = -(pm * logps[m]).sum(-1, keepdims=True)
                dlogit = dlogit + args.entr * pm * (logps[m] + Hm)
            G["heads.%d.weight" % m] += dlogit.T @ h2
            G["heads.%d.bias" % m] += dlogit.sum(0)
            dh = dh + dlogit @ p["heads.%d.weight" % m]
        # LSTM cell
        si, sf, tg, so, tc = cache["si"], cache["sf"], cache["tg"], cache["so"], cache["tc"]
        do = dh * tc
        dc = dc + dh * so * (1.0 - tc * tc)
        da = np.concatenate([dc * tg * si * (1.0 - si), dc * cache["c"] * sf * (1.0 - sf), dc * si * (1.0 - tg * tg),
                             do * so * (1.0 - so)], axis=1)
        dc_prev = dc * sf
        G["f_module.weight_ih"] += da.T @ cache["u"]
        G["f_module.weight_hh"] += da.T @ cache["h"]
        G["f_module.bias_ih"] += da.sum(0)
        G["f_module.bias_hh"] += da.sum(0)
        du = da @ W_ih
        dh_prev = da @ W_hh
        # communication and encoder
        G["C_modules.0.weight"] += du.T @ cache["S"]
        G["C_modules.0.bias"] += du.sum(0)
        dS = du @ W_c
        dh_prev = dh_prev + cache["M"].T @ dS
        G["encoder.weight"] += du.T @ cache["obs"]
        G["encoder.bias"] += du.sum(0)
        dh_next, dc_next = dh_prev, dc_prev
    for k in ("hidd_encoder.weight", "hidd_encoder.bias"):                          # unused by the forward: no grad
        if k in G:
            G[k] = None
    stat = dict(action_loss=float(action_loss), value_loss=float(value_loss), entropy=float(entropy),
                num_steps=int(values.shape[0]))
    return G, stat
