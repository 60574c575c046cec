"""CommNet / IC3Net policy step, CPU restatement (numpy float64).  TEST INFRASTRUCTURE.

Follows ``comm.py`` of the reference (recurrent / LSTM branch, comm_passes = 1,
the only branch any BASELINE config uses):
  forward_state_encoder :114-131  x = encoder(obs)        (no tanh when recurrent)
  get_agent_mask        :99-112   alive mask, n_alive counted BEFORE gating
  hard attention        :171-175  g = alive * comm_action
  comm                  :181-205  S[k] = sum_{j != k} h[j] / (n_alive-1) * g[j] * g[k]
  C + skip              :206,211  inp = x + C(S)           (bias added even when S = 0)
  LSTMCell              :213-218  torch.nn.LSTMCell, gate order i, f, g, o
  heads                 :228-239  value_head(h'), log_softmax(head(h'))
and ``action_utils.select_action`` (:32-36) with torch.multinomial replaced by
inverse-CDF sampling from a supplied uniform.

Parameters use the reference ``state_dict`` key names.
"""
import numpy as np

from . import philox


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _log_softmax(z):
    m = z.max(axis=-1, keepdims=True)
    return z - m - np.log(np.exp(z - m).sum(axis=-1, keepdims=True))


def params_to_f64(sd):
    return {k: np.asarray(v, dtype=np.float64) for k, v in sd.items()}


def nheads(params):
    k = 0
    while "heads.%d.weight" % k in params:
        k += 1
    return k


def comm_sum(h, g, n_alive, comm_mode="avg", comm_mask_zero=False):
    """S[k] = sum_{j != k} (h[j] / (n_alive-1)) * g[j] * g[k]   (comm.py:181-205)."""
    n = h.shape[0]
    S = np.zeros_like(h)
    if comm_mask_zero:
        return S
    scale = 1.0
    if comm_mode == "avg" and n_alive > 1:
        scale = 1.0 / (n_alive - 1)
    for k in range(n):
        acc = np.zeros(h.shape[1])
        for j in range(n):
            if j != k:
                acc = acc + h[j] * scale * g[j] * g[k]
        S[k] = acc
    return S


def forward(params, obs, h, c, comm_action=None, alive=None, hard_attn=True,
            comm_mode="avg", comm_mask_zero=False):
    """One policy step for ONE environment.

    obs [N,O], h,c [N,H], comm_action [N] (0/1) or None, alive [N] (0/1) or None.
    Returns (logps: list of [N,na], value [N], h' [N,H], c' [N,H], x [N,H]).
    """
    p = params
    obs = np.asarray(obs, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    n, H = h.shape
    x = obs @ p["encoder.weight"].T + p["encoder.bias"]
    alive_v = np.ones(n) if alive is None else np.asarray(alive, dtype=np.float64)
    n_alive = alive_v.sum()
    g = alive_v.copy()
    if hard_attn:
        g = g * np.asarray(comm_action, dtype=np.float64)
    S = comm_sum(h, g, n_alive, comm_mode, comm_mask_zero)
    cvec = S @ p["C_modules.0.weight"].T + p["C_modules.0.bias"]
    inp = x + cvec
    gates = (inp @ p["f_module.weight_ih"].T + p["f_module.bias_ih"]
             + h @ p["f_module.weight_hh"].T + p["f_module.bias_hh"])
    gi, gf, gg, go = (gates[:, k * H:(k + 1) * H] for k in range(4))
    c2 = _sigmoid(gf) * c + _sigmoid(gi) * np.tanh(gg)
    h2 = _sigmoid(go) * np.tanh(c2)
    value = (h2 @ p["value_head.weight"].T + p["value_head.bias"])[:, 0]
    logps = []
    for k in range(nheads(p)):
        logps.append(_log_softmax(h2 @ p["heads.%d.weight" % k].T + p["heads.%d.bias" % k]))
    return logps, value, h2, c2, x


def roles_of(params, model="commnet", recurrent=True, passes=1):
    """Parameter arrays by kernel role for the reference's model families (comm.py:31-96, models.py:8-96)."""
    p = params
    heads = [(p["heads.%d.weight" % k], p["heads.%d.bias" % k]) for k in range(nheads(p))]
    r = dict(value=(p["value_head.weight"], p["value_head.bias"]), heads=heads)
    if model == "commnet":
        r["enc"] = (p["encoder.weight"], p["encoder.bias"])
        r["C"] = [(p["C_modules.%d.weight" % i], p["C_modules.%d.bias" % i]) for i in range(passes)]
        if recurrent:
            r["lstm"] = tuple(p["f_module." + k] for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))
        else:
            r["f"] = [(p["f_modules.%d.weight" % i], p["f_modules.%d.bias" % i]) for i in range(passes)]
    else:                                   # models.MLP / models.RNN: independent controllers, no comm
        H = p["value_head.weight"].shape[1]
        r["enc"] = (p["affine1.weight"], p["affine1.bias"])
        r["C"] = [(np.zeros((H, H)), np.zeros(H))]
        if "lstm_unit.weight_ih" in p:
            r["lstm"] = tuple(p["lstm_unit." + k] for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))
        else:
            r["f"] = [(p["affine2.weight"], p["affine2.bias"])]
    return r


def forward_variant(roles, obs, h, c, comm_action=None, alive=None, hard_attn=False, comm_mode="avg",
                    comm_mask_zero=False, passes=1, x_tanh=False, h_from_x=False):
    """Policy step of every variant (comm.py:134-244; models.py:22-25,68-85) for ONE environment:
      x = encoder(obs) [tanh'd in the non-recurrent branch]; hidden = x or the recurrent state;
      per pass: S = gated mean of hidden, c_i = C_i(S); LSTM: (h, c) = LSTMCell(x + c_i, (h, c));
      tanh cell: h = tanh(x + f_i(h) + c_i).   Returns (logps, value [N], h' [N,H], c' or None)."""
    obs = np.asarray(obs, dtype=np.float64)
    x = obs @ roles["enc"][0].T + roles["enc"][1]
    if x_tanh:
        x = np.tanh(x)
    hid = x if h_from_x else np.asarray(h, dtype=np.float64)
    n, H = hid.shape
    alive_v = np.ones(n) if alive is None else np.asarray(alive, dtype=np.float64)
    n_alive = alive_v.sum()
    g = alive_v.copy()
    if hard_attn:
        g = g * np.asarray(comm_action, dtype=np.float64)
    c2 = None if c is None else np.asarray(c, dtype=np.float64)
    for i in range(passes):
        S = comm_sum(hid, g, n_alive, comm_mode, comm_mask_zero)
        cvec = S @ roles["C"][i][0].T + roles["C"][i][1]
        if "lstm" in roles:
            w_ih, w_hh, b_ih, b_hh = roles["lstm"]
            gates = (x + cvec) @ w_ih.T + b_ih + hid @ w_hh.T + b_hh
            gi, gf, gg, go = (gates[:, k * H:(k + 1) * H] for k in range(4))
            c2 = _sigmoid(gf) * c2 + _sigmoid(gi) * np.tanh(gg)
            hid = _sigmoid(go) * np.tanh(c2)
        else:
            hid = np.tanh(x + hid @ roles["f"][i][0].T + roles["f"][i][1] + cvec)
    value = (hid @ roles["value"][0].T + roles["value"][1])[:, 0]
    logps = [_log_softmax(hid @ w.T + b) for w, b in roles["heads"]]
    return logps, value, hid, c2


def sample_from_logp(logp_row, u):
    """Inverse CDF: smallest a with sum_{i<=a} exp(logp_i) > u, else the last index.

    Returns (action, margin) where margin is the distance of u to the nearest CDF
    edge (used by parity tests to flag draws that fp32 rounding may flip).
    """
    pr = np.exp(np.asarray(logp_row, dtype=np.float64))
    cdf = np.cumsum(pr)
    a = len(pr) - 1
    for i in range(len(pr)):
        if cdf[i] > u:
            a = i
            break
    margin = float(np.min(np.abs(cdf[:-1] - u))) if len(pr) > 1 else 1.0
    return a, margin


def sample_actions(logps, u24):
    """logps: list over heads of [N,na]; u24 [N,heads] ints.  Returns (act [N,heads], margin [N,heads])."""
    n = logps[0].shape[0]
    act = np.zeros((n, len(logps)), dtype=np.int64)
    margin = np.ones((n, len(logps)))
    for k, lp in enumerate(logps):
        for i in range(n):
            act[i, k], margin[i, k] = sample_from_logp(lp[i], float(u24[i][k]) * 2.0 ** -24)
    return act, margin


def action_draws(seed, env_id, tick, n, heads):
    """u24 [N,heads] from the Philox ACTION stream (oracle/philox.py)."""
    out = np.zeros((n, heads), dtype=np.int64)
    for i in range(n):
        out[i] = philox.draw_u24(seed, env_id, tick, philox.STREAM_ACTION, i)[:heads]
    return out
