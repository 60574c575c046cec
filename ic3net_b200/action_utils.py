"""Counterpart of the reference ``action_utils.py`` (:5-63) for batched CUDA tensors.

``select_action`` replaces ``torch.multinomial`` (action_utils.py:35) by inverse-CDF
sampling in a CUDA kernel, driven by the library's Philox action stream or by
explicit 24-bit draws; ``translate_action`` keeps the reference's return shape
``(action, actual)`` = per-head lists, now of ``[B, N]`` int32 CUDA tensors.
"""
import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib


def parse_action_args(args):
    """Derive ``args.continuous`` / ``args.naction_heads`` (action_utils.py:5-25).  A discrete environment
    (``num_actions[0] > 0``) contributes one head per action dimension; otherwise ``--nactions`` decides: "1" means
    continuous control, "k" means ``dim_actions`` heads of k actions, "a:b:..." lists the heads explicitly."""
    if args.num_actions[0] > 0:
        args.continuous = False
        args.naction_heads = [int(args.num_actions[d]) for d in range(args.dim_actions)]
        return
    spec = [int(tok) for tok in args.nactions.split(':')]        # int('') raises ValueError like the reference
    if len(spec) > 1:
        args.continuous, args.naction_heads = False, spec
    elif spec[0] == 1:
        args.continuous = True
    elif spec[0] > 1:
        args.continuous, args.naction_heads = False, [spec[0]] * args.dim_actions
    else:
        raise RuntimeError("--nactions wrong format!")


_ticks = {}          # id(args) -> (weak reference to args, per-env call counter [B] int32)
_cfgs = {}          # sampler launch descriptors by (B, N, heads, env_id0, seed)


def _joined_view(action_out, heads):
    """The per-head log-prob tensors of CommNetMLP.forward are consecutive column slices of ONE contiguous float32
    buffer: hand that buffer to the sampling kernel instead of concatenating the slices again (None: not the case)."""
    first = action_out[0]
    if first.dtype != torch.float32 or first.dim() != 3 or not first.is_cuda:
        return None
    atot = sum(heads)
    B, N = first.shape[0], first.shape[1]
    if first.stride() != (N * atot, atot, 1):
        return None
    base, off = first.data_ptr(), 0
    for a, na in zip(action_out, heads):
        if a.dtype != torch.float32 or a.shape[:2] != (B, N) or a.stride() != (N * atot, atot, 1) or \
                a.data_ptr() != base + 4 * off:
            return None
        off += na
    return torch.as_strided(first, (B, N, atot), (N * atot, atot, 1))


def select_action(args, action_out, draws=None, tick=None):
    """action_out: list over heads of log-probs [B, N, na].  Returns int32 [B, N, heads].

    draws: optional explicit 24-bit uniforms [B, N, heads]; otherwise the Philox action
    stream (seed = args.seed) at ``tick`` ([B] int32 tensor; by default an internal
    per-args call counter)."""
    if args.continuous:
        raise NotImplementedError("continuous actions are outside the accelerated path")
    heads = [int(a.shape[-1]) for a in action_out]
    logp = _joined_view(action_out, heads)            # CommNetMLP.forward returns views of one [B, N, sum(na)] buffer
    if logp is None:
        logp = torch.cat([a.to(torch.float32) for a in action_out], dim=-1).contiguous()
    B, N = logp.shape[0], logp.shape[1]
    ckey = (B, N, tuple(heads), int(getattr(args, 'env_id0', 0)), int(getattr(args, 'seed', 0)))
    cfg = _cfgs.get(ckey)
    if cfg is None:
        hd = (C.c_int32 * _lib.MAX_HEADS)(*(heads + [0] * (_lib.MAX_HEADS - len(heads))))
        cfg = _cfgs[ckey] = _lib.PolicyCfg(B=B, N=N, H=32, O=1, nheads=len(heads), head_dim=hd, hard_attn=0,
                                           comm_avg=0, comm_mask_zero=0, env_id0=ckey[3],
                                           seed=ckey[4] & 0xFFFFFFFFFFFFFFFF)
    d, bump = None, False
    if draws is not None:
        d = torch.as_tensor(np.asarray(draws.cpu() if torch.is_tensor(draws) else draws, dtype=np.int64))
        d = d.to(logp.device, torch.int32).reshape(B, N, len(heads)).contiguous()
    elif tick is None:
        ent = _ticks.get(id(args))
        if ent is None or ent[0]() is not args or ent[1].shape[0] != B or ent[1].device != logp.device:
            try:
                ref = weakref.ref(args)          # a recycled id() of a dead namespace must not inherit its counter
            except TypeError:
                ref = (lambda a=args: a)
            ent = _ticks[id(args)] = (ref, torch.zeros(B, dtype=torch.int32, device=logp.device))
        tick, bump = ent[1], True
    action = torch.empty(B, N, len(heads), dtype=torch.int32, device=logp.device)
    _lib.check(_lib.load().ic3_sample_actions(C.byref(cfg), logp.data_ptr(), _lib.ptr(tick), _lib.ptr(d),
                                              action.data_ptr(), _lib.stream()))
    if bump:
        tick += 1                 # after the sampler on the same stream: the kernel read the old count
    return action


def translate_action(args, env, action):
    # action_utils.py:39-43 (discrete branch): per-head arrays; `actual` is what the env gets
    if args.num_actions[0] > 0:
        action = [action[..., k].contiguous() for k in range(action.shape[-1])]
        actual = action
        return action, actual
    raise NotImplementedError("continuous / scaled actions are outside the accelerated path")
