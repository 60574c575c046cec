"""Batched GPU rollout behind the reference's ``trainer.Trainer`` surface (trainer.py:14-262).

``Trainer(args, policy_net, env)`` drives ``env.nenvs`` independent environment slots in
lock-step on one GPU.  Each slot plays the role of one reference process: it runs
episode after episode (auto-reset, hidden state zeroed, nobody talks at t = 0), and
``run_batch`` returns once every slot has produced ``>= batch_size`` steps
(``ceil(batch_size / max_steps) * max_steps`` lock-step iterations; an episode still
open at the end of the batch is cut there, which is the only deviation from
trainer.py:231-237, where the last episode may overshoot instead).

One lock-step iteration is 3 kernel launches and no host synchronisation:
  encoder (index form from the env state, or obs-gather + dense encoder)
  -> policy step (comm mean, C, LSTM, heads, sampling)
  -> env step + Trainer.get_episode bookkeeping + auto-reset (ic3_rollout_io).
The whole T-step sequence can be captured once into a CUDA graph (``use_graph``).
"""
import ctypes as C
import math
from collections import namedtuple

import numpy as np
import torch
from torch import optim

from . import _lib
from .utils import merge_stat

Transition = namedtuple('Transition', ('state', 'action', 'action_out', 'value', 'episode_mask',
                                       'episode_mini_mask', 'next_state', 'reward', 'misc'))

RolloutBatch = namedtuple('RolloutBatch', ('action', 'logp', 'value', 'reward', 'episode_mask',
                                           'episode_mini_mask', 'alive_mask', 'snapshot'))


class Trainer(object):
    def __init__(self, args, policy_net, env):
        self.args = args
        self.policy_net = policy_net
        self.env = env                       # GymWrapper
        self.display = False
        self.last_step = False
        self.optimizer = optim.RMSprop(policy_net.parameters(), lr=args.lrate, alpha=0.97, eps=1e-6)
        self.params = [p for p in self.policy_net.parameters()]
        self.obs_mode = getattr(args, 'obs_mode', 'index')      # 'index' | 'dense'
        self.use_graph = bool(getattr(args, 'use_graph', False))
        self.is_tj = args.env_name == 'traffic_junction'
        self._buf = None
        self._graph = None
        self.launches_per_step = 3 if self.obs_mode == 'index' else 4

    # ------------------------------------------------------------------ buffers
    def _alloc(self, T):
        e = self.env.env
        B, N, H = e.nenvs, self.args.nagents, self.args.hid_size
        dev = e.device
        nh = len(self.args.naction_heads)
        A = sum(self.args.naction_heads)
        z = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype, device=dev)
        b = dict(T=T, h=z(B * N, H), c=z(B * N, H), x=z(B * N, H),
                 comm=z(B, N, dtype=torch.uint8), alive=torch.ones(B, N, dtype=torch.uint8, device=dev),
                 fresh=torch.ones(B, dtype=torch.uint8, device=dev), t_ep=z(B, dtype=torch.int32),
                 action=z(T, B, N, nh, dtype=torch.int32), logp=z(T, B, N, A), value=z(T, B * N),
                 reward=z(T, B, N), emask=z(T, B, dtype=torch.uint8), mini=z(T, B, N, dtype=torch.uint8),
                 ralive=z(T, B, N, dtype=torch.uint8), step_reward=z(B, N),
                 stat_reward=z(B, N), stat_comm=z(B, N), stat_success=z(B, dtype=torch.int32),
                 stat_episodes=z(B, dtype=torch.int32), stat_steps=z(B, dtype=torch.int32),
                 err=z(1, dtype=torch.int32))
        if self.obs_mode == 'dense':
            b['obs'] = torch.empty(B, N, self.env.observation_dim, dtype=torch.float32, device=dev)
        if self.is_tj:
            b['snap_obs'] = None
        self._buf = b
        self._graph = None
        return b

    # ------------------------------------------------------------------ rollout
    def _enqueue(self, T):
        """Enqueue T lock-step iterations on the current stream (no host sync)."""
        b, e, net, args = self._buf, self.env.env, self.policy_net, self.args
        lib = _lib.load()
        B, N = e.nenvs, args.nagents
        nh = len(args.naction_heads)
        cfg = net.policy_cfg(B)
        cfg.seed, cfg.env_id0 = e.cfg.seed, e.cfg.env_id0
        w = net.packed()
        hard = int(bool(args.hard_attn) and bool(args.commnet))
        s = _lib.stream()
        ws, _ = net.workspace(B)          # tensor-core path scratch (None for the fp32 SIMT kernel)
        for t in range(T):
            if self.obs_mode == 'dense':
                if self.is_tj:
                    _lib.check(lib.ic3_tj_obs(C.byref(e.cfg), C.byref(e.state), b['obs'].data_ptr(), s))
                else:
                    _lib.check(lib.ic3_pp_obs(C.byref(e.cfg), C.byref(e.state), b['obs'].data_ptr(), s))
                _lib.check(lib.ic3_encoder_dense(C.byref(cfg), C.byref(w), b['obs'].data_ptr(), b['x'].data_ptr(), s))
            elif self.is_tj:
                _lib.check(lib.ic3_tj_encoder_index(C.byref(e.cfg), C.byref(e.state), C.byref(cfg), C.byref(w),
                                                    b['x'].data_ptr(), s))
            else:
                _lib.check(lib.ic3_pp_encoder_index(C.byref(e.cfg), C.byref(e.state), C.byref(cfg), C.byref(w),
                                                    b['x'].data_ptr(), s))
            io = _lib.PolicyIO(x=b['x'].data_ptr(), h=b['h'].data_ptr(), c=b['c'].data_ptr(),
                               comm_action=b['comm'].data_ptr() if hard else None, alive=b['alive'].data_ptr(),
                               fresh=b['fresh'].data_ptr(), tick=e.tick.data_ptr(), draws=None,
                               h_out=b['h'].data_ptr(), c_out=b['c'].data_ptr(), value=b['value'][t].data_ptr(),
                               logp=b['logp'][t].data_ptr(), action=b['action'][t].data_ptr(),
                               workspace=_lib.ptr(ws), err=b['err'].data_ptr())
            _lib.check(lib.ic3_policy_step(C.byref(cfg), C.byref(w), C.byref(io), s))
            r = _lib.RolloutIO(t=t, max_steps=args.max_steps, nheads=nh, hard_attn=hard,
                               comm_action_one=int(bool(args.comm_action_one)), last=int(t == T - 1),
                               action=b['action'][t].data_ptr(), t_ep=b['t_ep'].data_ptr(),
                               fresh=b['fresh'].data_ptr(), comm_next=b['comm'].data_ptr(),
                               alive_next=b['alive'].data_ptr(), rec_reward=b['reward'].data_ptr(),
                               rec_episode_mask=b['emask'].data_ptr(), rec_mini_mask=b['mini'].data_ptr(),
                               rec_alive=b['ralive'].data_ptr(), stat_reward=b['stat_reward'].data_ptr(),
                               stat_comm=b['stat_comm'].data_ptr(), stat_success=b['stat_success'].data_ptr(),
                               stat_episodes=b['stat_episodes'].data_ptr(), stat_steps=b['stat_steps'].data_ptr())
            if self.is_tj:
                _lib.check(lib.ic3_tj_step(C.byref(e.cfg), C.byref(e.state), b['action'][t].data_ptr(), nh, None,
                                           b['step_reward'].data_ptr(), None, b['err'].data_ptr(), C.byref(r), s))
            else:
                _lib.check(lib.ic3_pp_step(C.byref(e.cfg), C.byref(e.state), b['action'][t].data_ptr(), nh,
                                           b['step_reward'].data_ptr(), None, b['err'].data_ptr(), C.byref(r), s))

    def rollout(self, T, epoch=0):
        """T lock-step iterations from fresh episodes in every slot.  Returns (RolloutBatch, stat);
        everything stays on the device except the small stat reductions."""
        e = self.env.env
        if self._buf is None or self._buf['T'] != T:
            self._alloc(T)
        b = self._buf
        # episode boundary for every slot (trainer.py:28-32, 45-51)
        if self.is_tj:
            e.reset(epoch, want_obs=False)
        else:
            e.reset(want_obs=False)
        for k in ('stat_reward', 'stat_comm', 'stat_success', 'stat_episodes', 'stat_steps', 't_ep', 'err'):
            b[k].zero_()
        b['fresh'].fill_(1)
        self.policy_net.packed()                # (re)pack weights outside any graph capture
        if self.use_graph:
            if self._graph is None:
                self._enqueue(T)                # warm-up (lazy function attributes, allocator)
                torch.cuda.synchronize()
                if self.is_tj:
                    e.reset(epoch, want_obs=False)
                else:
                    e.reset(want_obs=False)
                for k in ('stat_reward', 'stat_comm', 'stat_success', 'stat_episodes', 'stat_steps', 't_ep'):
                    b[k].zero_()
                b['fresh'].fill_(1)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(T)
                self._graph = g
                g.replay()
            else:
                self._graph.replay()
        else:
            self._enqueue(T)
        batch = RolloutBatch(action=b['action'], logp=b['logp'], value=b['value'].view(T, e.nenvs, -1),
                             reward=b['reward'], episode_mask=b['emask'], episode_mini_mask=b['mini'],
                             alive_mask=b['ralive'], snapshot=None)
        return batch

    def collect_stat(self):
        """Host-side stat dict with the reference's keys (trainer.py:73-75,86-88,109-110,124-125),
        summed over the env slots of this GPU."""
        b, e, args = self._buf, self.env.env, self.args
        if int(b['err'].item()):
            raise RuntimeError("device-side error flag %d during rollout" % int(b['err'].item()))
        stat = dict()
        stat['num_episodes'] = int(b['stat_episodes'].sum().item())
        stat['num_steps'] = int(b['stat_steps'].sum().item())
        stat['steps_taken'] = stat['num_steps']
        stat['reward'] = b['stat_reward'].sum(0).double().cpu().numpy()
        if args.hard_attn and args.commnet:
            stat['comm_action'] = b['stat_comm'].sum(0).double().cpu().numpy()
        if not (not self.is_tj and args.mode == 'competitive'):
            stat['success'] = int(b['stat_success'].sum().item())
        if self.is_tj:
            stat['add_rate'] = e.add_rate * stat['num_episodes']
        return stat

    # ------------------------------------------------------------------ reference surface
    def get_episode(self, epoch):
        """One episode horizon (max_steps lock-step iterations) for every env slot."""
        batch = self.rollout(self.args.max_steps, epoch)
        return batch, self.collect_stat()

    def steps_per_batch(self):
        return int(math.ceil(self.args.batch_size / float(self.args.max_steps))) * self.args.max_steps

    def run_batch(self, epoch):
        batch = self.rollout(self.steps_per_batch(), epoch)
        self.stats = self.collect_stat()
        return batch, self.stats

    def compute_grad(self, batch):
        raise NotImplementedError("REINFORCE gradient (trainer.py:128-225) is the next row of the scope table")

    def train_batch(self, epoch):
        batch, stat = self.run_batch(epoch)
        self.optimizer.zero_grad()
        s = self.compute_grad(batch)
        merge_stat(s, stat)
        for p in self.params:
            if p._grad is not None:
                p._grad.data /= stat['num_steps']
        self.optimizer.step()
        return stat

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, state):
        self.optimizer.load_state_dict(state)
