"""CPU tests (gloo, world_size 2) of the data-parallel reduction that replaces the reference's
multi_processing.py: gradient = sum over ranks / GLOBAL num_steps, stats merged like merge_stat,
replicas identical after the optimizer step, exactly one gradient collective per update."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn, optim


class FakeTrainer(object):
    """Stands in for ic3net_b200.trainer.Trainer: deterministic per-rank 'rollout' and gradient."""

    def __init__(self, rank):
        torch.manual_seed(0)                       # identical replicas
        self.net = nn.ModuleDict(dict(used=nn.Linear(4, 3), unused=nn.Linear(3, 3)))   # 'unused' gets no grad
        self.params = [p for p in self.net.parameters()]
        self.optimizer = optim.RMSprop(self.net.parameters(), lr=0.01, alpha=0.97, eps=1e-6)
        self.rank = rank

    def run_batch(self, epoch):
        steps = 500 + 60 * self.rank
        return ("batch", self.rank), dict(num_steps=steps, num_episodes=7 + self.rank,
                                          reward=np.arange(3, dtype=np.float64) * (self.rank + 1), success=self.rank)

    def compute_grad(self, batch):
        g = torch.Generator().manual_seed(100 + self.rank)
        x = torch.randn(16, 4, generator=g)
        loss = self.net['used'](x).pow(2).sum() * (self.rank + 1)
        loss.backward()
        return dict(action_loss=float(loss.item()), value_loss=1.5 * (self.rank + 1), entropy=0.25)

    def state_dict(self):
        return self.optimizer.state_dict()


def expected(world):
    """Single-process restatement of multi_processing.py:74-98 over `world` workers."""
    trs = [FakeTrainer(r) for r in range(world)]
    stat, grads = {}, None
    from ic3net_b200.utils import merge_stat
    for t in trs:
        b, s = t.run_batch(0)
        t.optimizer.zero_grad(set_to_none=False)
        merge_stat(t.compute_grad(b), s)
        merge_stat(s, stat)
        gs = [p.grad.clone() for p in t.params if p.grad is not None]
        grads = gs if grads is None else [a + b_ for a, b_ in zip(grads, gs)]
    grads = [g / stat['num_steps'] for g in grads]
    master = trs[0]
    for p, g in zip([p for p in master.params if p.grad is not None], grads):
        p.grad.copy_(g)
    master.optimizer.step()
    return stat, [p.detach().clone() for p in master.params]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ic3net_b200.multi_gpu import MultiGPUTrainer
    import argparse
    mt = MultiGPUTrainer(argparse.Namespace(random=False), lambda: FakeTrainer(rank))
    stat = mt.train_batch(0)
    q.put((rank, {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in stat.items()},
           [p.detach().numpy().copy() for p in mt.trainer.params], mt.collectives))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_reduction_matches_reference_semantics():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda x: x[0])
    [p.join(timeout=60) for p in ps]
    stat, params = expected(world)
    for rank, s, ps_, ncoll in res:
        assert ncoll == 1
        assert s['num_steps'] == stat['num_steps'] == 500 + 560
        assert s['num_episodes'] == stat['num_episodes'] and s['success'] == stat['success']
        assert np.allclose(s['reward'], stat['reward'])
        assert np.isclose(s['action_loss'], stat['action_loss']) and np.isclose(s['value_loss'], stat['value_loss'])
        for a, b in zip(ps_, params):
            assert np.allclose(a, b.numpy(), rtol=1e-6, atol=1e-7)
    for a, b in zip(res[0][2], res[1][2]):
        assert np.array_equal(a, b)                 # replicas bit-identical


def test_single_process_is_plain_train_batch():
    from ic3net_b200.multi_gpu import MultiGPUTrainer
    import argparse
    mt = MultiGPUTrainer(argparse.Namespace(random=False), lambda: FakeTrainer(0))
    stat = mt.train_batch(0)
    ref_stat, ref_params = expected(1)
    assert stat['num_steps'] == 500 and mt.collectives == 0
    for p, b in zip(mt.trainer.params, ref_params):
        assert torch.allclose(p.detach(), b, rtol=1e-6, atol=1e-7)


def test_stat_packing_roundtrip():
    from ic3net_b200.multi_gpu import pack_stat, unpack_stat
    s = dict(num_steps=123456789, reward=np.array([1.5, -2.0]), success=3, add_rate=0.05, name="skip-me")
    vec, shapes = pack_stat(s, torch.device("cpu"))
    out = unpack_stat(vec, shapes, {})
    assert out['num_steps'] == 123456789 and out['success'] == 3 and out['add_rate'] == 0.05
    assert np.array_equal(out['reward'], s['reward']) and 'name' not in out


class FlatOptCPU(object):
    """CPU stand-in for ic3net_b200.optim.FlatRMSprop (same surface: flat_grads, zero_grad, step(grad_div)); the
    update is the torch formula of oracle/optim.py so the world-size-2 result can be compared with `expected`."""

    def __init__(self, params, lr, alpha=0.97, eps=1e-6):
        self.params = list(params)
        self.lr, self.alpha, self.eps = lr, alpha, eps
        self._off, n = [], 0
        for p in self.params:
            self._off.append(n)
            n += p.numel()
        self.flat_grads = torch.zeros(n)
        self.flat_sq = torch.zeros(n)
        self.seen = torch.zeros(n, dtype=torch.bool)
        for p, off in zip(self.params, self._off):
            p.grad = self.flat_grads[off:off + p.numel()].view_as(p)

    def zero_grad(self, set_to_none=False):
        self.flat_grads.zero_()

    def step(self, grad_div=1.0):
        g = self.flat_grads / grad_div
        self.flat_sq.mul_(self.alpha).addcmul_(g, g, value=1 - self.alpha)
        upd = self.lr * g / (self.flat_sq.sqrt() + self.eps)
        with torch.no_grad():
            for p, off in zip(self.params, self._off):
                p.sub_(upd[off:off + p.numel()].view_as(p))


class FlatFakeTrainer(FakeTrainer):
    def __init__(self, rank):
        super().__init__(rank)
        self.optimizer = FlatOptCPU(self.net.parameters(), lr=0.01)


def _flat_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ic3net_b200.multi_gpu import MultiGPUTrainer
    import argparse
    mt = MultiGPUTrainer(argparse.Namespace(random=False), lambda: FlatFakeTrainer(rank))
    stat = mt.train_batch(0)
    q.put((rank, int(stat['num_steps']), [p.detach().numpy().copy() for p in mt.trainer.params], mt.collectives))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_flat_gradient_buffer():
    """The FlatRMSprop fast path of MultiGPUTrainer: the flat gradient buffer is all-reduced in place, the division
    by the GLOBAL step count happens inside optimizer.step(grad_div) -- same result as multi_processing.py:90-97."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_flat_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda x: x[0])
    [p.join(timeout=60) for p in ps]
    stat, params = expected(world)
    for rank, nsteps, ps_, ncoll in res:
        assert ncoll == 1 and nsteps == stat['num_steps']
        for a, b in zip(ps_, params):
            assert np.allclose(a, b.numpy(), rtol=1e-6, atol=1e-7)
    for a, b in zip(res[0][2], res[1][2]):
        assert np.array_equal(a, b)


class SkewedTrainer(FakeTrainer):
    """Every rank builds its policy from a DIFFERENT seed (what an unseeded `--seed -1` launch used to do)."""

    def __init__(self, rank):
        super().__init__(rank)
        torch.manual_seed(1000 + rank)
        with torch.no_grad():
            for p in self.params:
                p.copy_(torch.randn_like(p))


def _skew_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ic3net_b200.multi_gpu import MultiGPUTrainer
    import argparse
    tr = SkewedTrainer(rank)
    before = [p.detach().numpy().copy() for p in tr.params]
    mt = MultiGPUTrainer(argparse.Namespace(random=False), lambda: tr)
    after = [p.detach().numpy().copy() for p in mt.trainer.params]
    diff0 = mt.replica_checksum()
    mt.train_batch(0)
    q.put((rank, before, after, diff0, mt.replica_checksum()))
    dist.barrier()
    dist.destroy_process_group()


def test_replicas_take_rank0_parameters_at_construction():
    """The reference's workers share ONE policy (main.py:177-179); ranks that start from different parameters must
    be made identical by MultiGPUTrainer itself (broadcast from rank 0) and stay identical after an update."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_skew_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda x: x[0])
    [p.join(timeout=60) for p in ps]
    (_, b0, a0, d0, e0), (_, b1, a1, d1, e1) = res
    assert any(not np.array_equal(x, y) for x, y in zip(b0, b1))       # they really started apart
    for x, y, z in zip(a0, a1, b0):
        assert np.array_equal(x, y) and np.array_equal(x, z)          # both now hold rank 0's values
    assert d0 == d1 == 0.0 and e0 == e1 == 0.0
