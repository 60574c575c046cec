"""RMSprop step + checkpoint format: the fused flat-buffer kernel against torch.optim.RMSprop
fixtures, the trainer's update through it, and save / load interchange with the reference's formats."""
import io

import numpy as np
import pytest
import torch

from helpers import load_golden

pytestmark = pytest.mark.gpu


def _fixture_params(z, n, dev):
    return [torch.nn.Parameter(torch.from_numpy(z["p0_%d" % i]).float().to(dev)) for i in range(n)]


def test_flat_rmsprop_matches_torch_fixture():
    from ic3net_b200.optim import FlatRMSprop
    meta, z = load_golden("rmsprop_ref")
    n, dev = len(meta["shapes"]), torch.device("cuda")
    params = _fixture_params(z, n, dev)
    opt = FlatRMSprop(params, lr=meta["lr"], alpha=meta["alpha"], eps=meta["eps"])
    for u, ns_ in enumerate(meta["num_steps"]):
        opt.zero_grad()
        v0 = [p._version for p in params]
        for i in range(n):
            if meta["live"][i]:
                params[i].grad.add_(torch.from_numpy(z["g%d_%d" % (u, i)]).float().to(dev))   # accumulate like autograd
        opt.step(grad_div=ns_)
        torch.cuda.synchronize()
        for i in range(n):
            want = z["p%d_%d" % (u + 1, i)]
            got = params[i].detach().cpu().numpy().astype(np.float64)
            assert np.allclose(got, want, rtol=2e-6, atol=2e-7), (u, i, np.abs(got - want).max())
            assert params[i]._version > v0[i]                      # weight caches keyed on _version see the update
            if meta["live"][i]:                                    # p.grad holds grad / num_steps (trainer.py:251-253)
                g = params[i].grad.cpu().numpy().astype(np.float64)
                assert np.allclose(g, z["g%d_%d" % (u, i)] / ns_, rtol=1e-6, atol=1e-30)
    sd = opt.state_dict()
    for i in range(n):
        if meta["live"][i]:
            assert np.allclose(sd["state"][i]["square_avg"].cpu().numpy(), z["v_%d" % i], rtol=2e-6, atol=1e-30)
        else:
            assert np.array_equal(params[i].detach().cpu().numpy(), z["p0_%d" % i].astype(np.float32))
            assert float(sd["state"][i]["square_avg"].abs().max()) == 0.0


def test_state_dict_interchanges_with_torch_rmsprop():
    """Our optimizer state loads into torch.optim.RMSprop (what a reference checkpoint holds) and back."""
    from ic3net_b200.optim import FlatRMSprop
    meta, z = load_golden("rmsprop_ref")
    n, dev = len(meta["shapes"]), torch.device("cuda")
    pa, pb = _fixture_params(z, n, dev), _fixture_params(z, n, dev)
    ours = FlatRMSprop(pa, lr=meta["lr"], alpha=meta["alpha"], eps=meta["eps"])
    ref = torch.optim.RMSprop(pb, lr=meta["lr"], alpha=meta["alpha"], eps=meta["eps"])
    g = torch.Generator().manual_seed(5)
    grads = [[torch.randn(*s, generator=g).to(dev) for s in meta["shapes"]] for _ in range(4)]
    for u in range(2):                                   # two updates on both
        ours.zero_grad()
        ref.zero_grad(set_to_none=False)
        for i in range(n):
            pa[i].grad.add_(grads[u][i])
            pb[i].grad = grads[u][i].clone()
        ours.step()
        ref.step()
    buf = io.BytesIO()
    torch.save(ours.state_dict(), buf)                   # through the serialised form, like main.py's checkpoints
    buf.seek(0)
    ref2 = torch.optim.RMSprop(pb, lr=0.5, alpha=0.5, eps=1e-3)
    ref2.load_state_dict(torch.load(buf, weights_only=False))
    ours2 = FlatRMSprop(pa, lr=0.5, alpha=0.5, eps=1e-3)
    ours2.load_state_dict(ref.state_dict())              # a torch / reference checkpoint into ours
    assert (ours2.lr, ours2.alpha, ours2.eps) == (meta["lr"], meta["alpha"], meta["eps"])
    for u in range(2, 4):                                # continue on both: trajectories must agree
        ours2.zero_grad()
        ref2.zero_grad(set_to_none=False)
        for i in range(n):
            pa[i].grad.add_(grads[u][i])
            pb[i].grad = grads[u][i].clone()
        ours2.step()
        ref2.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)


def test_rmsprop_argument_checks():
    from ic3net_b200 import _lib
    lib = _lib.load()
    t = torch.zeros(64, device="cuda")
    p = t.data_ptr()
    assert lib.ic3_rmsprop_step(0, 1e-3, 0.97, 1e-6, 1.0, p, p, p, None) != 0          # n <= 0
    assert lib.ic3_rmsprop_step(8, 1e-3, 0.97, 1e-6, 0.0, p, p, p, None) != 0          # grad_div must be > 0
    assert lib.ic3_rmsprop_step(8, 1e-3, 0.97, 1e-6, 1.0, None, p, p, None) != 0       # NULL
    assert lib.ic3_rmsprop_step(8, 1e-3, 0.97, 1e-6, 1.0, p + 4, p, p, None) != 0      # misaligned
    # odd length: float4 body + scalar tail
    prm, g, v = torch.ones(7, device="cuda"), torch.full((7,), 2.0, device="cuda"), torch.zeros(7, device="cuda")
    n = 7
    flat = torch.zeros(24, device="cuda")
    flat[:7], flat[8:15] = prm, g
    assert lib.ic3_rmsprop_step(n, 0.1, 0.9, 1e-6, 2.0, flat[8:].data_ptr(), flat.data_ptr(), flat[16:].data_ptr(), None) == 0
    torch.cuda.synchronize()
    want_v = 0.1 * 1.0
    assert torch.allclose(flat[16:23], torch.full((7,), want_v, device="cuda"))
    assert torch.allclose(flat[:7], torch.full((7,), 1.0 - 0.1 * 1.0 / (want_v ** 0.5 + 1e-6), device="cuda"))
    assert torch.allclose(flat[8:15], torch.ones(7, device="cuda"))                     # divided gradient written back


def test_trainer_update_is_reference_rmsprop():
    """Trainer.train_batch's update = grad / num_steps followed by torch's RMSprop formula (trainer.py:245-256)."""
    from helpers import finish_args, ns
    from ic3net_b200 import data
    from ic3net_b200.comm import CommNetMLP
    from ic3net_b200.trainer import Trainer
    meta, _ = load_golden("grad_pp_easy_ic3net")
    args = ns(meta["args"], nenvs=6, seed=11, env_id0=0, obs_mode="index", use_graph=False, record_for_grad=True,
              grad_window=16, lrate=0.002)
    env = data.init(args.env_name, args)
    finish_args(args, env)
    torch.manual_seed(1)
    net = CommNetMLP(args, args.num_inputs)
    tr = Trainer(args, net, env)
    before = [p.detach().double().clone() for p in tr.params]
    # one update, step by step, keeping the summed gradient
    batch, stat = tr.run_batch(0)
    tr.optimizer.zero_grad()
    tr.compute_grad(batch)
    gsum = [p.grad.detach().double().clone() for p in tr.params]
    tr.optimizer.step(grad_div=stat["num_steps"])
    torch.cuda.synchronize()
    moved = 0
    for p, p0, g in zip(tr.params, before, gsum):
        g = g / stat["num_steps"]
        v = (1 - 0.97) * g * g
        want = p0 - 0.002 * g / (v.sqrt() + 1e-6)
        assert torch.allclose(p.detach().double(), want, rtol=1e-5, atol=1e-7)
        moved += int((p.detach().double() != p0).any())
    assert moved >= 10                                   # every live tensor moved (hidd_encoder does not)
    assert torch.equal(net.hidd_encoder.weight.detach().double(), before[[id(q) for q in tr.params].index(
        id(net.hidd_encoder.weight))])
    # the forward path sees the new weights (packed-weight cache invalidated by the version bump)
    x = torch.zeros(6, args.nagents, args.num_inputs, device="cuda")
    hc = net.init_hidden(6)
    info = {"comm_action": np.zeros((6, args.nagents), dtype=np.int64)}
    a1, v1, _ = net([x, hc], info)
    ref = torch.nn.functional.linear(torch.nn.functional.linear(x, net.encoder.weight, net.encoder.bias),
                                     net.f_module.weight_ih)   # touches the updated tensors: no stale views
    assert torch.isfinite(v1).all() and torch.isfinite(ref).all()
    net2 = CommNetMLP(args, args.num_inputs)
    net2.load_state_dict(net.state_dict())
    a2, v2, _ = net2([x, hc], info)
    assert torch.allclose(v1, v2, atol=1e-6)


def test_main_save_and_load_round_trip(tmp_path):
    """main.py:260-272 checkpoint = {'policy_net', 'log', 'trainer'}: written, re-loaded, training continues."""
    from ic3net_b200 import main as m
    common = ["--env_name", "predator_prey", "--nagents", "3", "--dim", "5", "--vision", "0", "--max_steps", "20",
              "--hid_size", "128", "--ic3net", "--recurrent", "--nenvs", "8", "--num_epochs", "1", "--epoch_size", "2",
              "--batch_size", "40", "--seed", "3"]
    p1, p2 = str(tmp_path / "a.pt"), str(tmp_path / "b.pt")
    assert m.main(common + ["--save", p1]) == 0
    with m._utils_alias():        # the log is pickled under the reference's class path (utils.LogField, main.py:260-265)
        d = torch.load(p1, weights_only=False)
    assert set(d) == {"policy_net", "log", "trainer"}
    keys = set(d["policy_net"])
    for k in ("encoder.weight", "f_module.weight_ih", "f_module.bias_hh", "C_modules.0.weight", "heads.0.weight",
              "heads.1.bias", "value_head.weight", "hidd_encoder.weight"):
        assert k in keys                                 # reference state_dict names (comm.py:31-96)
    tr = d["trainer"]
    assert tr["param_groups"][0]["alpha"] == 0.97 and tr["param_groups"][0]["eps"] == 1e-6
    assert len(tr["state"]) == len(tr["param_groups"][0]["params"]) and "square_avg" in tr["state"][0]
    assert len(d["log"]["epoch"].data) == 1
    assert m.main(common + ["--load", p1, "--save", p2]) == 0
    with m._utils_alias():
        d2 = torch.load(p2, weights_only=False)
    assert len(d2["log"]["epoch"].data) == 2             # the loaded log continues
    assert float(d2["trainer"]["state"][0]["step"]) == 4.0          # 2 + 2 optimizer steps
    w1, w2 = d["policy_net"]["encoder.weight"], d2["policy_net"]["encoder.weight"]
    assert not torch.equal(w1, w2)
