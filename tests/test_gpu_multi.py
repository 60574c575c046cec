"""Two ranks under NCCL: MultiGPUTrainer.train_batch on real gradients.

The reference semantics to preserve (multi_processing.py:74-98): gradient = sum over all workers' backward passes
divided by the GLOBAL num_steps, statistics merged over workers, ONE optimizer step, all workers on one set of
parameters.  Checked against a single process that owns both ranks' env slots (same Philox streams).
Skipped with the reason when the box has fewer than 2 GPUs (`gpurun --gpus 2` runs it)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_rank_nccl_train_batch_equals_single_process_sum():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (have %d)" % (torch.cuda.device_count() if torch.cuda.is_available() else 0))
    sys.path.insert(0, HERE)
    from helpers import load_golden
    import nccl_worker
    B = 6
    with tempfile.TemporaryDirectory() as d:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29731", os.path.join(HERE, "nccl_worker.py"), d, str(B)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
        assert r.returncode == 0, r.stdout.decode()[-4000:]
        ranks = [dict(np.load(os.path.join(d, "rank%d.npz" % k))) for k in range(2)]
    r0, r1 = ranks
    # replicas: different initial seeds were broadcast away, and they stay bit-identical through two updates
    assert float(r0["init_diff"]) == 0.0 and float(r0["final_diff"]) == 0.0
    for k in ("p0", "p1", "p2", "g0", "g1"):
        assert np.array_equal(r0[k], r1[k]), k
    assert int(r0["collectives"]) == 2                      # exactly one gradient all-reduce per update
    # single process owning both shards: env ids [0, 2B), same initial parameters as rank 0
    meta, _ = load_golden("grad_pp_easy_ic3net")
    args, net, tr = nccl_worker.build(meta, 2 * B, 0, 17, torch_seed=100)
    assert np.array_equal(tr.optimizer.flat_params.detach().cpu().numpy(), r0["p0"])
    for u in range(2):
        stat = tr.train_batch(u)
        assert stat["num_steps"] == int(r0["steps%d" % u]) and stat["num_episodes"] == int(r0["episodes%d" % u])
        assert np.allclose(stat["reward"], r0["reward%d" % u], rtol=1e-6, atol=1e-6)
        want_l = np.array([stat[k] for k in ("action_loss", "value_loss", "entropy")])
        assert np.allclose(want_l, r0["losses%d" % u], rtol=1e-4, atol=1e-3)
        g = tr.optimizer.flat_grads.detach().cpu().numpy()          # summed over the 2B slots, / num_steps
        scale = np.abs(g).max()
        assert np.abs(g - r0["g%d" % u]).max() <= 2e-5 * scale, u
        p = tr.optimizer.flat_params.detach().cpu().numpy()
        assert np.abs(p - r0["p%d" % (u + 1)]).max() <= 1e-5, u
