"""CPU baseline OF RECORD: the reference's own ``MultiProcessTrainer`` (multi_processing.py:41-104) at
``nprocesses`` forked workers, ``OMP_NUM_THREADS=1`` (README.md:46-48), float64 as shipped (main.py:20), timed on
the host cores.  TEST / BENCH INFRASTRUCTURE -- only bench.py's ``cpu_baseline`` / ``--impl reference`` legs run
this module (in a fresh interpreter); it is never the thing shipped.

The reference code is imported UNMODIFIED from /root/reference (development container) or from its staged copy
``oracle/_ref`` (GPU box; oracle/build_ref.py) through the five import shims of oracle/ref_shims.py (SURVEY.md
section 8(c)); ``main.py`` itself is not imported (it needs visdom and runs at import): this module does what
main.py:134-186 does -- build args, env, ``CommNetMLP``, share the parameters, ``MultiProcessTrainer(args, lambda:
Trainer(args, policy_net, env))`` -- and then times ``trainer.train_batch(epoch)`` calls.

Two modes:
  train_batch   what the reference does per update: run_batch + compute_grad in every worker, gradient sum,
                RMSprop step;
  rollout       the same call with ``Trainer.compute_grad`` patched to a no-op in every process, i.e. run_batch
                only -- the like-for-like of the GPU arm's rollout metric.
agent-env-steps = stat['num_steps'] (already summed over the processes, multi_processing.py:86-88) * nagents.

    python -m oracle.ref_baseline '<json cfg>'     ->  one JSON line
cfg: {"args": {...reference flags...}, "nprocesses": 16, "modes": ["rollout", "train_batch"], "warmup": 1,
      "iters": 3}
"""
import json
import os
import sys
import time

os.environ["OMP_NUM_THREADS"] = "1"
os.environ.setdefault("MKL_NUM_THREADS", "1")


def _build(cfg):
    import torch

    from oracle import ref_shims

    ref_shims.install()
    torch.set_num_threads(1)
    torch.set_default_dtype(torch.float64)          # main.py:20 (torch.set_default_tensor_type('torch.DoubleTensor'))
    import comm
    import multi_processing
    import trainer as ref_trainer

    a = ref_shims.make_args(**cfg["args"])
    a.nprocesses = int(cfg["nprocesses"])
    env0 = ref_shims.make_ref_env(a)
    ref_shims.finish_args(a, env0)
    torch.manual_seed(int(a.seed))
    policy_net = comm.CommNetMLP(a, a.num_inputs)
    for p in policy_net.parameters():               # main.py:177-179
        p.data.share_memory_()
    return a, policy_net, ref_trainer, multi_processing, ref_shims


def run(cfg):
    a, policy_net, ref_trainer, mpmod, ref_shims = _build(cfg)
    out = dict(nprocesses=a.nprocesses, nagents=a.nagents, batch_size=a.batch_size, modes={})
    orig_grad = ref_trainer.Trainer.compute_grad
    for mode in cfg.get("modes", ["rollout", "train_batch"]):
        # the patch must be in place before MultiProcessTrainer forks its workers (multi_processing.py:47-51)
        ref_trainer.Trainer.compute_grad = orig_grad if mode == "train_batch" else (lambda self, batch: dict())
        t_make = time.perf_counter()
        mpt = mpmod.MultiProcessTrainer(a, lambda: ref_trainer.Trainer(a, policy_net, ref_shims.make_ref_env(a)))
        t_make = time.perf_counter() - t_make
        samples = []
        for it in range(int(cfg.get("warmup", 1)) + int(cfg.get("iters", 3))):
            t0 = time.perf_counter()
            stat = mpt.train_batch(it)
            samples.append((int(stat["num_steps"]), time.perf_counter() - t0))
        mpt.quit()
        out["modes"][mode] = dict(samples=samples, setup_s=t_make)
    ref_trainer.Trainer.compute_grad = orig_grad
    return out


if __name__ == "__main__":
    res = run(json.loads(sys.argv[1]))
    sys.stdout.write(json.dumps(res) + "\n")
    sys.stdout.flush()
    os._exit(0)                                      # main.py:292-295: forked workers may still be draining
