"""CPU baseline: the oracle's restatement of the reference rollout (one env per process,
float64, OMP_NUM_THREADS=1, reference README.md:46-48) timed on the host cores.
TEST / BENCH INFRASTRUCTURE -- only bench.py's ``cpu_baseline`` and ``--impl reference``
legs call this; it is never the thing shipped.

Each worker process plays one reference worker (multi_processing.py:6-38, run_batch
half): it owns one environment and loops ``Trainer.get_episode`` (oracle/rollout.py)
until its time budget is spent.  Throughput = sum of env steps * nagents / wall time,
i.e. ``stat['num_steps'] * nagents / seconds`` of the reference's run_batch.
"""
import argparse
import json
import multiprocessing as mp
import os
import time

import numpy as np


def _worker(cfg, rank, budget_s, nsamples, q):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import policy, pp_env, tj_env
    from oracle.gen_golden import make_weights
    from oracle.rollout import run_episode
    a = argparse.Namespace(**cfg["args"])
    if a.env_name == "predator_prey":
        env = pp_env.PredatorPreyOracle(a.nagents, a.dim, a.vision, a.mode, 1, False)
        obs_dim = env.obs_dim
    else:
        z = np.load(cfg["tables"])
        ln, cells = z["route_len"], z["route_cells"]
        routes = [[cells[g, k, :ln[g, k]] for k in range(ln.shape[1])] for g in range(ln.shape[0])]
        env = tj_env.TrafficJunctionOracle(a.nagents, a.dim, a.vision, a.difficulty,
                                           {"grid": z["grid"], "routes": routes}, a.add_rate_min, a.add_rate_max,
                                           a.curr_start, a.curr_end)
        obs_dim = env.obs_dim
    params = policy.params_to_f64(make_weights(0, obs_dim, a.hid_size, cfg["heads"]))
    k, tick = 0, 0
    run_episode(env, params, a, 1, rank, tick0=0, episode=0, max_steps=2)      # warm-up
    out = []
    for _ in range(nsamples):
        steps = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < budget_s:
            ep = run_episode(env, params, a, 1, rank, tick0=tick, episode=k,
                             max_steps=min(a.max_steps, max(2, int(a.max_steps * min(1.0, budget_s)))))
            steps += ep["num_steps"]
            tick += ep["num_steps"]
            k += 1
        out.append((steps, time.perf_counter() - t0))
    q.put(out)


def run(cfg, nprocs, budget_s, nsamples=1):
    """nsamples back-to-back samples of budget_s seconds in every worker.
    Returns dict(value=agent-env-steps/s, env_steps, seconds, nprocs, samples=[(env_steps, seconds)...])."""
    os.environ["OMP_NUM_THREADS"] = "1"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(cfg, r, budget_s, nsamples, q)) for r in range(nprocs)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    samples = [(sum(r[i][0] for r in res), max(r[i][1] for r in res)) for i in range(nsamples)]
    wall = sum(x[1] for x in samples)
    steps = sum(x[0] for x in samples)
    return dict(value=steps * cfg["args"]["nagents"] / wall, env_steps=steps, seconds=wall, nprocs=nprocs,
                samples=samples, launch_wall=time.perf_counter() - t0)


if __name__ == "__main__":
    import sys
    cfg = json.loads(sys.argv[1])
    print(json.dumps(run(cfg, int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 1)))
