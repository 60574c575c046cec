"""Worker of tests/test_gpu_multi.py: one rank of a 2-GPU NCCL job (launched with torch.distributed.run).
Runs MultiGPUTrainer.train_batch twice on its shard of env slots and dumps what the test compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(meta, B, env_id0, seed, torch_seed):
    from helpers import finish_args, ns
    from ic3net_b200 import data
    from ic3net_b200.comm import CommNetMLP
    from ic3net_b200.trainer import Trainer
    args = ns(meta["args"], nenvs=B, seed=seed, env_id0=env_id0, obs_mode="index", use_graph=False,
              record_for_grad=True, grad_window=16, lrate=0.002)
    env = data.init(args.env_name, args)
    finish_args(args, env)
    torch.manual_seed(torch_seed)
    net = CommNetMLP(args, args.num_inputs)
    return args, net, Trainer(args, net, env)


def main():
    from helpers import load_golden
    from ic3net_b200.multi_gpu import MultiGPUTrainer
    out_dir, B = sys.argv[1], int(sys.argv[2])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    meta, _ = load_golden("grad_pp_easy_ic3net")
    # every rank deliberately starts from DIFFERENT parameters: MultiGPUTrainer must broadcast rank 0's
    args, net, tr = build(meta, B, rank * B, 17, torch_seed=100 + rank)
    mgt = MultiGPUTrainer(args, lambda: tr)
    res = dict(init_diff=np.array(mgt.replica_checksum()))
    res["p0"] = tr.optimizer.flat_params.detach().cpu().numpy().copy()
    stats = []
    for u in range(2):
        stat = mgt.train_batch(u)
        stats.append(stat)
        res["g%d" % u] = tr.optimizer.flat_grads.detach().cpu().numpy().copy()     # all-reduced, already / num_steps
        res["p%d" % (u + 1)] = tr.optimizer.flat_params.detach().cpu().numpy().copy()
        res["steps%d" % u] = np.array(stat["num_steps"])
        res["episodes%d" % u] = np.array(stat["num_episodes"])
        res["reward%d" % u] = np.asarray(stat["reward"])
        res["losses%d" % u] = np.array([stat[k] for k in ("action_loss", "value_loss", "entropy")])
    res["final_diff"] = np.array(mgt.replica_checksum())
    res["collectives"] = np.array(mgt.collectives)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
