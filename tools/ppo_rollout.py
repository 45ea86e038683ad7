#!/usr/bin/env python
"""BASELINE.json configs[3]: bread_crumb_env PPO iteration — device-resident rollout (tcgen05 policy forward + fused env step)
followed by a clipped-surrogate update whose gradients are all-reduced over NCCL in one flat bucket.

  python tools/ppo_rollout.py [--envs 4096] [--steps 16] [--iters 2]                     # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/ppo_rollout.py

Prints one JSON line per iteration on rank 0 (rollout substeps/s over all ranks, update time, all-reduced rollout statistics)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from mujoco_rl_manipulate_unknown_objects_b200 import make_config
from mujoco_rl_manipulate_unknown_objects_b200.rollout import RolloutWorker, PPOLearner, dist_info

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="bread_crumb"); ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--steps", type=int, default=16); ap.add_argument("--iters", type=int, default=2); ap.add_argument("--minibatch", type=int, default=4096)
a = ap.parse_args()
rank, world, local = dist_info()
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = RolloutWorker(make_config(sim_env="/xmls/%s_env.xml" % a.scene), envs_per_gpu=a.envs, device=local, seed=0)
learner = PPOLearner(w, minibatch=a.minibatch, seed=1)
storage = w.make_storage(a.steps)
for it in range(a.iters):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    stats = w.collect(a.steps, storage=storage)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    loss = learner.update(storage, w.sim.obs)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    tm = torch.tensor([t1 - t0, t2 - t1], device=w.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    stats.all_reduce()
    d = stats.as_dict()
    ar = learner.opt.harvest_timings()
    learner.opt.allreduce_ms = []
    if rank == 0:
        print(json.dumps({"iter": it, "optimizer_steps": len(ar) if world > 1 else None,
                          "grad_allreduce_ms": {"mean": sum(ar) / len(ar), "min": min(ar), "max": max(ar), "bucket_bytes": 4 * learner.opt.count} if ar else None, "n_gpus": world, "envs_per_gpu": a.envs, "rollout_steps": a.steps, "rollout_s": tm[0].item(), "update_s": tm[1].item(),
                          "substeps_per_s": d.get("substeps", 0) / tm[0].item(), "transitions_per_s": a.envs * world * a.steps / tm[0].item(), "loss": loss, "stats": d}))
w.close()
if world > 1:
    dist.destroy_process_group()
