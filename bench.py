#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched gripper environment (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scene acorn] [--envs 4096]

Workload (BASELINE.json configs[1]): acorn_env (stand-in mesh), 4 096 environments per GPU, actions
U(-1,1)^6 from a seeded device generator, physics + reward + observation, auto-reset on.  One "step" = one VecEnv.step
over the whole batch (one pass of the hot path).  Environments shard across ranks with no data-path collective
("weak" scaling); NCCL carries only the rollout statistics.

Unit: the headline `value` is S = physics substeps/s (one substep == one MuJoCo `mj_step`, 2 ms of simulated time in one
environment; BASELINE.md §3 reads the 1e7/s target in this unit).  T = agent transitions/s (VecEnv.step rows) and the
mean substeps per transition that links the two are printed beside it.

The reference arm (--impl reference) times the CPU implementation of the same path on the host cores: the oracle's
fp64 C restatement (oracle/engine.c, kind "port" — MuJoCo itself cannot be installed here) with one
environment per thread, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (acorn_env, 4096 envs/GPU)"
UNIT = "physics substeps/s (one substep = one mj_step of one env)"
BYTES_PER_SUBSTEP = 428 + 64  # 56 f32 read + 51 f32 written + ~64 B contact summary, if state round-tripped HBM
BYTES_PER_TRANSITION = 21100  # state in/out + action + obs write (20 480 B) + goals/reward/done


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)  # >= 200 timed agent steps after >= 50 warm-up
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="acorn")
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--direction", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--policy", action="store_true", help="actions from the tensor-core policy forward on the previous observation "
                    "(device-resident rollout loop, BASELINE.json configs[3]) instead of U(-1,1)^6")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target wall time of the CPU baseline sample")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_rollout(scene, direction, seed, seconds, threads=None):
    """The oracle's C restatement, one environment per host thread, same action distribution, auto-reset.
    Returns dict(value substeps/s, transitions/s, cores, sample)."""
    from oracle import engine
    from mujoco_rl_manipulate_unknown_objects_b200 import compile_model
    cm = compile_model("/xmls/%s_env.xml" % scene)
    om = engine.Model(engine.model_dict_from_export(cm))
    cores = threads or os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    # calibration: one env per core, a few transitions
    nenv, nst = cores, 4
    a = rng.uniform(-1, 1, (nenv, nst, 6))
    t = time.perf_counter()
    sub, tr, _ = engine.rollout_threads(om, a, cores, direction=direction)
    dt = max(time.perf_counter() - t, 1e-4)
    per_tr = dt / (tr / cores)
    nst = int(max(8, min(400, seconds / max(per_tr * 4, 1e-6))))  # 4 environments per core
    nenv = cores * 4
    a = rng.uniform(-1, 1, (nenv, nst, 6))
    t = time.perf_counter()
    sub, tr, rsum = engine.rollout_threads(om, a, cores, direction=direction)
    dt = time.perf_counter() - t
    return dict(value=sub / dt, unit=UNIT, cores=cores, kind="port",
                sample="%d envs x %d transitions of %s_env on %d host threads (%d substeps, %.1f s), oracle/engine.c fp64 restatement, not MuJoCo" % (
                    nenv, nst, scene, cores, sub, dt),
                transitions_per_s=tr / dt, substeps_per_transition=sub / max(tr, 1), seconds=dt)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    res = []
    for _ in range(max(1, min(args.steps, 3))):  # each "step" of this arm is one bounded sample
        res.append(cpu_rollout(args.scene, args.direction, args.seed, max(2.0, args.cpu_seconds)))
    best = max(res, key=lambda r: r["value"])
    line = {"impl": "reference", "metric": METRIC, "value": best["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": best["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s_env, U(-1,1)^6 actions, physics + reward + scalar obs channels (no image render on the CPU arm)" % args.scene,
                       "direction": args.direction},
            "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "transitions_per_s": best["transitions_per_s"], "substeps_per_transition": best["substeps_per_transition"],
            "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
    from mujoco_rl_manipulate_unknown_objects_b200._native import INFO

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, K, W = args.envs, args.steps, args.warmup
    cfg = make_config(sim_env="/xmls/%s_env.xml" % args.scene, direction=args.direction)
    sim = GripperSim(cfg, num_envs=N, device=local, auto_reset=True)
    gen = torch.Generator(device=dev).manual_seed(args.seed + rank)  # per-rank seed
    actions = torch.rand((W + K, N, sim.action_dim), device=dev, generator=gen) * 2 - 1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    sub_acc = torch.zeros(1, device=dev, dtype=torch.float64)
    done_acc = torch.zeros(1, device=dev, dtype=torch.float64)
    ret_acc = torch.zeros(1, device=dev, dtype=torch.float64)
    chain_acc = torch.zeros(1, device=dev, dtype=torch.float64)  # sum over steps of the longest substep chain of the batch
    c0, c1 = INFO["NSUB_A"], INFO["NSUB_A"] + 3
    from mujoco_rl_manipulate_unknown_objects_b200.policy import GripperPolicy
    policy = GripperPolicy(max_envs=N, obs_shape=sim.obs_shape, action_dim=sim.action_dim, device=local, seed=args.seed)
    noise = torch.randn((W + K, N, sim.action_dim), device=dev, generator=gen)
    pol_act = torch.empty((N, sim.action_dim), device=dev)

    def one_step(i):
        if args.policy:
            policy.forward(sim.obs, deterministic=False, noise=noise[i], out=pol_act)
            sim.step(pol_act)
        else:
            sim.step(actions[i])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sim.reset()
    for i in range(W):
        one_step(i)
    barrier()
    sim.step_kernel_ms(reset=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = sim.launch_count + policy.launch_count
    barrier()
    wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        one_step(W + i)
        ev[i][1].record()
        nsub = sim.info[:, c0:c1].sum(1)
        sub_acc += nsub.sum()
        chain_acc += nsub.max()
        done_acc += sim.done.sum()
        ret_acc += sim.reward.sum()
    barrier()
    wall = time.perf_counter() - wall0
    launches = sim.launch_count + policy.launch_count - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    kernel_ms = sim.step_kernel_ms(reset=True)
    clk = clocks.stop() if rank == 0 else None
    # ---- policy forward alone on the last observation (tensor roofline): CUDA events on the launching stream, L2 flushed
    pol_ms = []
    for i in range(3 + 10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        policy.forward(sim.obs, out=pol_act)
        e1.record()
        torch.cuda.synchronize(dev)
        if i >= 3:
            pol_ms.append(e0.elapsed_time(e1))
    pol_ms = float(np.median(pol_ms))
    # ---- end to end through the C-ABI with HOST buffers (the call a VecEnv makes): H2D actions, step, D2H results
    Cc, H, Wd = sim.obs_shape
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
    h_act = pin((N, sim.action_dim), torch.float32)
    h_obs, h_rew, h_done = pin((N, Cc, H, Wd), torch.uint8), pin((N,), torch.float32), pin((N,), torch.uint8)
    h_ag, h_dg, h_info = pin((N, 2), torch.float32), pin((N, 2), torch.float32), pin((N, INFO["STRIDE"]), torch.float32)
    acts_host = actions[W:].cpu().numpy()
    KE = K
    e2e_sub = 0.0
    if not args.policy:
        # replay: same reset state, same warm-up and timed action tapes -> the very trajectories (and substep chains) of the
        # device-timed region, now through host buffers, so `e2e` and `value` differ only by the copies
        sim.reset()
        for i in range(W):
            one_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(KE):
        h_act[...] = acts_host[i % K]
        sim.step_host(h_act, obs=h_obs, achieved=h_ag, desired=h_dg, reward=h_rew, done=h_done, info=h_info)
        e2e_sub += float(h_info[:, c0:c1].sum())
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = h_act.nbytes
    d2h = h_obs.nbytes + h_rew.nbytes + h_done.nbytes + h_ag.nbytes + h_dg.nbytes + h_info.nbytes
    # ---- aggregate over ranks: units summed, time = max
    stats = torch.tensor([sub_acc.item(), float(N * K), done_acc.item(), ret_acc.item(), e2e_sub, float(launches), chain_acc.item()], device=dev, dtype=torch.float64)
    tmax = torch.tensor([dev_ms, wall * 1e3, e2e_s * 1e3, kernel_ms, pol_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)  # rollout-statistics reduction: the only collective on this path
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    stats, tmax = stats.cpu().numpy(), tmax.cpu().numpy()
    if rank == 0:
        substeps, transitions = stats[0], stats[1]
        dev_s = tmax[0] / 1e3
        value = substeps / dev_s
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": tmax[0] / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "%s_env (%s), %d envs/GPU, %s, physics + progress reward + RGB-D observation, auto-reset" % (
                    args.scene, "synthetic stand-in mesh: acorn.stl is absent from the reference tree" if args.scene == "acorn" else "reference mesh", N,
                    "actions from the tcgen05 policy forward on the previous observation (random-init weights, stochastic)" if args.policy
                    else "U(-1,1)^6 device-generated actions"),
                    "direction": args.direction, "envs_per_gpu": N, "l2": "256 MiB buffer written between timed steps (outside the event pairs)",
                    "parallelism": "env-sharded x%d, no data-path collective" % world},
                "transitions_per_s": transitions / dev_s, "substeps_per_transition": substeps / transitions, "episodes_finished": stats[2],
                "mean_reward_per_transition": stats[3] / transitions, "wall_ms_per_step": tmax[1] / K,
                "e2e": {"value": stats[4] / (tmax[2] / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": KE,
                        "transitions_per_s": N * KE * world / (tmax[2] / 1e3), "api": "grs_step_host (C-ABI, pinned host buffers)"},
                "gpu_launches": int(stats[5] // world), "clocks": clk,
                # an agent step is a chain of dependent substeps; the synchronous VecEnv step lasts as long as the batch's longest chain
                "longest_chain_substeps": stats[6] / world / K, "us_per_substep_of_longest_chain": tmax[3] * 1e3 / max(stats[6] / world / K, 1.0)}
        per_launch_sub = substeps / world / K
        kms = tmax[3]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = per_launch_sub * BYTES_PER_SUBSTEP / (kms / 1e3) / 1e9 if kms > 0 else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_env_step_bytes_per_launch")
        except Exception:
            pass
        line["roofline"] = {"bound": "hbm", "kernel": "k_env_step_ls, physics phase (fused agent step: controller + <=1200 substeps + reward, state on chip; "
                                                      "device-timed: first block start -> last block leaving the substep loop)",
                            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                            "kernel_ms": kms, "bytes_per_substep": BYTES_PER_SUBSTEP, "substeps_per_launch": per_launch_sub,
                            "note": "fused kernel keeps state on chip: it is latency bound by the batch's longest chain of dependent substeps, not HBM bound "
                                    "(SURVEY.md §8d, DESIGN.md §3.1; traffic = physics + observation phases of one launch); "
                                    "per-transition accounting (%d B incl. the 20 480 B observation) gives %.1f GB/s over the whole step" % (
                                        BYTES_PER_TRANSITION, transitions / world / K * BYTES_PER_TRANSITION / (tmax[0] / K / 1e3) / 1e9)}
        flop = 2 * (225 * 64 * (sim.obs_shape[0] - 1) * 32 + 36 * 512 * 64 + 16 * 576 * 64 + 1024 * 512 + 514 * 256 + 256 * 256 + 256 * 2 * sim.action_dim)
        tpeak = float(peaks.get("bf16_tflops_sustained", 1408.0))
        tfl = N * flop / (tmax[4] / 1e3) / 1e12 if tmax[4] > 0 else None
        line["policy"] = {"kernel": "k_layer x7 (tcgen05.mma kind::f16, TMEM accumulators): NatureCNN + actor MLP forward for %d observations" % N,
                          "ms": tmax[4], "in_timed_region": bool(args.policy), "flop_per_obs": flop,
                          "roofline": {"bound": "tensor", "achieved": tfl, "peak": tpeak, "unit": "TFLOP/s", "frac": tfl / tpeak if tfl else None,
                                       "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback"}}
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_rollout(args.scene, args.direction, args.seed, args.cpu_seconds)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
                line["cpu_baseline"]["transitions_per_s"] = cb["transitions_per_s"]
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    policy.close()
    sim.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
