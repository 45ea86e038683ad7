#!/usr/bin/env python
"""Development aid: the step kernel in the STEADY STATE of a random-action rollout (agent steps >= 150 from reset), where
a quarter of the environments run a timed-out gripper-close phase (fingers pressed against each other, 400 substeps).
Prints per-step kernel time, chain-length classes and throughput for a few batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
from mujoco_rl_manipulate_unknown_objects_b200._native import INFO as I
scene = sys.argv[1] if len(sys.argv) > 1 else "acorn"
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4096]
PRE, K = int(os.environ.get("PRE", 150)), int(os.environ.get("K", 30))
for N in sizes:
    sim = GripperSim(make_config(sim_env="/xmls/%s_env.xml" % scene), num_envs=N)
    gen = torch.Generator(device="cuda").manual_seed(0)
    for i in range(PRE):
        sim.step(torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1)
    torch.cuda.synchronize(); sim.step_kernel_ms(reset=True)
    ms, sub, mx, frac, ncon, its, tops = [], [], [], [], [], [], []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_ms = 0.0
    for i in range(K):
        a = torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1
        ev0.record(); sim.step(a); ev1.record(); torch.cuda.synchronize()
        tot_ms += ev0.elapsed_time(ev1)
        info = sim.info.cpu().numpy()
        ns = info[:, I["NSUB_A"]:I["NSUB_A"] + 3].sum(1)
        srt = np.sort(ns)[::-1]; top3 = np.argsort(ns)[::-1][:3]
        tops.append((int(srt[0]), int(srt[1]), int(srt[7]), int(srt[63]), int(srt[295]), " ".join("%d+%d+%d" % tuple(int(x) for x in info[e, I["NSUB_A"]:I["NSUB_A"] + 3]) for e in top3)))
        ms.append(sim.step_kernel_ms(reset=True)); sub.append(ns.sum()); mx.append(ns.max()); frac.append((ns >= 400).mean())
        ncon.append(info[:, I["NCON_MAX"]].mean()); its.append(info[:, I["SOLVER_ITERS"]].sum() / max(ns.sum(), 1))
    print("%s N=%d steps %d-%d: %.2f M substeps/s (events), kernel-phase ms mean %.2f (min %.1f max %.1f), step ms %.2f, substeps/transition %.0f, "
          "max chain mean %.0f, frac>=400 %.3f, ncon_max mean %.2f, newton iters/substep %.2f, us per round of longest chain %.1f" % (
              scene, N, PRE, PRE + K, sum(sub) / tot_ms / 1e3, np.mean(ms), np.min(ms), np.max(ms), tot_ms / K, np.mean(sub) / N, np.mean(mx), np.mean(frac),
              np.mean(ncon), np.mean(its), 1e3 * np.mean(ms) / np.mean(mx)), flush=True)
    if os.environ.get("VERBOSE"):
        print("per step: kernel ms | longest chain, 2nd, 8th, 64th, 296th longest | substeps/transition")
        for a_, t_, s_ in zip(ms, tops, sub):
            print("  %6.2f | %4d %4d %4d %4d %4d | %.0f | top-3 (A+B+C): %s" % ((a_,) + t_[:5] + (s_ / N, t_[5])))
    sim.close()
