#!/usr/bin/env python
"""Development aid (GRS_STEP_TIMING=1): for each of a few agent steps in the steady state, the block whose physics phase ends
last: rounds, time per round, per-stage block-max cycles per round, rounds by active warps.  Answers: what makes a slow step slow?"""
import os, sys
os.environ["GRS_STEP_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
from mujoco_rl_manipulate_unknown_objects_b200._native import INFO as I
N, PRE, K = 4096, int(os.environ.get("PRE", 150)), int(os.environ.get("K", 24))
sim = GripperSim(make_config(sim_env="/xmls/acorn_env.xml"), num_envs=N)
gen = torch.Generator(device="cuda").manual_seed(0)
names = ["tick", "smooth", "constr", "solve", "eul+kin", "crb", "collis"]
G = 296
for i in range(PRE + K):
    a = torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1
    sim.step(a)
    if i < PRE:
        continue
    torch.cuda.synchronize()
    dbg = sim.debug.reshape(-1)
    d = dbg[: 32 * G].reshape(G, 32).cpu().numpy()
    bd = dbg[32 * G: 64 * G].reshape(G, 32).cpu().numpy()
    info = sim.info.cpu().numpy()
    ns = info[:, I["NSUB_A"]:I["NSUB_A"] + 3].sum(1)
    dur = bd[:, 0] / 1965.0e3
    order = dur.argsort()[::-1]
    b = order[0]
    r = d[b, 16]
    hist = bd[b, 1:13]
    print("step %d: kernel physics %.1f ms (p50 block %.1f, 2nd %.1f, 8th %.1f); slowest block %d: %d rounds, %.0f us/round; stage max kcyc/round: %s; mean-warp: %s; rounds by active: %s; its/substep batch %.2f ncon mean %.2f max %d" % (
        i, dur[b], np.median(dur), dur[order[1]], dur[order[7]], b, r, 1e3 * dur[b] / max(r, 1),
        " ".join("%s %.1f" % (n, x / r / 1e3) for n, x in zip(names, d[b, 8:15])),
        " ".join("%.1f" % (x / r / 1e3) for x in d[b, 0:7]),
        " ".join("%d:%d" % (k, hist[k]) for k in range(1, 9) if hist[k] > 0),
        info[:, I["SOLVER_ITERS"]].sum() / ns.sum(), info[:, I["NCON_MAX"]].mean(), info[:, I["NCON_MAX"]].max()), flush=True)
