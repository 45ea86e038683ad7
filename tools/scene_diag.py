#!/usr/bin/env python
"""Development aid: why a scene's agent step is slow — substep chain lengths, Newton iterations per substep, contact counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
from mujoco_rl_manipulate_unknown_objects_b200._native import INFO as I
N = 4096
for scene in sys.argv[1:] or ["acorn", "sugar_cube", "sand_ball", "bread_crumb"]:
    sim = GripperSim(make_config(sim_env="/xmls/%s_env.xml" % scene), num_envs=N)
    gen = torch.Generator(device="cuda").manual_seed(0)
    rows = []
    for i in range(30):
        sim.step(torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1)
        if i >= 10:
            info = sim.info.cpu().numpy()
            ns = info[:, I["NSUB_A"]:I["NSUB_A"] + 3].sum(1)
            it = info[:, I["SOLVER_ITERS"]] / np.maximum(ns, 1)
            rows.append((sim.step_kernel_ms(), ns.max(), ns.mean(), it.mean(), np.percentile(it, 99), it.max(), info[:, I["NCON_MAX"]].mean(), info[:, I["NCON_MAX"]].max(),
                         (info[:, I["FLAGS"]] != 0).sum(), info[:, I["SOLVER_ITERS"]].max()))
    r = np.array(rows)
    print("%-12s kernel %.1f ms | chain max %.0f mean %.0f | Newton iters/substep mean %.2f p99 %.2f max %.2f (max total per agent step %.0f) | ncon_max mean %.1f max %.0f | flagged %d" % (
        scene, r[:, 0].mean(), r[:, 1].mean(), r[:, 2].mean(), r[:, 3].mean(), r[:, 4].mean(), r[:, 5].max(), r[:, 9].max(), r[:, 6].mean(), r[:, 7].max(), r[:, 8].sum()))
    sim.close()
