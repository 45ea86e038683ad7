#!/usr/bin/env python
"""Development aid: the step kernel with one environment per SM (N = 148) — every warp runs alone, so an ncu source-page
capture of this launch shows where a lone warp's substep chain (the critical path of the kernel) spends its cycles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
sim = GripperSim(make_config(sim_env="/xmls/acorn_env.xml"), num_envs=N)
gen = torch.Generator(device="cuda").manual_seed(0)
for i in range(16):
    sim.step(torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1)
    torch.cuda.synchronize()
    ns = sim.info[:, 10:13].sum(1)
    print("step %d: kernel %.2f ms, max chain %d, mean %.0f -> %.1f us per substep of the longest chain" % (
        i, sim.step_kernel_ms(reset=True), int(ns.max()), float(ns.mean()), 1e3 * sim.step_kernel_ms() / max(1, int(ns.max()))))
