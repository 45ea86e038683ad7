#!/usr/bin/env python
"""Development aid: substep latency of a lone warp vs. throughput of a full GPU, and the per-step substep-count tail."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
cfg = make_config(sim_env="/xmls/acorn_env.xml")
for n in (1, 148, 1480, 2960, 4096):
    sim = GripperSim(cfg, num_envs=n, auto_reset=False)
    sim.set_state(ctrl=np.array([0.3, 0.1, -0.2, 0.1, 0.1, 0, 0], np.float32))
    sim.substep(50); torch.cuda.synchronize()
    t = time.perf_counter(); sim.substep(1000); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("N=%5d: 1000 substeps in %.1f ms -> %.1f us per substep round, %.2f M substeps/s" % (n, dt * 1e3, dt * 1e3, n * 1000 / dt / 1e6))
    sim.close()
N = 4096
sim = GripperSim(cfg, num_envs=N)
gen = torch.Generator(device="cuda").manual_seed(0)
for i in range(25):
    sim.step(torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1)
    if i >= 20:
        ns = sim.info[:, 10:13].sum(1).cpu().numpy()
        print("step %d: nsub mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f  #>=800: %d  kernel %.1f ms" % (i, ns.mean(), np.percentile(ns, 50), np.percentile(ns, 90), np.percentile(ns, 99), ns.max(), (ns >= 800).sum(), sim.step_kernel_ms()))
