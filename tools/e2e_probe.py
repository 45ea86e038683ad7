#!/usr/bin/env python
"""Development aid: what the host-buffer path (grs_step_host: H2D actions, step, D2H results) adds to the device-only step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
from mujoco_rl_manipulate_unknown_objects_b200._native import INFO
N = 4096
sim = GripperSim(make_config(sim_env="/xmls/acorn_env.xml"), num_envs=N)
C, H, W = sim.obs_shape
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
h_act = pin((N, 6), torch.float32); h_obs = pin((N, C, H, W), torch.uint8); h_rew = pin((N,), torch.float32); h_done = pin((N,), torch.uint8)
h_ag = pin((N, 2), torch.float32); h_dg = pin((N, 2), torch.float32); h_info = pin((N, INFO["STRIDE"]), torch.float32)
rng = np.random.default_rng(0)
rows = []
for i in range(40):
    h_act[...] = rng.uniform(-1, 1, (N, 6))
    t0 = time.perf_counter()
    sim.step_host(h_act, obs=h_obs, achieved=h_ag, desired=h_dg, reward=h_rew, done=h_done, info=h_info)
    t1 = time.perf_counter()
    if i >= 10:
        rows.append(((t1 - t0) * 1e3, sim.step_kernel_ms(), h_info[:, 10:13].sum(1).max()))
r = np.array(rows)
print("step_host wall %.2f ms | physics phase %.2f ms | wall - physics %.2f ms (min %.2f max %.2f) | max chain mean %.0f" % (
    r[:, 0].mean(), r[:, 1].mean(), (r[:, 0] - r[:, 1]).mean(), (r[:, 0] - r[:, 1]).min(), (r[:, 0] - r[:, 1]).max(), r[:, 2].mean()))
t0 = time.perf_counter()
for _ in range(5):
    torch.as_tensor(h_obs).copy_(sim.obs, non_blocking=True); torch.cuda.synchronize()
print("obs D2H alone: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
