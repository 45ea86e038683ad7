#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): reset, a few agent steps (fused step kernel:
physics + observation phases), raw substeps, contacts query, raw render, policy forward.
  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
from mujoco_rl_manipulate_unknown_objects_b200.policy import GripperPolicy
N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for scene in ("sugar_cube_env", "gripper_two_fingers"):
    sim = GripperSim(make_config(sim_env="/xmls/%s.xml" % scene, time_horizon=3), num_envs=N)
    gen = torch.Generator(device="cuda").manual_seed(0)
    for i in range(steps):
        a = torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1
        a[:, 0] = a[:, 0].abs()
        sim.step(a)
    sim.substep(3)
    sim.contacts()
    sim.render(camera_id=1, width=32, height=32)
    torch.cuda.synchronize()
    print(scene, "ok: reward sum %.4f" % float(sim.reward.sum()))
    if scene == "sugar_cube_env":
        pol = GripperPolicy(max_envs=N)
        act = pol(sim.obs)
        torch.cuda.synchronize()
        print("policy ok", tuple(act.shape))
        pol.close()
    sim.close()
