#!/usr/bin/env python
"""Development aid: the critical-path class of the steady state in isolation.  A fifth of the environments of a random-action
rollout run a gripper-close phase that times out after 400 substeps (fingers pressed against each other); they are queued first
and their 475-substep chains decide the step time.  This tool rolls a batch to agent step 150, picks environments whose last
step timed out, copies their states over the whole batch and steps them with a pure `close` action: every warp then runs that
class, so the kernel time / 475 is its cost per substep in full blocks, and an ncu capture of the launch shows where it goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
from mujoco_rl_manipulate_unknown_objects_b200._native import INFO as I
scene = sys.argv[1] if len(sys.argv) > 1 else "acorn"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sim = GripperSim(make_config(sim_env="/xmls/%s_env.xml" % scene), num_envs=N, auto_reset=False)
gen = torch.Generator(device="cuda").manual_seed(0)
for i in range(150):
    sim.step(torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1)
info = sim.info.cpu().numpy()
st = sim.get_state()
stalled = np.flatnonzero((info[:, I["NSUB_C"]] >= 400) & (st["flags"][:, 0] == 1) & (st["flags"][:, 2] == 0))
print("stalled-close environments at step 150: %d of %d" % (len(stalled), N))
src = stalled[np.arange(N) % len(stalled)]
sim.set_state(qpos=st["qpos"][src], qvel=st["qvel"][src], ctrl=st["ctrl"][src], warmstart=st["warmstart"][src], flags=st["flags"][src], xfrc_z=st["xfrc_z"][src])
a = torch.zeros((N, 6), device="cuda"); a[:, 5] = -1.0
for k in range(3):
    sim.step_kernel_ms(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.step(a); e1.record(); torch.cuda.synchronize()
    info = sim.info.cpu().numpy()
    ns = info[:, I["NSUB_A"]:I["NSUB_A"] + 3].sum(1)
    print("step %d: %.2f ms (physics phase %.2f ms), substeps mean %.0f max %.0f, frac timed out %.3f, ncon_max mean %.2f, newton iters/substep %.2f -> %.1f us per substep of the chain, %.2f M substeps/s" % (
        k, e0.elapsed_time(e1), sim.step_kernel_ms(), ns.mean(), ns.max(), (info[:, I["NSUB_C"]] >= 400).mean(), info[:, I["NCON_MAX"]].mean(),
        info[:, I["SOLVER_ITERS"]].sum() / ns.sum(), 1e3 * sim.step_kernel_ms() / ns.max(), ns.sum() / e0.elapsed_time(e1) / 1e3))
sim.close()
