#!/usr/bin/env python
"""Development aid: how an environment's substep chain length relates to what is known BEFORE the agent step (action,
gripper_open flag) — the input of the longest-first queue order — plus the pinned D2H bandwidth of this box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
from mujoco_rl_manipulate_unknown_objects_b200._native import INFO as I
N = 4096
sim = GripperSim(make_config(sim_env="/xmls/acorn_env.xml"), num_envs=N)
gen = torch.Generator(device="cuda").manual_seed(0)
rec = []
for i in range(40):
    a = torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1
    gopen = sim.state[:, 47].clone() if hasattr(sim, "state") else None
    sim.step(a)
    if i >= 15:
        info = sim.info.cpu().numpy()
        rec.append(dict(a=a.cpu().numpy(), gopen=None if gopen is None else gopen.cpu().numpy(), ns=info[:, I["NSUB_A"]:I["NSUB_A"] + 3].copy(),
                        reached=info[:, I["REACHED_TARGET"]].copy(), ms=sim.step_kernel_ms()))
a = np.concatenate([r["a"] for r in rec]); ns = np.concatenate([r["ns"] for r in rec]); tot = ns.sum(1)
print("kernel ms per step:", np.round([r["ms"] for r in rec], 1))
print("per-step max chain:", [int(r["ns"].sum(1).max()) for r in rec])
print("us per round of the longest chain: mean %.1f  (steps with max<300: %.1f, others: %.1f); mean kernel %.2f ms" % (
    np.mean([1e3 * r["ms"] / r["ns"].sum(1).max() for r in rec]), np.mean([1e3 * r["ms"] / r["ns"].sum(1).max() for r in rec if r["ns"].sum(1).max() < 300] or [0]),
    np.mean([1e3 * r["ms"] / r["ns"].sum(1).max() for r in rec if r["ns"].sum(1).max() >= 300] or [0]), np.mean([r["ms"] for r in rec])))
tn = np.linalg.norm(a[:, :3], axis=1)
for lo, hi in ((0, .5), (.5, .8), (.8, 1.0), (1.0, 1.2), (1.2, 2)):
    m = (tn >= lo) & (tn < hi)
    print("|a_xyz| in [%.1f,%.1f): n=%6d  nsA mean %.0f p10 %.0f p90 %.0f max %.0f" % (lo, hi, m.sum(), ns[m, 0].mean(), np.percentile(ns[m, 0], 10), np.percentile(ns[m, 0], 90), ns[m, 0].max()))
if rec[0]["gopen"] is not None:
    g = np.concatenate([r["gopen"] for r in rec])
    will = ((a[:, 5] > 0) & (g == 0)) | ((a[:, 5] < 0) & (g != 0))
    for nm, m in (("gripper phase expected", will), ("not expected", ~will)):
        print("%-24s n=%6d  nsC mean %.0f p10 %.0f p90 %.0f max %.0f  total mean %.0f p90 %.0f" % (nm, m.sum(), ns[m, 2].mean(), np.percentile(ns[m, 2], 10), np.percentile(ns[m, 2], 90), ns[m, 2].max(), tot[m].mean(), np.percentile(tot[m], 90)))
    np.savez_compressed("gpurun_out/chain_stats.npz", a=a, g=g, ns=ns)
# pinned D2H / H2D bandwidth
h = torch.empty(84 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(84 << 20, dtype=torch.uint8, device="cuda")
for name, src, dst in (("D2H", d, h), ("H2D", h, d)):
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print("%s pinned 84 MiB: %.2f ms = %.1f GB/s" % (name, dt * 1e3, (84 << 20) / dt / 1e9))
