#!/usr/bin/env python
"""Development aid: per-stage cost of the lock-step step kernel (GRS_STEP_TIMING=1): mean warp time vs. what the block
barrier makes every warp pay (the per-round maximum)."""
import os, sys
os.environ["GRS_STEP_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
N = 4096
sim = GripperSim(make_config(sim_env="/xmls/acorn_env.xml"), num_envs=N)
gen = torch.Generator(device="cuda").manual_seed(0)
WARM = int(sys.argv[1]) if len(sys.argv) > 1 else 0  # agent steps before the measured ones (later states are contact-heavier)
for i in range(WARM + 12):
    if i == WARM:
        torch.cuda.synchronize(); sim.debug.zero_()
    sim.step(torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1)
torch.cuda.synchronize()
d = sim.debug.reshape(-1)[: 32 * 296].reshape(296, 32).cpu().numpy()
names = ["tick", "smooth", "constraint", "solve", "euler+kin", "crb", "collision"]
mean, mx = d[:, :7].sum(0), d[:, 8:15].sum(0)
print("rounds per block: mean %.0f" % d[:, 16].mean())
print("kcycles per round: warp mean %s | block max %s" % (" ".join("%s %.1f" % (n, a / d[:, 16].sum() / 1e3) for n, a in zip(names, d[:, :7].sum(0))), " ".join("%.1f" % (b / d[:, 16].sum() / 1e3) for b in d[:, 8:15].sum(0))))
for n, a, b in zip(names, mean, mx):
    print("%-12s mean-warp %6.1f%%   block-max %6.1f%%   max/mean %.2f" % (n, 100 * a / mean.sum(), 100 * b / mx.sum(), b / max(a, 1)))
lone = d[:, 18:25].sum(0); nl = d[:, 17].sum()
print("rounds with a single active warp in the block: %d of %d; cycles per such round %.0f (= %.1f us at 1.965 GHz); stage shares: %s" % (
    nl, d[:, 16].sum(), lone.sum() / max(nl, 1), lone.sum() / max(nl, 1) / 1965.0, ", ".join("%s %.0f%%" % (n, 100 * x / max(lone.sum(), 1)) for n, x in zip(names, lone))))
print("total: sum of per-round maxima / sum of warp means = %.2f" % (mx.sum() / mean.sum()))

# the blocks that run the longest chains (queued first: gripper-close environments, a fifth of which time out at 400 substeps)
order = d[:, 16].argsort()[::-1]
for name, sel in (("48 blocks with most rounds", order[:48]), ("48 blocks with fewest rounds", order[-48:])):
    r = d[sel, 16].sum()
    print("%s: rounds/block %.0f; kcycles per round, warp mean: %s | block max: %s | sum of max %.1f" % (
        name, d[sel, 16].mean(), " ".join("%s %.1f" % (n, a / r / 1e3) for n, a in zip(names, d[sel, :7].sum(0))),
        " ".join("%.1f" % (b / r / 1e3) for b in d[sel, 8:15].sum(0)), d[sel, 8:15].sum() / r / 1e3))

# block-level record: physics-phase duration, the number of rounds with k active warps and the cycles those rounds took
nsteps = 1  # the kernel overwrites its record: the numbers are those of the last agent step
bd = sim.debug.reshape(-1)[32 * 296: 64 * 296].reshape(296, 32).cpu().numpy()
dur, hist, cyc = bd[:, 0] / 1965.0e3, bd[:, 1:13], bd[:, 13:25]
last = dur.argsort()[::-1]
print("block physics-phase ms per agent step: mean %.2f p50 %.2f p90 %.2f max %.2f" % (dur.mean() / nsteps, np.median(dur) / nsteps, np.percentile(dur, 90) / nsteps, dur.max() / nsteps))
for name, sel in (("all blocks", np.arange(296)), ("16 blocks with the longest physics phase", last[:16])):
    h, c = hist[sel].sum(0), cyc[sel].sum(0)
    r = d[sel, 16].sum()
    print("%s: ms/step %.2f, rounds/step %.0f, us/round %.1f; rounds by active warps k = share (us per round): %s" % (
        name, dur[sel].mean() / nsteps, r / len(sel) / nsteps, 1e3 * dur[sel].sum() / max(r, 1),
        " ".join("%d = %.0f%% (%.0f)" % (k, 100 * h[k] / max(h.sum(), 1), c[k] / max(h[k], 1) / 1965.0) for k in range(1, 11) if h[k] > 0)))
print("slowest blocks (id, ms/step, rounds/step):", [(int(b), round(float(dur[b] / nsteps), 2), int(d[b, 16] / nsteps)) for b in last[:10]])
