#!/usr/bin/env python
"""Development aid: how predictable is an environment's substep count (its chain length) before the agent step?
Per chain-length class of the queue order (close gripper | open | arm only): share of long chains, and how well the previous
agent step's counts / the state predict them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
from mujoco_rl_manipulate_unknown_objects_b200._native import INFO as I
scene = sys.argv[1] if len(sys.argv) > 1 else "acorn"
N, PRE, K = 4096, int(os.environ.get("PRE", 150)), int(os.environ.get("K", 30))
sim = GripperSim(make_config(sim_env="/xmls/%s_env.xml" % scene), num_envs=N)
gen = torch.Generator(device="cuda").manual_seed(0)
prev = None
rows = []
for i in range(PRE + K):
    a = torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1
    st = sim.get_state() if i >= PRE else None
    sim.step(a)
    info = sim.info.cpu().numpy()
    ns = info[:, I["NSUB_A"]:I["NSUB_A"] + 3].copy()
    if i >= PRE and prev is not None:
        an = a.cpu().numpy()
        go = st["flags"] if "flags" in st else None
        rows.append(dict(ns=ns, prev=prev, act=an, qpos=st["qpos"], done=info[:, I["DONE"]].copy(), gopen=st.get("gripper_open")))
    prev = ns
ns = np.concatenate([r["ns"] for r in rows]); pv = np.concatenate([r["prev"] for r in rows]); act = np.concatenate([r["act"] for r in rows])
qpos = np.concatenate([r["qpos"] for r in rows])
tot, ptot = ns.sum(1), pv.sum(1)
print("keys of get_state:", list(st.keys()))
print("all: mean chain %.0f, share >= 400: %.3f; by phase: A>=400 %.3f, B>0 %.3f, C>=400 %.3f" % (tot.mean(), (tot >= 400).mean(), (ns[:, 0] >= 400).mean(), (ns[:, 1] > 0).mean(), (ns[:, 2] >= 400).mean()))
armonly = ns[:, 2] == 0
print("arm-only envs (no gripper phase): share %.3f, of which chain >= 300: %.4f" % (armonly.mean(), (tot[armonly] >= 300).mean()))
L = armonly & (tot >= 300)
print("  long arm-only: prev step also A-timeout %.3f (base rate %.4f); z of ee qpos[2]: mean %.3f vs %.3f overall; action z mean %.2f vs %.2f" % (
    (pv[L, 0] >= 400).mean(), (pv[armonly, 0] >= 400).mean(), qpos[L, 2].mean(), qpos[armonly, 2].mean(), act[L, 2].mean(), act[armonly, 2].mean()))
for name, cond in (("prev A-timeout", pv[:, 0] >= 400), ("prev chain >= 300", ptot >= 300), ("ee low (qpos z < p10) & action z < 0", (qpos[:, 2] < np.percentile(qpos[:, 2], 10)) & (act[:, 2] < 0))):
    c = cond & armonly
    print("  predictor '%s': flags %.3f of arm-only envs, precision %.3f, recall %.3f" % (name, c.sum() / armonly.sum(), (tot[c] >= 300).mean() if c.any() else 0, (c & L).sum() / max(L.sum(), 1)))
