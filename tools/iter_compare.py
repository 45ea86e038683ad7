#!/usr/bin/env python
"""Development aid: Newton iterations per substep, fp32 CUDA solver vs the fp64 oracle (tolerance 1e-10) at matched states."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import test_physics_parity as T
for scene in ("acorn", "sugar_cube"):
    N = 96
    sim, om = T.get_sim(scene, N)
    states = T.visited_states(om, seed=11, nsteps=N // 3)[:N]
    q, v, c, w = T.to_f32_states(states)
    sim.reset(); sim.set_state(qpos=q, qvel=v, ctrl=c, warmstart=w)
    dbg = sim.debug_step()
    gi, oi, nc = [], [], []
    for i in range(N):
        d = T.oracle_at(om, q[i], v[i], c[i], w[i]); nc.append(len(d.contacts())); d.step(); oi.append(d.solver_iter); gi.append(int(dbg["iters"][i]))
    gi, oi, nc = np.array(gi), np.array(oi), np.array(nc)
    print(scene, "mean iters gpu %.2f oracle %.2f | gpu>oracle in %d, gpu<oracle in %d of %d states | by contact count:" % (gi.mean(), oi.mean(), (gi > oi).sum(), (gi < oi).sum(), N),
          {int(k): (round(float(gi[nc == k].mean()), 2), round(float(oi[nc == k].mean()), 2)) for k in np.unique(nc)})
