#!/usr/bin/env python
"""Quick on-GPU comparison of the CUDA path with the oracle (development aid; the real checks are in tests/)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import oracle_model, golden  # noqa: E402
from oracle import engine  # noqa: E402
from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config  # noqa: E402


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "sugar_cube"
    cfg = make_config(sim_env="/xmls/%s_env.xml" % scene)
    t = time.time()
    sim = GripperSim(cfg, num_envs=8, auto_reset=False)
    print("create %.2fs" % (time.time() - t), "launches", sim.launch_count)
    om = oracle_model(sim)
    od = engine.Data(om)
    od.reset()
    st = sim.get_state()
    print("reset qpos diff", np.abs(st["qpos"][0] - od.qpos).max(), "warm diff", np.abs(st["warmstart"][0] - od.qacc_warmstart).max(), "warm scale", np.abs(od.qacc_warmstart).max())
    print("xfrc_z", st["xfrc_z"][0], "flags", st["flags"][0])
    c = sim.contacts()
    print("gpu ncon", c["ncon"][0], c["geom"][0][:c["ncon"][0]].tolist(), "oracle ncon", od.ncon, [(x["geom1"], x["geom2"]) for x in od.contacts()])
    # substep trajectory with a constant control
    ctrl = np.array([0.6, 0.2, -0.3, 0.1, 0.2, -1, -1], np.float32)
    sim.set_state(ctrl=ctrl)
    od.xfrc_applied[om.body_id("ee"), 2] = 0.438 * 9.81
    od.ctrl[:] = ctrl
    for k in range(6):
        dbg = sim.debug_step()
        # oracle intermediates are those of the step being taken
        od.step()
        s = sim.get_state()
        print("step %d: qpos err %.2e qvel err %.2e | ncon gpu %d orc %d | iters gpu %d orc %d | M err %.1e bias err %.1e qacc err %.1e (scale %.1e)" % (
            k, np.abs(s["qpos"][0] - od.qpos).max(), np.abs(s["qvel"][0] - od.qvel).max(), dbg["ncon"][0], -1, dbg["iters"][0], od.solver_iter,
            np.abs(dbg["M"][0] - od.qM).max(), np.abs(dbg["qfrc_bias"][0] - od.qfrc_bias).max(), np.abs(dbg["qacc"][0] - od.qacc).max(), np.abs(od.qacc).max()))
    t = time.time()
    sim.substep(200)
    sim.synchronize()
    print("200 substeps x 8 envs: %.3fs" % (time.time() - t))
    od.step(200)
    s = sim.get_state()
    print("after 206 substeps: qpos err %.2e qvel err %.2e ncon %d" % (np.abs(s["qpos"][0] - od.qpos).max(), np.abs(s["qvel"][0] - od.qvel).max(), od.ncon))
    print("   gpu qpos", s["qpos"][0]); print("   orc qpos", od.qpos)
    # agent steps against the golden rollout
    name = {"sugar_cube": "rollout_sugar_cube_dir0_seed0.npz", "sand_ball": "rollout_sand_ball_dir0_seed2.npz", "bread_crumb": "rollout_bread_crumb_dir0_seed3.npz"}.get(scene)
    if name:
        g = golden(name)
        sim.reset()
        acts = g["actions"]
        for i in range(len(acts)):
            a = torch.tensor(np.tile(acts[i], (8, 1)), device="cuda")
            sim.step(a)
            info = sim.info.cpu().numpy()[0]
            s = sim.get_state()
            nsub = int(info[10] + info[11] + info[12])
            print("astep %2d nsub gpu %4d gold %4d | reward gpu %.5f gold %.5f | qpos err %.2e | done %d/%d pad %s/%s grasped %d/%d" % (
                i, nsub, g["nsub"][i], info[0], g["reward"][i], np.abs(s["qpos"][0] - g["qpos"][i]).max(), info[1], g["done"][i],
                info[3:5].astype(int).tolist(), g["pad"][i].tolist(), info[5], g["object_grasped"][i]))
            if g["done"][i]:
                break
        print("obs sum", int(sim.obs.sum().item()), "obs[0] channel means", sim.obs[0].float().mean(dim=(1, 2)).tolist())
    # throughput probe
    N = 4096
    big = GripperSim(cfg, num_envs=N)
    gen = torch.Generator(device="cuda").manual_seed(0)
    for it in range(3):
        a = torch.rand((N, 6), device="cuda", generator=gen) * 2 - 1
        torch.cuda.synchronize(); t = time.time()
        big.step(a)
        torch.cuda.synchronize(); dt = time.time() - t
        info = big.info.cpu().numpy()
        nsub = info[:, 10:13].sum()
        print("N=%d step %d: %.1f ms, substeps %d (%.1f/env), %.2f M substeps/s, kernel %.2f ms, done %d, ncon max %d iters/substep %.2f flags %s" % (
            N, it, dt * 1e3, nsub, nsub / N, nsub / dt / 1e6, big.step_kernel_ms(), info[:, 1].sum(), info[:, 29].max(), info[:, 28].sum() / max(nsub, 1),
            np.unique(info[:, 30]).tolist()))


if __name__ == "__main__":
    main()
