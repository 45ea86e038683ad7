#!/usr/bin/env python
"""Dump MuJoCo golden vectors from the UNMODIFIED reference (SURVEY.md §8c "required deliverable").

Cannot run in the build container (mujoco / dm_control / gym are not installable there).  On any machine with
    pip install dm-control==1.0.3.post1 "mujoco==2.2.1" gym==0.21.0 numpy scipy scikit-learn opencv-python
run
    MUJOCO_GL=egl python tools/dump_mujoco_golden.py --reference /path/to/mujoco_rl_manipulate_unknown_objects --out tests/golden/mujoco
For each compilable scene x direction x seeded action tape it records, per SUBSTEP: qpos, qvel, ctrl, ncon, (geom1, geom2,
dist) of every contact; per AGENT STEP: observation, reward, done, the info record; plus the compiled model constants and
mujoco.__version__.  tests/test_physics_parity.py / test_env_parity.py consume the files when present and otherwise report
"GOLDEN ABSENT - parity vs MuJoCo unverified".
"""
import argparse
import os
import sys

import numpy as np

SCENES = ("sugar_cube", "sand_ball", "bread_crumb")  # acorn.stl is absent from the reference tree


def action_tape(seed, n):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, (n, 6))
    a[:, 0] = np.abs(a[:, 0])  # push towards the object so that contacts happen
    a[:, 5] = np.where((np.arange(n) // 3) % 2 == 0, -np.abs(a[:, 5]), np.abs(a[:, 5]))
    return a.astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--seeds", type=int, nargs="*", default=[0, 1])
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    import mujoco  # noqa: F401
    from config.base_config import BaseConfig  # noqa: F401  (reference)
    from simulation.environment.robot_env import RobotEnv
    os.makedirs(args.out, exist_ok=True)
    for scene in SCENES:
        for direction in (0, 45):
            for seed in args.seeds:
                cfg = argparse.Namespace(sim_env="/xmls/%s_env.xml" % scene, width_capture=64, height_capture=64, full_observation=True,
                                         camera_id=3, show_obs=False, max_rotation=0.15, max_translation=0.05, grasp_tolerance=0.03,
                                         pos_tolerance=0.002, include_roll=True, max_steps=400, im_reward=False, her_buffer=False,
                                         direction=direction, time_horizon=400, rendering_zoom_width=10, rendering_zoom_height=7.5)
                env = RobotEnv(cfg)
                phys = env.physics
                sub = dict(qpos=[], qvel=[], ctrl=[], ncon=[], contacts=[], agent_step=[])
                step_orig = phys.step
                cur = [0]

                def step_logged(*a, **k):
                    r = step_orig(*a, **k)
                    d = phys.data
                    sub["qpos"].append(d.qpos.copy()); sub["qvel"].append(d.qvel.copy()); sub["ctrl"].append(d.ctrl.copy())
                    sub["ncon"].append(int(d.ncon))
                    sub["contacts"].append(np.array([[c.geom1, c.geom2, c.dist] for c in d.contact[:d.ncon]], dtype=np.float64).reshape(-1, 3))
                    sub["agent_step"].append(cur[0])
                    return r

                phys.step = step_logged
                obs0 = env.reset()
                tape = action_tape(seed, args.steps)
                rec = dict(obs=[obs0["observation"]], reward=[], done=[], ag=[obs0["achieved_goal"]], dg=[obs0["desired_goal"]], info=[])
                for t, a in enumerate(tape):
                    cur[0] = t
                    o, r, done, info = env.step(a)
                    rec["obs"].append(o["observation"]); rec["reward"].append(r); rec["done"].append(done)
                    rec["ag"].append(o["achieved_goal"]); rec["dg"].append(o["desired_goal"])
                    rec["info"].append([info["total_distance"], info["line_distance"], info["object_grasped"], info["gripper_open"], info["episode_step"]])
                    if done:
                        break
                m = phys.model
                ncmax = max([c.shape[0] for c in sub["contacts"]] + [1])
                con = np.full((len(sub["contacts"]), ncmax, 3), np.nan)
                for i, c in enumerate(sub["contacts"]):
                    con[i, :c.shape[0]] = c
                path = os.path.join(args.out, "mujoco_%s_dir%d_seed%d.npz" % (scene, direction, seed))
                np.savez_compressed(
                    path, mujoco_version=mujoco.__version__, scene=scene, direction=direction, seed=seed, actions=tape,
                    sub_qpos=np.array(sub["qpos"]), sub_qvel=np.array(sub["qvel"]), sub_ctrl=np.array(sub["ctrl"]), sub_ncon=np.array(sub["ncon"]),
                    sub_contacts=con, sub_agent_step=np.array(sub["agent_step"]), obs=np.array(rec["obs"]), reward=np.array(rec["reward"]),
                    done=np.array(rec["done"]), achieved_goal=np.array(rec["ag"]), desired_goal=np.array(rec["dg"]), info=np.array(rec["info"], dtype=np.float64),
                    body_mass=m.body_mass.copy(), body_inertia=m.body_inertia.copy(), body_ipos=m.body_ipos.copy(), body_iquat=m.body_iquat.copy(),
                    body_invweight0=m.body_invweight0.copy(), dof_invweight0=m.dof_invweight0.copy(), geom_rbound=m.geom_rbound.copy(),
                    qpos0=m.qpos0.copy(), stat_extent=m.stat.extent, stat_meaninertia=m.stat.meaninertia)
                print("wrote", path)


if __name__ == "__main__":
    main()
