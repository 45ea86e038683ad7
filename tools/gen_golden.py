#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the reference's UNMODIFIED python (RobotEnv, Actuator, Reward,
IntrinsicReward, utils, transformations) in this container.

What is pinned by these fixtures, and what is not:
  * PINNED to the reference's own code: the agent-step state machine (robot_env.py:77-241), the controller
    (actuator.py), both rewards (reward.py), transform_depth / projection (utils.py), the observation packing
    (robot_env.py:275-293).  They are executed for real, imported from /root/reference.
  * NOT pinned: the physics underneath.  `dm_control.mujoco.Physics` is replaced by oracle/fake_dm_control.py,
    which drives oracle/engine.c (a restatement of MuJoCo; MuJoCo itself is unavailable here).

Run:  python tools/gen_golden.py            (needs /root/reference; the GPU box only reads the .npz files)
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import fake_dm_control  # noqa: E402

fake_dm_control.install()
sys.path.insert(0, REF)
from simulation.environment.robot_env import RobotEnv  # noqa: E402
from simulation.environment.reward import Reward, IntrinsicReward  # noqa: E402
from simulation.utils import utils as ref_utils  # noqa: E402
from simulation.utils import transformations as ref_tf  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make_config(scene, direction, **kw):
    c = argparse.Namespace(sim_env="/xmls/%s_env.xml" % scene, width_capture=64, height_capture=64,
                           full_observation=True, camera_id=3, show_obs=False, max_rotation=0.15,
                           max_translation=0.05, grasp_tolerance=0.03, pos_tolerance=0.002, include_roll=True,
                           max_steps=400, im_reward=False, her_buffer=False, direction=direction, time_horizon=400,
                           rendering_zoom_width=10, rendering_zoom_height=7.5)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def action_tape(seed, n, push_bias=True):
    """Seeded U(-1,1)^6 tape; the first component is biased towards +x so the gripper reaches the object and
    the open/close channel alternates in blocks so both gripper phases are exercised."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, (n, 6))
    if push_bias:
        a[:, 0] = np.abs(a[:, 0])
        a[:, 5] = np.where((np.arange(n) // 3) % 2 == 0, -np.abs(a[:, 5]), np.abs(a[:, 5]))
    return a.astype(np.float32)


def rollout(scene, direction, seed, nsteps, **cfgkw):
    cfg = make_config(scene, direction, **cfgkw)
    env = RobotEnv(cfg)
    phys = env.physics
    obs0 = env.reset()
    acts = action_tape(seed, nsteps)
    if not cfg.include_roll:
        acts = acts[:, [0, 1, 2, 4, 5]]
    rec = dict(actions=acts, qpos=[], qvel=[], ctrl=[], reward=[], done=[], status=[], pad=[], achieved=[], desired=[],
               nsub=[], object_grasped=[], gripper_open=[], reached=[], total_distance=[], line_distance=[],
               target_qpos=[], ncon=[], contact_geoms=[])
    rec["reset_pad"] = obs0["observation"][-1, 0, :2].copy()
    rec["reset_achieved"] = obs0["achieved_goal"].copy()
    rec["reset_desired"] = obs0["desired_goal"].copy()
    rec["reset_qpos"] = phys.data.qpos.copy()
    orig_target = env._actuator.get_target_pose
    last = {}

    def spy(action):
        t = orig_target(action)
        last["t"] = np.array(t, dtype=np.float64)
        return t
    env._actuator.get_target_pose = spy
    for k in ("pre_qpos", "pre_qvel", "pre_ctrl", "pre_warmstart", "pre_gripper_open", "pre_episode_step"):
        rec[k] = []
    for a in acts:
        rec["pre_qpos"].append(phys.data.qpos.copy()); rec["pre_qvel"].append(phys.data.qvel.copy())
        rec["pre_ctrl"].append(phys.data.ctrl.copy()); rec["pre_warmstart"].append(phys._d.qacc_warmstart.copy())
        rec["pre_gripper_open"].append(bool(env.gripper_open)); rec["pre_episode_step"].append(int(env.episode_step))
        n0 = fake_dm_control.Physics.step_count
        obs, r, done, info = env.step(a.astype(np.float64))
        rec["nsub"].append(fake_dm_control.Physics.step_count - n0)
        rec["qpos"].append(phys.data.qpos.copy()); rec["qvel"].append(phys.data.qvel.copy())
        rec["ctrl"].append(phys.data.ctrl.copy())
        rec["reward"].append(float(r)); rec["done"].append(bool(done)); rec["status"].append(info["status"].value)
        rec["pad"].append(obs["observation"][-1, 0, :2].copy())
        rec["achieved"].append(obs["achieved_goal"].copy()); rec["desired"].append(obs["desired_goal"].copy())
        rec["object_grasped"].append(info["object_grasped"]); rec["gripper_open"].append(bool(info["gripper_open"]))
        pr = info["position_reached"]
        rec["reached"].append([pr["target"], pr["initial"], pr["fail"]])
        rec["total_distance"].append(info["total_distance"]); rec["line_distance"].append(float(info["line_distance"]))
        rec["target_qpos"].append(last["t"])
        cons = phys._d.contacts()
        rec["ncon"].append(len(cons))
        g = np.full((16, 2), -1, np.int32)
        for i, c in enumerate(cons[:16]):
            g[i] = (c["geom1"], c["geom2"])
        rec["contact_geoms"].append(g)
        if done:
            env.reset()
    out = {k: np.array(v) for k, v in rec.items()}
    out["scene"], out["direction"] = np.array(scene), np.array(direction)
    return out


def unit_vectors():
    """Known-answer vectors for the pure functions, straight from the reference modules."""
    rng = np.random.default_rng(7)
    out = {}
    # Reward.agent_reward (reward.py:18-41)
    R = Reward(robot=None, config=None)
    n = 64
    init = rng.uniform(-0.2, 0.2, (n, 3)); final = init + rng.uniform(-0.05, 0.12, (n, 3))
    dirs = np.array([[1, 0], [1, 1]])[rng.integers(0, 2, n)]
    gopen = rng.integers(0, 2, n).astype(bool); ctr = rng.choice([0.0, -1.0, 0.5], (n, 2)); grasp = rng.integers(0, 4, n)
    out["rw_init"], out["rw_final"], out["rw_dir"], out["rw_open"], out["rw_ctrl"], out["rw_grasp"] = init, final, dirs, gopen, ctr, grasp
    out["rw_out"] = np.array([R.agent_reward(init[i], final[i], dirs[i], gopen[i], ctr[i], grasp[i]) for i in range(n)])
    # IntrinsicReward.intrinsic_reward (reward.py:57-77), full and rgb-only observations
    imgs_a = rng.integers(0, 256, (6, 5, 64, 64)).astype(np.uint8)
    imgs_b = imgs_a.copy()
    imgs_b[:, :, 10:40, 5:50] = rng.integers(0, 64, (6, 5, 30, 45)).astype(np.uint8)
    imgs_b[0] = imgs_a[0]
    imgs_a[1, :4] = 17  # degenerate single-bin histograms
    for full in (True, False):
        IR = IntrinsicReward(robot=None, config=argparse.Namespace(full_observation=full))
        key = "ir_full" if full else "ir_rgb"
        out[key] = np.array([IR.intrinsic_reward(imgs_a[i] if full else imgs_a[i][[0, 1, 2, 4]],
                                                 imgs_b[i] if full else imgs_b[i][[0, 1, 2, 4]]) for i in range(6)])
    out["ir_a"], out["ir_b"] = imgs_a, imgs_b
    # transform_depth (utils.py:11-19)
    dep = rng.uniform(0.05, 4.0, (4, 64, 64)).astype(np.float32)
    dep[1] += 1.5
    out["td_in"] = dep.copy()
    out["td_out"] = np.array([ref_utils.transform_depth(d.copy()) for d in dep])
    out["td_u8"] = out["td_out"].astype(np.uint8)
    # euler_from_quaternion 'rzyx' (transformations.py:1093) and compose/euler_from_matrix round trip
    q = rng.normal(size=(32, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    out["eq_in"] = q
    out["eq_out"] = np.array([ref_tf.euler_from_quaternion(x, axes=(0, 0, 0, 1)) for x in q])
    ang = rng.uniform(-1.5, 1.5, (32, 3)); tr = rng.uniform(-1, 1, (32, 3))
    out["cm_ang"], out["cm_tr"] = ang, tr
    out["cm_out"] = np.array([ref_tf.compose_matrix(angles=ang[i], translate=tr[i]) for i in range(32)])
    out["em_out"] = np.array([ref_tf.euler_from_matrix(m) for m in out["cm_out"]])
    # project_to_target_direction (utils.py:30-31)
    p = rng.uniform(-1, 1, (16, 2))
    out["pj_in"] = p
    out["pj_out0"] = np.array([ref_utils.project_to_target_direction(x, np.array([1, 0])) for x in p])
    out["pj_out45"] = np.array([ref_utils.project_to_target_direction(x, np.array([1, 1])) for x in p])
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "unit_vectors.npz"), **unit_vectors())
    print("unit vectors done")
    jobs = [("sugar_cube", 0, 0, 40, {}), ("sugar_cube", 45, 1, 40, {}), ("sand_ball", 0, 2, 40, {}),
            ("bread_crumb", 0, 3, 30, {}), ("sand_ball", 45, 4, 30, dict(include_roll=False)),
            ("sugar_cube", 0, 5, 30, dict(her_buffer=True, time_horizon=12))]
    for scene, direction, seed, n, kw in jobs:
        tag = "%s_dir%d_seed%d" % (scene, direction, seed) + ("".join("_%s" % k for k in kw))
        r = rollout(scene, direction, seed, n, **kw)
        r["cfg_keys"] = np.array(list(kw.keys()))
        r["cfg_vals"] = np.array([float(v) for v in kw.values()])
        np.savez_compressed(os.path.join(OUT, "rollout_%s.npz" % tag), **r)
        print(tag, "substeps", int(r["nsub"].sum()), "reward", float(r["reward"].sum()), "done", int(r["done"].sum()),
              "grasp", r["pad"][:, 0].max())


if __name__ == "__main__":
    main()
