"""CPU: pin the oracle.  The C restatement of RobotEnv / Actuator / Reward (oracle/engine.c, environment layer) must
reproduce the fixtures that tools/gen_golden.py recorded by running the reference's UNMODIFIED python on the same
engine, and the pure functions must reproduce the reference's own outputs (unit_vectors.npz)."""
import glob
import os

import numpy as np
import pytest

from helpers import GOLDEN, scene_file
from oracle import engine, mjcf, envmath

ROLLOUTS = sorted(glob.glob(os.path.join(GOLDEN, "rollout_*.npz")))


def _cfg_of(g):
    kw = {}
    for k, v in zip(g["cfg_keys"].tolist(), g["cfg_vals"].tolist()):
        kw[str(k)] = int(v)
    return kw


@pytest.mark.parametrize("path", ROLLOUTS, ids=[os.path.basename(p)[8:-4] for p in ROLLOUTS])
def test_c_env_reproduces_reference_python(path):
    g = np.load(path)
    scene, direction = str(g["scene"]), int(g["direction"])
    kw = _cfg_of(g)
    model = engine.Model(mjcf.compile_mjcf(scene_file(scene)))
    env = engine.Env(model, direction=direction, **kw)
    out = env.reset()
    assert [out.grasp, out.pheromone] == g["reset_pad"].tolist()
    np.testing.assert_array_equal(np.float32(out.achieved_goal[:]), g["reset_achieved"])
    np.testing.assert_array_equal(np.float32(out.desired_goal[:]), g["reset_desired"])
    np.testing.assert_allclose(env.qpos, g["reset_qpos"], rtol=0, atol=0)
    qerr = []
    for i, a in enumerate(g["actions"]):
        # matched states: every agent step starts from the recorded pre-state (last-digit differences, e.g. SVD pinv in
        # numpy vs normal equations in C, would otherwise be amplified by contact chaos over the rollout)
        env.qpos[:] = g["pre_qpos"][i]
        env.data.qvel[:] = g["pre_qvel"][i]
        env.ctrl[:] = g["pre_ctrl"][i]
        env.data.qacc_warmstart[:] = g["pre_warmstart"][i]
        env.p.contents.gripper_open = int(g["pre_gripper_open"][i])
        env.p.contents.episode_step = int(g["pre_episode_step"][i])
        env.p.contents.status = 0
        env.data.forward_position()
        o = env.step(a.astype(np.float64))
        assert o.nsub_a + o.nsub_b + o.nsub_c == g["nsub"][i], "substep count, step %d" % i
        np.testing.assert_allclose(np.array(o.target_qpos[:]), g["target_qpos"][i], rtol=0, atol=1e-11)
        # contact on polyhedral hulls is a discontinuous map (support-vertex flips): a 1e-16 difference in the IK target
        # occasionally becomes 1e-3 within one grasping step, so the state tolerance is two-tier (see the end of the test)
        qerr.append(np.abs(env.qpos - g["qpos"][i]).max())
        assert qerr[-1] < 1e-2, "step %d" % i
        assert abs(o.reward - g["reward"][i]) < 1e-4
        assert bool(o.done) == bool(g["done"][i]) and o.status == g["status"][i]
        if qerr[-1] < 1e-7:  # contact-derived flags are only comparable when no contact flip happened in the step
            assert [o.grasp, o.pheromone] == g["pad"][i].tolist()
            assert o.object_grasped == g["object_grasped"][i]
        assert bool(o.gripper_open) == bool(g["gripper_open"][i])
        assert [bool(o.reached_target), bool(o.reached_initial), bool(o.fail)] == g["reached"][i].tolist()
        tol = 1e-6 if qerr[-1] < 1e-7 else 1e-2
        np.testing.assert_allclose(np.float32(o.achieved_goal[:]), g["achieved"][i], rtol=0, atol=max(tol, 1e-6))
        np.testing.assert_allclose(np.float32(o.desired_goal[:]), g["desired"][i], rtol=0, atol=max(tol, 1e-6))
        assert abs(o.total_distance - g["total_distance"][i]) < tol and abs(o.line_distance - g["line_distance"][i]) < tol
        if qerr[-1] < 1e-7:
            assert env.data.ncon == g["ncon"][i]
    qerr = np.array(qerr)
    assert (qerr < 1e-7).mean() >= 0.8, qerr
    assert np.median(qerr) < 1e-12


def test_golden_covers_the_state_machine():
    """The fixtures must exercise every branch worth pinning: contact with the object, both gripper phases,
    phase B (return to the initial pose), time-limit termination, the 5-dim action space and the HER term."""
    seen = dict(contact=False, close=False, open=False, phase_b=False, time_limit=False, reward=False, grasp_pad=False)
    for p in ROLLOUTS:
        g = np.load(p)
        seen["reward"] |= bool((g["reward"] > 0).any())
        seen["grasp_pad"] |= bool((g["pad"][:, 0] > 0).any())
        seen["time_limit"] |= bool((g["status"] == 2).any())
        seen["phase_b"] |= bool(g["reached"][:, 1].any())
        go = np.r_[True, g["gripper_open"][:-1]]
        seen["close"] |= bool((go & ~g["gripper_open"]).any())
        seen["open"] |= bool((~go & g["gripper_open"]).any())
        seen["contact"] |= bool((g["contact_geoms"][:, :, 1] == 6).any() and (g["contact_geoms"][:, :, 0] > 0).any())
    assert all(seen.values()), seen


def test_pure_functions_against_reference_outputs():
    u = np.load(os.path.join(GOLDEN, "unit_vectors.npz"))
    # Reward.agent_reward (reward.py:18-41)
    L = engine.lib()
    for i in range(len(u["rw_out"])):
        init, fin, d, c = [np.ascontiguousarray(u[k][i], dtype=np.float64) for k in ("rw_init", "rw_final", "rw_dir", "rw_ctrl")]
        r = L.orc_agent_reward(init.ctypes.data, fin.ctypes.data, d.ctypes.data, int(u["rw_open"][i]), c.ctypes.data, int(u["rw_grasp"][i]))
        assert abs(r - u["rw_out"][i]) < 1e-12
    # IntrinsicReward.intrinsic_reward (reward.py:57-77)
    for i in range(6):
        assert abs(envmath.intrinsic_reward(u["ir_a"][i], u["ir_b"][i], True) - u["ir_full"][i]) < 1e-6
        assert abs(envmath.intrinsic_reward(u["ir_a"][i][[0, 1, 2, 4]], u["ir_b"][i][[0, 1, 2, 4]], False) - u["ir_rgb"][i]) < 1e-6
    # transform_depth (utils.py:11-19)
    for i in range(4):
        out = envmath.transform_depth(u["td_in"][i])
        np.testing.assert_allclose(out, u["td_out"][i], rtol=0, atol=2e-4)
        assert (out.astype(np.uint8) != u["td_u8"][i]).mean() < 1e-3
    # euler_from_quaternion(axes=(0,0,0,1)) (transformations.py:1093) -> (yaw, pitch, roll)
    for q, e in zip(u["eq_in"], u["eq_out"]):
        np.testing.assert_allclose(envmath.euler_rzyx_from_quat(q), e, rtol=0, atol=1e-12)
    # compose_matrix / euler_from_matrix sxyz (transformations.py:789, 1035)
    for a, t, M, e in zip(u["cm_ang"], u["cm_tr"], u["cm_out"], u["em_out"]):
        np.testing.assert_allclose(envmath.compose_matrix(a, t), M, rtol=0, atol=1e-12)
        np.testing.assert_allclose(envmath.euler_sxyz_from_matrix(M), e, rtol=0, atol=1e-12)
    for p, a, b in zip(u["pj_in"], u["pj_out0"], u["pj_out45"]):
        assert abs(envmath.project(p, [1, 0]) - a) < 1e-15 and abs(envmath.project(p, [1, 1]) - b) < 1e-15


def test_oracle_physics_sanity():
    """Analytic properties the engine must satisfy whatever MuJoCo does in the last digit."""
    md = mjcf.compile_mjcf(scene_file("sand_ball"))
    m = engine.Model(md)
    d = engine.Data(m)
    d.reset()
    # free fall of the object (sand_ball spawns above the floor): semi-implicit Euler, v_k = -g h k, z_k = z_0 - g h^2 k(k+1)/2
    z0 = d.qpos[9]
    for k in range(1, 6):
        d.step()
        if d.ncon:
            break
        assert abs(d.qvel[9] + 9.81 * 0.002 * k) < 1e-12
        assert abs(d.qpos[9] - (z0 - 9.81 * 0.002 ** 2 * k * (k + 1) / 2)) < 1e-12
    # gripper mass from the mesh volumes (0.44719 kg, SURVEY.md §8 a-M) and gravity compensation literal
    names = md["body_names"]
    gm = sum(md["body_mass"][names.index(n)] for n in ("ee", "robotiq_85_base_link", "left_inner_knuckle", "left_inner_finger",
                                                       "right_inner_knuckle", "right_inner_finger"))
    assert abs(gm - 0.44719) < 2e-4
    # determinism
    e1, e2 = engine.Env(m), engine.Env(m)
    e1.reset(), e2.reset()
    rng = np.random.default_rng(0)
    for _ in range(5):
        a = rng.uniform(-1, 1, 6)
        e1.step(a), e2.step(a)
    assert np.array_equal(e1.qpos, e2.qpos) and np.array_equal(e1.data.qvel, e2.data.qvel)


def test_free_body_conserves_angular_momentum():
    """A torque-free rigid body: the object of sugar_cube (principal inertias 0.0322 / 0.0307 / 0.0281, tilted inertial frame) spun
    about an axis that is no principal axis, far above the floor.  World-frame angular momentum R I R^T w and rotational energy must
    stay constant up to the integrator's O(h) error — this pins the gyroscopic term of the RNE, the body-frame convention of the
    free joint's angular velocity and the quaternion integration, whatever MuJoCo does in the last digit."""
    def q2m(q):
        w, x, y, z = q
        return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])
    md = mjcf.compile_mjcf(scene_file("sugar_cube"))
    m = engine.Model(md)
    ob = md["body_names"].index("object")
    inertia = np.asarray(md["body_inertia"]).reshape(-1, 3)[ob]
    Ri = q2m(np.asarray(md["body_iquat"]).reshape(-1, 4)[ob])
    Ib = Ri @ np.diag(inertia) @ Ri.T
    d = engine.Data(m)
    d.reset()
    d.qpos[9] = 5.0                      # 5 m up: the object touches nothing for the 0.6 s of the test
    d.qvel[10:13] = [3.0, 1.0, 2.0]
    d.forward_position()
    L, E = [], []
    for _ in range(300):
        R, w = q2m(d.qpos[10:14]), d.qvel[10:13].copy()
        L.append(R @ (Ib @ w))
        E.append(0.5 * w @ Ib @ w)
        d.step()
    L, E = np.array(L), np.array(E)
    assert np.abs(L - L[0]).max() / np.linalg.norm(L[0]) < 1e-3      # measured 2.9e-4
    assert abs(E[-1] - E[0]) / E[0] < 5e-4                           # measured 4.5e-5
    assert np.abs(np.diff(d.qvel[10:13] - [3.0, 1.0, 2.0])).max() > 1e-3   # the body does tumble: w itself is NOT constant


def test_gravity_compensation_leaves_the_documented_residual():
    """robot_env.py:64-65 holds the gripper up with a constant +z force of 0.438 * 9.81 N on `ee`; the meshes weigh 0.44719 kg, so 9.2 g
    of weight remain and the gripper sinks at the terminal velocity of its damped z slide: (m - 0.438) g / damping(gripper_z = 20).
    Pins the mesh masses, the xfrc term of the smooth forces and the implicit joint damping in one number."""
    md = mjcf.compile_mjcf(scene_file("acorn"))
    m = engine.Model(md)
    names = md["body_names"]
    gm = sum(md["body_mass"][names.index(n)] for n in ("ee", "robotiq_85_base_link", "left_inner_knuckle", "left_inner_finger",
                                                       "right_inner_knuckle", "right_inner_finger"))
    d = engine.Data(m)
    d.reset()
    d.xfrc_applied[m.body_id("ee"), 2] = 0.438 * 9.81
    d.step(400)   # 0.8 s = 36 time constants (m / damping = 22 ms)
    assert abs(d.qvel[2] + (gm - 0.438) * 9.81 / 20.0) < 1e-9
    assert abs(d.qvel[:2]).max() < 1e-9 and abs(d.qvel[3:7]).max() < 1e-6
