"""GPU: reset randomisation.  The reference has none — RobotEnv.reset always starts from qpos0
(robot_env.py:56-75) — so the default (noise 0) must stay bit-identical to the deterministic reset, and the optional
noise is checked against its own specification: a counter-based hash of (seed, environment, episode) that moves the
object by U(-xy, xy)^2 and turns it by U(-yaw, yaw) about the vertical; the observation of the new episode is the
render of exactly that state; explicit and automatic resets draw from the same stream."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def reset_uniform(seed, env, episode, k):
    """Mirror of reset_uniform in csrc/env_kernels.cuh (uint32 arithmetic)."""
    M = 0xFFFFFFFF
    h = ((seed * 0x9E3779B1) & M) ^ ((env * 0x85EBCA77) & M) ^ ((episode * 0xC2B2AE3D) & M) ^ ((k * 0x27D4EB2F) & M)
    h ^= h >> 15; h = (h * 0x2C1B3C6D) & M; h ^= h >> 12; h = (h * 0x297A2D39) & M; h ^= h >> 15
    return np.float32(np.float32(h >> 8) * np.float32(2.0 / 16777216.0) - np.float32(1.0))


def expected_pose(q0, seed, env, episode, nxy, nyaw):
    dx, dy = np.float32(nxy) * reset_uniform(seed, env, episode, 0), np.float32(nxy) * reset_uniform(seed, env, episode, 1)
    yaw = float(np.float32(nyaw) * reset_uniform(seed, env, episode, 2))
    qz = np.array([np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)])
    w0, x0, y0, z0 = q0[10:14]
    q = np.array([qz[0] * w0 - qz[3] * z0, qz[0] * x0 - qz[3] * y0, qz[0] * y0 + qz[3] * x0, qz[0] * z0 + qz[3] * w0])
    return np.array([q0[7] + dx, q0[8] + dy]), q / np.linalg.norm(q)


def _obs_matches_render(sim, full=True):
    from oracle import render
    from mujoco_rl_manipulate_unknown_objects_b200._native import INFO as I
    obs = sim.obs.cpu().numpy()
    rgb, depth = sim.render(camera_id=2, width=64, height=64)
    rgb, depth = rgb.cpu().numpy(), depth.cpu().numpy()
    for e in range(obs.shape[0]):
        ref = render.observation(rgb[e], depth[e], 0, obs[e, 4, 0, 1], full)
        assert (obs[e, :3] != ref[:3]).mean() < 1e-3  # two kernels, same arithmetic: at most a stray rounding flip
        assert np.abs(obs[e, 3].astype(int) - ref[3].astype(int)).max() <= 1
        assert obs[e, 4, 0, 0] == 0 and obs[e, 4].sum() == obs[e, 4, 0, :2].sum()


def test_default_is_the_reference_reset():
    from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
    sim = GripperSim(make_config(sim_env="/xmls/sugar_cube_env.xml"), num_envs=16)
    st = sim.get_state()
    assert np.all(st["qpos"] == st["qpos"][0]) and bool((sim.obs == sim.reset_obs).all())
    sim.close()


def test_reset_noise_follows_its_specification():
    import torch
    from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
    N, NXY, NYAW, SEED, H = 48, 0.04, 0.6, 12345, 2
    ref = GripperSim(make_config(sim_env="/xmls/sugar_cube_env.xml"), num_envs=1)
    q0 = ref.get_state()["qpos"][0].astype(np.float64)
    ref.close()
    cfg = make_config(sim_env="/xmls/sugar_cube_env.xml", reset_noise_xy=NXY, reset_noise_yaw=NYAW, seed=SEED, time_horizon=H)
    sim = GripperSim(cfg, num_envs=N, auto_reset=True)
    # --- episode 1: the reset inside the constructor
    q = sim.get_state()["qpos"]
    assert np.abs(q[:, 7:9] - q0[7:9]).max() <= NXY + 1e-6 and np.unique(q[:, 7].round(6)).size > N // 2
    np.testing.assert_array_equal(q[:, :7], np.tile(np.float32(q0[:7]), (N, 1)))  # the gripper is untouched
    for e in range(N):
        xy, quat = expected_pose(q0, SEED, e, 1, NXY, NYAW)
        np.testing.assert_allclose(q[e, 7:9], xy, atol=2e-7)
        np.testing.assert_allclose(q[e, 10:14], quat, atol=2e-6)
    np.testing.assert_allclose(sim.achieved_goal.cpu().numpy(), q[:, 7:9], atol=1e-7)  # robot_env.py:71
    _obs_matches_render(sim)
    assert not bool((sim.obs == sim.reset_obs).all())
    # --- automatic reset at the time limit: episode 2 draws new poses, the returned observation shows them
    a = torch.zeros((N, 6), device="cuda")
    for _ in range(H):
        sim.step(a)
    assert bool(sim.done.all())
    q2 = sim.get_state()["qpos"]
    for e in range(0, N, 5):
        xy, quat = expected_pose(q0, SEED, e, 2, NXY, NYAW)
        np.testing.assert_allclose(q2[e, 7:9], xy, atol=2e-7)
        np.testing.assert_allclose(q2[e, 10:14], quat, atol=2e-6)
    np.testing.assert_allclose(sim.achieved_goal.cpu().numpy(), q2[:, 7:9], atol=1e-7)
    _obs_matches_render(sim)
    assert not torch.equal(sim.terminal_obs[0], sim.obs[0])
    # --- masked explicit reset: only the masked environments move on to episode 3
    mask = torch.zeros(N, dtype=torch.uint8, device="cuda")
    mask[::2] = 1
    sim.reset(mask)
    q3 = sim.get_state()["qpos"]
    np.testing.assert_array_equal(q3[1::2], q2[1::2])
    xy, _ = expected_pose(q0, SEED, 4, 3, NXY, NYAW)
    np.testing.assert_allclose(q3[4, 7:9], xy, atol=2e-7)
    _obs_matches_render(sim)
    sim.close()
    # --- same seed, same stream; another seed, another stream
    s1 = GripperSim(cfg, num_envs=N)
    np.testing.assert_array_equal(s1.get_state()["qpos"], q)
    s1.close()
    s2 = GripperSim(make_config(sim_env="/xmls/sugar_cube_env.xml", reset_noise_xy=NXY, reset_noise_yaw=NYAW, seed=SEED + 1), num_envs=N)
    assert np.abs(s2.get_state()["qpos"][:, 7:9] - q[:, 7:9]).max() > 1e-3
    s2.close()
