"""GPU: BatchedRobotVecEnv through the stable-baselines3 VecEnv protocol — the upper boundary of the hot path
(SURVEY.md §8b): reference RobotEnv under DummyVecEnv + Monitor (train_agent.py:17-23)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# robot_env.py:226-241 — the sixteen info keys of RobotEnv.step, in the reference's order
INFO_KEYS = ["old_obs", "new_obs", "init_obj_pos", "final_obj_pos", "target_dir", "gripper_open", "controls", "object_grasped", "episode_step",
             "episode_rewards", "status", "gripper_position", "object_position", "position_reached", "total_distance", "line_distance"]


def test_vec_env_protocol(tmp_path):
    from mujoco_rl_manipulate_unknown_objects_b200 import BatchedRobotVecEnv, make_config
    from mujoco_rl_manipulate_unknown_objects_b200.sb3_io import load_monitor
    N, H = 32, 4
    env = BatchedRobotVecEnv(make_config(sim_env="/xmls/sugar_cube_env.xml", time_horizon=H), num_envs=N, monitor_file=str(tmp_path))
    assert env.num_envs == N and env.action_space.shape == (6,) and env.observation_space["observation"].shape == (5, 64, 64)
    obs = env.reset()
    assert obs["observation"].shape == (N, 5, 64, 64) and obs["observation"].dtype == np.uint8
    assert obs["achieved_goal"].shape == (N, 2) and obs["achieved_goal"].dtype == np.float32
    np.testing.assert_array_equal(obs["desired_goal"], np.tile(np.float32([1, 0]), (N, 1)))  # robot_env.py:72
    assert (obs["observation"] == obs["observation"][0]).all()  # no reset randomisation in the reference
    rng = np.random.default_rng(0)
    returns = np.zeros(N)
    step_rewards = np.zeros((H, N))
    assert env.Status.TIME_LIMIT.value == 2 and [m.name for m in env.Status] == ["RUNNING", "FAIL", "TIME_LIMIT"]  # robot_env.py:19-22
    first_obs = obs["observation"]
    first_copy = first_obs.copy()
    with pytest.raises(ValueError):
        env.step_async(np.zeros((N, 5), np.float32))
    with pytest.raises(RuntimeError):
        env.step_wait()
    for t in range(H):
        a = rng.uniform(-1, 1, (N, 6)).astype(np.float32)
        a[:, 0] = np.abs(a[:, 0])
        env.step_async(a)
        obs, rew, dones, infos = env.step_wait()
        returns += rew
        step_rewards[t] = rew
        assert rew.shape == (N,) and rew.dtype == np.float32 and dones.shape == (N,) and dones.dtype == bool and len(infos) == N
        inf = infos[3]
        assert list(inf)[:16] == INFO_KEYS and isinstance(inf["status"], env.Status) and inf["status"].name in ("RUNNING", "FAIL", "TIME_LIMIT")
        if t == 0:
            # what reset() returned stays intact while the next step runs (two pinned buffer sets alternate) and is this step's old_obs
            np.testing.assert_array_equal(first_obs, first_copy)
            np.testing.assert_array_equal(inf["old_obs"], first_copy[3])
        if t < H - 1:
            np.testing.assert_array_equal(inf["new_obs"], obs["observation"][3])  # robot_env.py:186,227
            assert inf["episode_rewards"].shape == (H,) and inf["episode_step"] == t + 1
            np.testing.assert_allclose(inf["episode_rewards"][:t + 1], step_rewards[:t + 1, 3], atol=1e-6)  # robot_env.py:199
            assert (inf["episode_rewards"][t + 1:] == 0).all()
        assert set(inf["position_reached"]) == {"target", "initial", "fail"}
        # compute_reward (robot_env.py:243-273) on the stored transition reproduces the step's reward
        assert abs(env.compute_reward(obs["achieved_goal"][3], obs["desired_goal"][3], inf) - rew[3]) < 1e-4
        if t < H - 1:
            assert not dones.any() and "terminal_observation" not in inf
    assert dones.all()  # time limit: every episode ends on step H, the VecEnv has already reset
    for i in (0, N - 1):
        inf = infos[i]
        assert "TimeLimit.truncated" not in inf  # the reference never wraps RobotEnv in TimeLimit (train_agent.py:17-23)
        assert inf["status"] is env.Status.TIME_LIMIT and inf["status"].value == 2 and inf["episode"]["l"] == H
        assert abs(inf["episode"]["r"] - returns[i]) < 1e-4
        assert inf["terminal_observation"]["observation"].shape == (5, 64, 64)
        np.testing.assert_array_equal(inf["new_obs"], inf["terminal_observation"]["observation"])  # the finished episode's last image
        assert inf["new_obs"].std() > 0 and not np.array_equal(inf["new_obs"], obs["observation"][i])
        np.testing.assert_allclose(inf["episode_rewards"], step_rewards[:, i], atol=1e-6)
    assert (obs["observation"] == obs["observation"][0]).all()  # the returned observation is the next episode's first
    assert env.get_attr("episode_step") == [0] * N and env.get_attr("gripper_open", indices=[1]) == [True]
    assert env.env_is_wrapped(object) == [False] * N and env.seed(3) == [3] * N
    img = env.render(mode="rgb_array", camera_id=3, width=96, height=64)  # robot_env.py:330-334: cameras 0-2 side by side
    assert img.shape == (64, 3 * 96, 3) and img.dtype == np.uint8 and img.std() > 0
    dep = env.render(mode="depth_array", camera_id=1, width=32, height=32)
    assert dep.shape == (32, 32)
    env.close()
    header, rows = load_monitor(str(tmp_path / "monitor.csv"))
    assert rows.shape == (N, 3) and (rows[:, 1] == H).all()
    np.testing.assert_allclose(np.sort(rows[:, 0]), np.sort(returns), atol=1e-4)


def test_truncation_key_and_private_copies():
    from mujoco_rl_manipulate_unknown_objects_b200 import BatchedRobotVecEnv, make_config
    N, H = 8, 2
    env = BatchedRobotVecEnv(make_config(sim_env="/xmls/sand_ball_env.xml", time_horizon=H), num_envs=N, truncation_as_timeout=True, copy_outputs=True)
    o0 = env.reset()
    a = np.zeros((N, 6), np.float32)
    o1, r1, d1, i1 = env.step(a)
    o2, r2, d2, i2 = env.step(a)
    assert d2.all() and i2[0]["TimeLimit.truncated"] is True and "TimeLimit.truncated" not in i1[0]
    keep = o1["observation"].copy()
    for _ in range(3):
        env.step(a)
    np.testing.assert_array_equal(o1["observation"], keep)  # copy_outputs=True: private arrays, never overwritten
    env.close()


def test_intrinsic_reward_reaches_the_episode_return_and_compute_reward():
    """reward.py:45-55 adds the KL term to the step reward; robot_env.py:199 sums step rewards into the episode (Monitor's 'r');
    compute_reward on the stored transition (old_obs / new_obs in info, robot_env.py:250-251) reproduces it."""
    from mujoco_rl_manipulate_unknown_objects_b200 import BatchedRobotVecEnv, make_config
    from mujoco_rl_manipulate_unknown_objects_b200._native import INFO
    N, H = 16, 3
    env = BatchedRobotVecEnv(make_config(sim_env="/xmls/sugar_cube_env.xml", time_horizon=H, im_reward=True), num_envs=N)
    env.reset()
    rng = np.random.default_rng(1)
    total = np.zeros(N)
    saw_kl = False
    for t in range(H):
        a = rng.uniform(-1, 1, (N, 6)).astype(np.float32)
        a[:, 0] = np.abs(a[:, 0])
        obs, rew, dones, infos = env.step(a)
        total += rew
        for i in (0, 5):
            inf = infos[i]
            r = env.compute_reward(inf["achieved_goal"], inf["desired_goal"], inf)
            assert abs(r - rew[i]) < 2e-4, (t, i, r, rew[i])
            assert abs(env.last_info_rows[i, INFO["REWARD"]] - rew[i]) < 1e-6  # the info record carries the full reward too
        saw_kl = saw_kl or bool((rew > 0.01).any())
    assert dones.all() and saw_kl
    for i in range(N):
        assert abs(infos[i]["episode"]["r"] - total[i]) < 1e-4, (i, infos[i]["episode"]["r"], total[i])
        np.testing.assert_allclose(infos[i]["episode_rewards"].sum(), total[i], atol=1e-4)
    env.close()
