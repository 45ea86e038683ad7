"""GPU: BatchedRobotVecEnv through the stable-baselines3 VecEnv protocol — the upper boundary of the hot path
(SURVEY.md §8b): reference RobotEnv under DummyVecEnv + Monitor (train_agent.py:17-23)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# robot_env.py:226-241 — the info keys of RobotEnv.step
INFO_KEYS = {"init_obj_pos", "final_obj_pos", "target_dir", "gripper_open", "controls", "object_grasped", "episode_step", "status",
             "gripper_position", "object_position", "position_reached", "total_distance", "line_distance"}


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
        assert rew.shape == (N,) and rew.dtype == np.float32 and dones.shape == (N,) and dones.dtype == bool and len(infos) == N
        inf = infos[3]
        assert INFO_KEYS <= set(inf) and inf["status"] in ("RUNNING", "FAIL", "TIME_LIMIT")
        assert set(inf["position_reached"]) == {"target", "initial", "fail"}
        # compute_reward (robot_env.py:243-273) on the stored transition reproduces the step's reward
        assert abs(env.compute_reward(obs["achieved_goal"][3], obs["desired_goal"][3], inf) - rew[3]) < 1e-4
        if t < H - 1:
            assert not dones.any() and "terminal_observation" not in inf
    assert dones.all()  # time limit: every episode ends on step H, the VecEnv has already reset
    for i in (0, N - 1):
        inf = infos[i]
        assert inf["TimeLimit.truncated"] is True and inf["status"] == "TIME_LIMIT" and inf["episode"]["l"] == H
        assert abs(inf["episode"]["r"] - returns[i]) < 1e-4
        assert inf["terminal_observation"]["observation"].shape == (5, 64, 64)
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
