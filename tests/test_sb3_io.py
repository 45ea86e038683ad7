"""CPU: the stable-baselines3 file formats around the hot path (sb3_io.py) — model archives with SB3's parameter names
(reference eval_agent.py:34-48), Monitor's monitor.csv (train_agent.py:22) and EvalCallback's evaluations.npz
(train_agent.py:47-58).  GPU: an archive loads into the tensor-core policy and reproduces the fp32 torch actor."""
import json
import os
import zipfile

import numpy as np
import pytest

from mujoco_rl_manipulate_unknown_objects_b200 import sb3_io
from mujoco_rl_manipulate_unknown_objects_b200.policy import init_params, param_spec


def test_archive_round_trip_keeps_sb3_names(tmp_path):
    params = init_params(channels=5, action_dim=6, n_flatten=1024, seed=3)
    critic = {"critic.qf0.0.weight": np.ones((256, 520), np.float32)}
    path = sb3_io.write_sb3_zip(str(tmp_path / "best_model"), params, extra_state=critic)
    assert path.endswith("best_model.zip")
    with zipfile.ZipFile(path) as z:
        assert {"data", "policy.pth", "_stable_baselines3_version"} <= set(z.namelist())
        assert json.loads(z.read("data"))["policy_kwargs"]["net_arch"] == [256, 256]  # eval_agent.py:42-46
    sd, data = sb3_io.read_sb3_zip(str(tmp_path / "best_model"))  # SB3 accepts the path without '.zip'
    assert "critic.qf0.0.weight" in sd and data["policy_class"]["__name__"] == "MultiInputPolicy"
    actor = sb3_io.actor_state_dict(sd)
    assert sorted(actor) == sorted(n for n, _ in param_spec())
    for name, shape in param_spec():
        assert tuple(actor[name].shape) == shape
        np.testing.assert_array_equal(actor[name].numpy(), params[name])


def test_reads_an_archive_laid_out_like_sb3_save_to_zip_file(tmp_path):
    """Independent of this repo's writer: an archive assembled the way stable-baselines3's `save_to_zip_file` lays it out for an
    off-policy SAC model with a custom features extractor — JSON `data` whose non-JSON entries are {":type:", ":serialized:"}
    (base64 cloudpickle), `policy.pth` holding actor, critic and critic_target with the features extractor repeated under each,
    separate optimizer state dicts, `pytorch_variables.pth`, the version file and `system_info.txt` (SB3 >= 1.5).  stable-baselines3
    itself is not installed here, so this pins the READER to the published layout, not to SB3's code."""
    import base64
    import collections
    import io
    import cloudpickle
    import torch
    params = init_params(channels=5, action_dim=6, n_flatten=1024, seed=9)
    sd = collections.OrderedDict()
    for k, v in params.items():
        sd["actor." + k] = torch.as_tensor(v)
    fe = {k: torch.as_tensor(v) for k, v in params.items() if k.startswith("features_extractor.")}
    for net in ("critic", "critic_target"):
        for k, v in fe.items():
            sd[net + "." + k] = v.clone()
        for q in ("qf0", "qf1"):
            for i, (o, n) in enumerate(((256, 514 + 6), (256, 256), (1, 256))):
                sd["%s.%s.%d.weight" % (net, q, 2 * i)] = torch.zeros(o, n)
                sd["%s.%s.%d.bias" % (net, q, 2 * i)] = torch.zeros(o)

    def ser(obj):
        return {":type:": str(type(obj)), ":serialized:": base64.b64encode(cloudpickle.dumps(obj)).decode()}
    data = {"policy_class": ser(dict), "verbose": 1, "policy_kwargs": ser({"features_extractor_class": dict, "net_arch": [256, 256]}),
            "observation_space": ser(("Dict", (5, 64, 64))), "action_space": ser(("Box", 6)), "n_envs": 1, "num_timesteps": 100000,
            "buffer_size": 10000, "batch_size": 256, "learning_starts": 100, "tau": 0.005, "gamma": 0.99, "gradient_steps": 1,
            "lr_schedule": ser(lambda _: 3e-4), "target_entropy": -6.0, "ent_coef": "auto", "_last_obs": None}

    def pth(obj):
        b = io.BytesIO()
        torch.save(obj, b)
        return b.getvalue()
    path = tmp_path / "best_model.zip"
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", json.dumps(data, indent=4))
        z.writestr("pytorch_variables.pth", pth({"log_ent_coef": torch.zeros(1)}))
        z.writestr("policy.pth", pth(sd))
        for opt in ("actor.optimizer", "critic.optimizer", "ent_coef_optimizer"):
            z.writestr(opt + ".pth", pth({"state": {}, "param_groups": [{"lr": 3e-4, "params": [0]}]}))
        z.writestr("_stable_baselines3_version", "1.6.2")
        z.writestr("system_info.txt", "OS: Linux\nPython: 3.9\nStable-Baselines3: 1.6.2\nPyTorch: 1.12\nGym: 0.21.0\n")
    got, meta = sb3_io.read_sb3_zip(str(tmp_path / "best_model"))
    assert meta["gamma"] == 0.99 and meta["policy_kwargs"][":type:"].startswith("<class")
    actor = sb3_io.actor_state_dict(got)
    assert sorted(actor) == sorted(n for n, _ in param_spec())
    for name, shape in param_spec():
        assert tuple(actor[name].shape) == shape
        np.testing.assert_array_equal(actor[name].numpy(), params[name])
    assert "critic_target.qf1.4.weight" in got and "critic.features_extractor.cnn.0.weight" in got


def test_archive_errors(tmp_path):
    p = tmp_path / "not_a_model.zip"
    with zipfile.ZipFile(p, "w") as z:
        z.writestr("readme", "x")
    with pytest.raises(ValueError):
        sb3_io.read_sb3_zip(str(p))
    with pytest.raises(KeyError):
        sb3_io.actor_state_dict({"critic.x": 1})


def test_monitor_csv_and_evaluations(tmp_path):
    w = sb3_io.MonitorWriter(str(tmp_path), t_start=100.0)
    infos = [{"episode": {"r": 1.25, "l": 7, "t": 3.5}}, {}, {"episode": {"r": -0.5, "l": 400, "t": 9.0}}]
    w.write_step(np.array([True, False, True]), infos)
    w.write_step(np.array([False, False, False]), [{}, {}, {}])
    w.close()
    lines = open(tmp_path / "monitor.csv").read().splitlines()
    assert lines[0].startswith("#") and json.loads(lines[0][1:]) == {"t_start": 100.0, "env_id": "RobotEnv-v0"}
    assert lines[1] == "r,l,t" and lines[2] == "1.25,7,3.5" and lines[3] == "-0.5,400,9.0" and len(lines) == 4
    header, rows = sb3_io.load_monitor(str(tmp_path / "monitor.csv"))
    assert header["t_start"] == 100.0 and rows.shape == (2, 3) and rows[1, 1] == 400
    log = sb3_io.EvalLog(str(tmp_path))
    log.add(1000, [1.0, 2.0, 3.0], [10, 11, 12])
    log.add(2000, [2.0, 2.0, 5.0], [9, 9, 9])
    z = np.load(tmp_path / "evaluations.npz")
    assert z["timesteps"].tolist() == [1000, 2000] and z["results"].shape == (2, 3) and z["ep_lengths"][1].tolist() == [9, 9, 9]


@pytest.mark.gpu
def test_zip_loads_into_the_tensor_core_policy(tmp_path):
    import torch
    from mujoco_rl_manipulate_unknown_objects_b200.policy import GripperPolicy
    from oracle import policy_ref
    params = init_params(channels=5, action_dim=6, n_flatten=1024, seed=11)
    path = sb3_io.write_sb3_zip(str(tmp_path / "best_model"), params)
    pol = GripperPolicy(max_envs=32, seed=0)  # different random weights until the archive is loaded
    pol.load_sb3_zip(path)
    for name, _ in param_spec():
        np.testing.assert_array_equal(pol.state_dict()[name], params[name])
    obs = torch.randint(0, 256, (32, 5, 64, 64), dtype=torch.uint8, generator=torch.Generator().manual_seed(0))
    act = pol(obs.cuda()).cpu().numpy()
    with torch.no_grad():
        ref = policy_ref.actor(params, obs.numpy())["action"].numpy()
    assert np.abs(act - ref).max() < 1e-2
    out = pol.save_sb3_zip(str(tmp_path / "resaved"))
    sd, _ = sb3_io.read_sb3_zip(out)
    np.testing.assert_array_equal(sd["actor.mu.weight"].numpy(), params["mu.weight"])
    pol.close()
