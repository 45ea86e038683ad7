"""stable-baselines3 file formats around the hot path, without importing stable-baselines3.

* `read_sb3_zip` / `write_sb3_zip` — the `best_model.zip` archives that `SAC.load` / `model.save` exchange (reference
  eval_agent.py:34-48, train_agent.py:60-66).  An SB3 archive is a zip holding `data` (JSON: class and constructor
  arguments), `policy.pth` (a `torch.save`d state dict of the whole policy: `actor.*`, `critic.*`, `critic_target.*`),
  `pytorch_variables.pth`, optimizer state dicts and `_stable_baselines3_version`.  Only `policy.pth` matters to the
  rollout-time actor; `GripperPolicy.load_state_dict` picks the `actor.` entries by their SB3 names.
* `MonitorWriter` — `Monitor`'s `monitor.csv` (`#{json header}` line, then `r,l,t` rows; train_agent.py:22) fed from the
  vectorised environment's `infos[i]["episode"]` records, so the reference's plotting scripts (scripts/sub_plot_*.py,
  which read these files through `load_results`) keep working.
* `EvalLog` — `EvalCallback`'s `evaluations.npz` (`timesteps`, `results`, `ep_lengths`; train_agent.py:47-58).
"""
import io
import json
import os
import time
import zipfile

import numpy as np


# ------------------------------------------------------------------------------------------------ model archives
def read_sb3_zip(path, map_location="cpu"):
    """Returns (state_dict of the policy, data dict).  `path` may omit the '.zip' suffix, as SB3 allows."""
    import torch
    if not os.path.exists(path) and os.path.exists(path + ".zip"):
        path = path + ".zip"
    with zipfile.ZipFile(path) as z:
        names = set(z.namelist())
        if "policy.pth" not in names:
            raise ValueError("%s is not a stable-baselines3 archive: no policy.pth (members: %s)" % (path, sorted(names)))
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location=map_location, weights_only=True)
        data = {}
        if "data" in names:
            try:
                data = json.loads(z.read("data").decode())
            except (ValueError, UnicodeDecodeError):
                data = {}
    return sd, data


def actor_state_dict(sd):
    """The actor's entries of an SB3 SAC policy state dict, with the 'actor.' prefix stripped (names as in
    policy.param_spec: features_extractor.cnn.N.*, features_extractor.linear.0.*, latent_pi.N.*, mu.*, log_std.*)."""
    out = {k[len("actor."):]: v for k, v in sd.items() if k.startswith("actor.")}
    if not out:
        raise KeyError("no 'actor.*' parameters in the state dict (is this a SAC MultiInputPolicy archive?)")
    return out


def write_sb3_zip(path, actor_params, data=None, extra_state=None):
    """Writes an archive that `read_sb3_zip` (and SB3's own loader, given a matching `data`) reads back: `actor_params`
    maps actor parameter names (no prefix) to arrays; `extra_state` may carry critic tensors under their full names."""
    import torch
    sd = {"actor." + k: torch.as_tensor(np.asarray(v)) for k, v in actor_params.items()}
    for k, v in (extra_state or {}).items():
        sd[k] = torch.as_tensor(np.asarray(v))
    buf = io.BytesIO()
    torch.save(sd, buf)
    if not path.endswith(".zip"):
        path = path + ".zip"
    meta = {"policy_class": {":type:": "<class 'abc.ABCMeta'>", "__module__": "stable_baselines3.sac.policies", "__name__": "MultiInputPolicy"},
            "policy_kwargs": {"net_arch": [256, 256], "share_features_extractor": True, "features_extractor_class": "AugmentedNatureCNN"}}
    meta.update(data or {})
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("data", json.dumps(meta, indent=2))
        z.writestr("policy.pth", buf.getvalue())
        z.writestr("_stable_baselines3_version", "1.8.0")
        z.writestr("system_info.txt", "written by mujoco_rl_manipulate_unknown_objects_b200.sb3_io\n")
    return path


# ------------------------------------------------------------------------------------------------ Monitor / EvalCallback logs
class MonitorWriter:
    """monitor.csv of stable-baselines3's Monitor for a whole vectorised environment: one row per finished episode."""

    EXT = "monitor.csv"

    def __init__(self, filename, env_id="RobotEnv-v0", t_start=None):
        if os.path.isdir(filename):
            filename = os.path.join(filename, self.EXT)
        elif not filename.endswith(self.EXT):
            filename = filename + "." + self.EXT
        self.filename = filename
        self.t_start = time.time() if t_start is None else float(t_start)
        os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
        self._f = open(filename, "wt")
        self._f.write("#%s\n" % json.dumps({"t_start": self.t_start, "env_id": env_id}))
        self._f.write("r,l,t\n")
        self._f.flush()
        self.episodes = 0

    def write_step(self, dones, infos):
        """`dones` bool [N] and `infos` (LazyInfos or list of dicts) of one VecEnv.step: appends the finished episodes."""
        idx = np.flatnonzero(np.asarray(dones))
        for i in idx:
            ep = infos[int(i)].get("episode")
            if ep is not None:
                self._f.write("%s,%d,%s\n" % (repr(round(float(ep["r"]), 6)), int(ep["l"]), repr(round(float(ep["t"]), 6))))
                self.episodes += 1
        if len(idx):
            self._f.flush()

    def close(self):
        if self._f:
            self._f.close()
            self._f = None


def load_monitor(filename):
    """Reads a monitor.csv back: (header dict, array of (r, l, t) rows) — what SB3's `load_results` extracts."""
    with open(filename) as f:
        first = f.readline()
        assert first.startswith("#"), "monitor.csv must start with a '#{json}' header line"
        header = json.loads(first[1:])
        cols = f.readline().strip().split(",")
        assert cols == ["r", "l", "t"], cols
        rows = [tuple(float(x) for x in ln.strip().split(",")) for ln in f if ln.strip()]
    return header, np.array(rows, dtype=np.float64).reshape(-1, 3)


class EvalLog:
    """evaluations.npz of stable-baselines3's EvalCallback: cumulative `timesteps` [E], `results` [E][episodes],
    `ep_lengths` [E][episodes]; rewritten after every evaluation like the callback does."""

    def __init__(self, log_path):
        self.path = log_path if log_path.endswith(".npz") else os.path.join(log_path, "evaluations.npz")
        self.timesteps, self.results, self.ep_lengths = [], [], []

    def add(self, num_timesteps, episode_rewards, episode_lengths):
        self.timesteps.append(int(num_timesteps))
        self.results.append([float(x) for x in episode_rewards])
        self.ep_lengths.append([int(x) for x in episode_lengths])
        os.makedirs(os.path.dirname(os.path.abspath(self.path)), exist_ok=True)
        np.savez(self.path, timesteps=np.array(self.timesteps), results=np.array(self.results), ep_lengths=np.array(self.ep_lengths))
        return float(np.mean(episode_rewards))
