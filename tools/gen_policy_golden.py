#!/usr/bin/env python
"""Generate tests/golden/policy_vectors.npz by running the reference's UNMODIFIED `AugmentedNatureCNN`
(/root/reference/models/feature_extractor.py) in this container.

gym and stable-baselines3 are not installable here, so the two names the file imports are stubbed: `gym` (only used in a type
annotation) and `BaseFeaturesExtractor` (SB3's 10-line nn.Module base that stores `features_dim`).  The class body — layer
stack, channel split, concat of the two direct features — is the reference's own code.  Parameters come from
`init_params(seed)` (numpy, deterministic) and are loaded into the reference module, so the fixture only stores the seed, the
observations and the outputs.  The SB3 actor head (third-party) is restated in oracle/policy_ref.py.

Run:  python tools/gen_policy_golden.py      (needs /root/reference; tests only read the .npz)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mujoco_rl_manipulate_unknown_objects_b200.policy import init_params  # noqa: E402  (numpy only)
from oracle import policy_ref  # noqa: E402

REF = "/root/reference/models/feature_extractor.py"


def load_reference_class():
    gym = types.ModuleType("gym")
    gym.spaces = types.SimpleNamespace(Dict=dict)
    sb3 = types.ModuleType("stable_baselines3")
    common = types.ModuleType("stable_baselines3.common")
    layers = types.ModuleType("stable_baselines3.common.torch_layers")

    class BaseFeaturesExtractor(torch.nn.Module):  # SB3 torch_layers.py: stores the space and features_dim, nothing else
        def __init__(self, observation_space, features_dim=0):
            super().__init__()
            self._observation_space, self._features_dim = observation_space, features_dim

    layers.BaseFeaturesExtractor = BaseFeaturesExtractor
    sys.modules.update({"gym": gym, "stable_baselines3": sb3, "stable_baselines3.common": common,
                        "stable_baselines3.common.torch_layers": layers})
    spec = importlib.util.spec_from_file_location("ref_feature_extractor", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.AugmentedNatureCNN


class _Box:
    def __init__(self, shape):
        self.shape = shape

    def sample(self):
        return np.random.randint(0, 256, self.shape).astype(np.uint8)


def main():
    Ref = load_reference_class()
    out = {}
    for tag, C in (("full", 5), ("nodepth", 4)):
        seed = 7 if C == 5 else 8
        params = init_params(channels=C, action_dim=6, n_flatten=1024, seed=seed)
        net = Ref({"observation": _Box((C, 64, 64))})
        sd = {k.replace("features_extractor.", ""): torch.as_tensor(v) for k, v in params.items() if k.startswith("features_extractor.")}
        net.load_state_dict(sd)
        rng = np.random.default_rng(100 + C)
        n = 6
        obs = rng.integers(0, 256, (n, C, 64, 64), dtype=np.uint8)
        obs[:, -1] = 0
        obs[:, -1, 0, 0] = rng.integers(0, 4, n)
        obs[:, -1, 0, 1] = rng.integers(0, 4, n)
        obs[0, :-1] = 0      # all-black image
        obs[1, :-1] = 255    # saturated image
        with torch.no_grad():
            feat = net({"observation": torch.as_tensor(obs).float() / 255.0}).numpy()  # SB3 preprocess_obs: /255
            mine = policy_ref.actor(params, obs)
        assert np.abs(mine["features"].numpy() - feat).max() < 1e-5, "oracle restatement disagrees with the reference class"
        noise = rng.standard_normal((n, 6)).astype(np.float32)
        with torch.no_grad():
            sto = policy_ref.actor(params, obs, noise)
        out.update({tag + "_seed": seed, tag + "_obs": obs, tag + "_features": feat, tag + "_mu": mine["mu"].numpy(),
                    tag + "_log_std": mine["log_std"].numpy(), tag + "_action": mine["action"].numpy(), tag + "_noise": noise,
                    tag + "_action_stochastic": sto["action"].numpy()})
    path = os.path.join(ROOT, "tests", "golden", "policy_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
