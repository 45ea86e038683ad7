"""ORACLE (test infrastructure, never on the product path): fp32 torch restatement of the rollout policy forward.

Follows
  * SB3 `preprocess_obs` — uint8 image observations are divided by 255 (all C channels, so the two direct features arrive
    as v/255),
  * reference models/feature_extractor.py:19-49 `AugmentedNatureCNN`: NatureCNN on the first C-1 channels -> 512, concat
    observation[:, -1, 0, :2] -> 514,
  * SB3 sac/policies.py `Actor`: latent_pi = MLP(514 -> 256 -> 256, ReLU) (net_arch from reference train_agent.py:18-20),
    mu = Linear(256, A), log_std = Linear(256, A) clamped to [-20, 2], action = tanh(mu [+ exp(log_std) * noise]).
Pinned by tests/golden/policy_vectors.npz: the feature part against the reference's own AugmentedNatureCNN class executed in
the build container (tools/gen_policy_golden.py); the SB3 actor head is absent from /root/reference (third-party,
un-vendored) and is restated from its published definition.
"""
import numpy as np
import torch
import torch.nn.functional as F

LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0


def _t(params, name):
    v = params[name]
    return v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v), dtype=torch.float32)


def features(params, obs_u8):
    """obs uint8 [n, C, H, W] (tensor or array) -> float32 [n, 514]."""
    obs = torch.as_tensor(np.asarray(obs_u8)) if not isinstance(obs_u8, torch.Tensor) else obs_u8
    x = obs.to(torch.float32) / 255.0
    img, pad = x[:, :-1], x[:, -1]
    p = "features_extractor."
    h = F.relu(F.conv2d(img, _t(params, p + "cnn.0.weight"), _t(params, p + "cnn.0.bias"), stride=4))
    h = F.relu(F.conv2d(h, _t(params, p + "cnn.2.weight"), _t(params, p + "cnn.2.bias"), stride=2))
    h = F.relu(F.conv2d(h, _t(params, p + "cnn.4.weight"), _t(params, p + "cnn.4.bias"), stride=1))
    h = h.flatten(1)
    h = F.relu(F.linear(h, _t(params, p + "linear.0.weight"), _t(params, p + "linear.0.bias")))
    return torch.cat((h, pad[:, 0, :2]), dim=1)


def actor(params, obs_u8, noise=None):
    """-> dict(features, mu, log_std, action) float32 tensors."""
    f = features(params, obs_u8)
    h = F.relu(F.linear(f, _t(params, "latent_pi.0.weight"), _t(params, "latent_pi.0.bias")))
    h = F.relu(F.linear(h, _t(params, "latent_pi.2.weight"), _t(params, "latent_pi.2.bias")))
    mu = F.linear(h, _t(params, "mu.weight"), _t(params, "mu.bias"))
    log_std = torch.clamp(F.linear(h, _t(params, "log_std.weight"), _t(params, "log_std.bias")), LOG_STD_MIN, LOG_STD_MAX)
    pre = mu if noise is None else mu + torch.exp(log_std) * torch.as_tensor(np.asarray(noise), dtype=torch.float32)
    return dict(features=f, mu=mu, log_std=log_std, action=torch.tanh(pre))
