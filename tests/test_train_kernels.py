"""GPU: the on-device PPO post-processing kernels (csrc/train_kernels.cu) against plain torch:
grl_gae (GAE + returns + advantage moments over the [T][N] rollout storage) and grl_adam_step (Adam over one flat bucket
that doubles as the gradient all-reduce buffer) through FlatAdam.  fp32 arithmetic on both sides; tolerances are a few ulps of
the quantities involved."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_gae_kernel_matches_the_recurrence():
    import torch
    from mujoco_rl_manipulate_unknown_objects_b200.rollout import gae_device
    g = torch.Generator(device="cuda").manual_seed(0)
    for T, N in ((1, 1), (16, 4096), (33, 1000), (128, 37)):
        rew = torch.rand((T, N), device="cuda", generator=g) * 3
        done = (torch.rand((T, N), device="cuda", generator=g) < 0.1).to(torch.uint8)
        val = torch.randn((T + 1, N), device="cuda", generator=g)
        gamma, lam = 0.99, 0.95
        adv, ret, mom = gae_device(rew, done, val, gamma, lam)
        ref = torch.zeros((T, N), device="cuda", dtype=torch.float64)
        last = torch.zeros(N, device="cuda", dtype=torch.float64)
        r64, v64 = rew.double(), val.double()
        for k in reversed(range(T)):
            nd = 1.0 - done[k].double()
            delta = r64[k] + gamma * v64[k + 1] * nd - v64[k]
            last = delta + gamma * lam * nd * last
            ref[k] = last
        scale = max(1.0, float(ref.abs().max()))
        assert float((adv.double() - ref).abs().max()) <= 2e-5 * scale, (T, N)
        assert float((ret.double() - (ref + v64[:T])).abs().max()) <= 2e-5 * scale
        assert abs(float(mom[0]) - float(adv.double().sum())) <= 1e-6 * max(1.0, float(adv.double().abs().sum()))
        assert abs(float(mom[1]) - float((adv.double() ** 2).sum())) <= 1e-6 * max(1.0, float((adv.double() ** 2).sum()))
    # an episode boundary cuts the recurrence: with done everywhere the advantage is the one-step TD error
    rew = torch.ones((4, 8), device="cuda"); done = torch.ones((4, 8), device="cuda", dtype=torch.uint8); val = torch.full((5, 8), 0.5, device="cuda")
    adv, ret, _ = gae_device(rew, done, val, 0.99, 0.95)
    assert torch.allclose(adv, torch.full_like(adv, 0.5)) and torch.allclose(ret, torch.ones_like(ret))


def test_flat_adam_matches_torch_adam_and_keeps_gradients_in_the_bucket():
    import torch
    from torch import nn
    from mujoco_rl_manipulate_unknown_objects_b200.rollout import FlatAdam
    torch.manual_seed(0)
    def make():
        torch.manual_seed(1)
        return nn.Sequential(nn.Linear(37, 64), nn.ReLU(), nn.Linear(64, 5, bias=True), nn.Tanh(), nn.Linear(5, 3)).cuda()
    a, b = make(), make()
    ref = torch.optim.Adam(a.parameters(), lr=3e-3, betas=(0.9, 0.999), eps=1e-8)
    opt = FlatAdam(b, lr=3e-3, betas=(0.9, 0.999), eps=1e-8)
    assert opt.count % 4 == 0 and all(p.data_ptr() % 16 == 0 for p in b.parameters())
    ptrs = [p.grad.data_ptr() for p in b.parameters()]
    x = torch.randn(256, 37, device="cuda")
    y = torch.randn(256, 3, device="cuda")
    for it in range(25):
        ref.zero_grad(set_to_none=True)
        ((a(x) - y) ** 2).mean().backward()
        ref.step()
        opt.zero_grad()
        ((b(x) - y) ** 2).mean().backward()
        assert [p.grad.data_ptr() for p in b.parameters()] == ptrs, "autograd replaced a gradient view"
        world = opt.all_reduce()   # no process group: a no-op that reports world size 1
        opt.step(world)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert float((pa - pb).abs().max()) <= 2e-6 * max(1.0, float(pa.abs().max())), float((pa - pb).abs().max())
    # the module computes with the bucket: writing the flat array changes the parameters
    opt.flat.zero_()
    assert all(float(p.abs().max()) == 0.0 for p in b.parameters())
    # grad_scale (1 / world after a SUM all-reduce): two steps with doubled gradients and world=2 equal one plain step
    c, d = make(), make()
    o1, o2 = FlatAdam(c, lr=1e-2), FlatAdam(d, lr=1e-2)
    o1.zero_grad(); ((c(x) - y) ** 2).mean().backward(); o1.step(1)
    o2.zero_grad(); (2 * ((d(x) - y) ** 2).mean()).backward(); o2.step(2)
    assert float((o1.flat - o2.flat).abs().max()) <= 1e-6
