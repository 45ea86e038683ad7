"""Multi-GPU host logic on CPU: world_size-2 `gloo` process groups exercise the environment sharding, the rollout-statistics
reduction and the flat gradient bucket.  The GPU test runs the device-resident rollout + PPO update once."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_envs, q):
    import torch
    import torch.distributed as dist
    from mujoco_rl_manipulate_unknown_objects_b200._native import INFO
    from mujoco_rl_manipulate_unknown_objects_b200.rollout import RolloutStats, allreduce_flat_, shard_range, rank_seed, dist_info
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert dist_info()[:2] == (rank, world)
    lo, hi = shard_range(total_envs, rank, world)
    # synthetic per-environment info rows whose content is a function of the GLOBAL environment index
    idx = torch.arange(lo, hi, dtype=torch.float32)
    info = torch.zeros((hi - lo, INFO["STRIDE"]))
    info[:, INFO["NSUB_A"]] = idx
    info[:, INFO["NSUB_C"]] = 1
    info[:, INFO["EPISODE_RETURN"]] = 2 * idx
    info[:, INFO["EPISODE_STEP"]] = 7
    info[:, INFO["STATUS"]] = (idx % 3 == 0).float()
    info[:, INFO["GRASP"]] = 3 * (idx % 2 == 0).float()
    reward = idx / 10
    done = (idx % 4 == 0).to(torch.uint8)
    st = RolloutStats("cpu")
    st.update(info, reward, done)
    st.update(info, reward, done)
    st.all_reduce()
    g = torch.Generator().manual_seed(rank_seed(5, rank))
    grads = [torch.full((3, 2), float(rank + 1)), torch.full((5,), 10.0 * (rank + 1))]
    allreduce_flat_(grads, average=True)
    q.put((rank, (lo, hi), st.as_dict(), [x.numpy().copy() for x in grads], float(torch.rand(1, generator=g))))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_sharding_stats_and_gradient_bucket():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, total, port = 2, 101, _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, (a0, b0), s0, g0, u0), (r1, (a1, b1), s1, g1, u1) = out
    assert (a0, b0, a1, b1) == (0, 51, 51, 101)  # contiguous, balanced, complete
    idx = np.arange(total, dtype=np.float64)
    done = idx % 4 == 0
    for s in (s0, s1):  # both ranks hold the global sums
        assert s["transitions"] == 2 * total
        assert s["substeps"] == pytest.approx(2 * (idx.sum() + total))
        assert s["reward_sum"] == pytest.approx(2 * (idx / 10).sum(), rel=1e-6)
        assert s["episodes"] == 2 * done.sum()
        assert s["episode_return_sum"] == pytest.approx(2 * (2 * idx[done]).sum())
        assert s["mean_episode_length"] == pytest.approx(7.0)
        assert s["fails"] == 2 * np.sum(done & (idx % 3 == 0))
        assert s["grasps"] == 2 * np.sum(idx % 2 == 0)
    for g in (g0, g1):
        assert np.allclose(g[0], 1.5) and np.allclose(g[1], 15.0)
    assert u0 != u1  # per-rank seeds differ


def _bucket_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from torch import nn
    from mujoco_rl_manipulate_unknown_objects_b200.rollout import FlatAdam
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(5, 7), nn.Tanh(), nn.Linear(7, 2))
    opt = FlatAdam(net, lr=1e-3)
    ptrs = [p.grad.data_ptr() for p in net.parameters()]
    x = torch.full((4, 5), float(rank + 1))
    opt.zero_grad()
    net(x).sum().backward()
    local = opt.grad.clone()
    assert [p.grad.data_ptr() for p in net.parameters()] == ptrs       # autograd accumulated into the bucket views
    w = opt.all_reduce()                                               # ONE collective on the flat bucket, no cat / scatter
    q.put((rank, w, local.numpy().copy(), opt.grad.numpy().copy(), [p.grad.detach().numpy().copy() for p in net.parameters()]))
    try:
        opt.step(w)
        q.put((rank, "stepped"))
    except Exception as e:  # noqa: BLE001 - the fused Adam kernel is CUDA only: no CPU fallback
        q.put((rank, "raised"))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_flat_gradient_bucket_of_the_fused_optimiser():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2 * world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    first = sorted(g for g in got if len(g) == 5)
    (r0, w0, l0, s0, g0), (r1, w1, l1, s1, g1) = first
    assert w0 == w1 == 2
    np.testing.assert_allclose(s0, l0 + l1, rtol=1e-6)
    np.testing.assert_allclose(s1, s0, rtol=0, atol=0)
    assert np.abs(l0 - l1).max() > 0
    o = 0
    for g in g0:   # the parameters' .grad are views of the reduced bucket (16-byte aligned slots)
        np.testing.assert_array_equal(g.reshape(-1), s0[o:o + g.size])
        o += (g.size + 3) // 4 * 4
    assert sorted(g[1] for g in got if len(g) == 2) == ["raised", "raised"]


def test_shard_range_properties():
    from mujoco_rl_manipulate_unknown_objects_b200.rollout import shard_range
    for total in (0, 1, 7, 4096, 4097, 32768):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


@pytest.mark.gpu
def test_device_resident_rollout_and_ppo_update():
    import torch
    from mujoco_rl_manipulate_unknown_objects_b200 import make_config
    from mujoco_rl_manipulate_unknown_objects_b200.rollout import RolloutWorker, PPOLearner
    w = RolloutWorker(make_config(sim_env="/xmls/sugar_cube_env.xml", time_horizon=6), envs_per_gpu=64, seed=3)
    learner = PPOLearner(w, minibatch=128, seed=1)
    # the torch twin and the tensor-core forward agree on the same weights
    mu_t, ls_t, _ = learner.net(w.sim.obs)
    w.policy.forward(w.sim.obs)
    assert (w.policy.mu[:64] - mu_t).abs().max().item() < 2e-2
    T = 8
    storage = w.make_storage(T)
    stats = w.collect(T, storage=storage).all_reduce().as_dict()
    assert stats["transitions"] == 64 * T and stats["substeps"] > 64 * T
    assert stats["episodes"] >= 64  # time_horizon 6: every environment finishes at least one episode in 8 steps
    assert storage["obs"].float().std() > 0 and torch.isfinite(storage["actions"]).all()
    before = {k: v.clone() for k, v in learner.net.state_dict().items()}
    loss = learner.update(storage, w.sim.obs)
    assert np.isfinite(loss)
    assert any((before[k] - v).abs().max() > 0 for k, v in learner.net.state_dict().items())
    # the kernels now run the updated weights
    mu_t, _, _ = learner.net(w.sim.obs)
    w.policy.forward(w.sim.obs)
    assert (w.policy.mu[:64] - mu_t).abs().max().item() < 2e-2
    w.close()
