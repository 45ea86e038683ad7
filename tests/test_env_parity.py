"""GPU: RobotEnv.step / reset through the C-ABI against the oracle's environment layer at matched pre-step states,
and against the golden fixtures (reference python) where the trajectory is contact-free.

Tolerances (BASELINE.json north_star): reward and goals within 1e-4; pad channel scalars, done, status and the
per-phase substep counts exact at matched state.  A grasping step is a chaotic map (see test_oracle_golden.py), so
steps in which the fp32 state ends > 1e-4 away from the fp64 one are counted as 'flipped', bounded, and excluded
from the exact comparisons."""
import numpy as np
import pytest

from helpers import golden, oracle_model
from mujoco_rl_manipulate_unknown_objects_b200._native import INFO as I

pytestmark = pytest.mark.gpu

CASES = [("sugar_cube", 0, {}, 3), ("sugar_cube", 45, {}, 4), ("sand_ball", 0, {}, 5), ("bread_crumb", 0, {}, 6), ("acorn", 0, {}, 7),
         ("sand_ball", 45, dict(include_roll=False), 8), ("sugar_cube", 0, dict(her_buffer=True, time_horizon=12), 9),
         ("gripper_two_fingers", 0, {}, 10), ("sugar_cube", 120, {}, 11)]  # 120 degrees: the `_get_direction` unit vector (robot_env.py:46-54)  # the primitive-box scene (mjc_PlaneBox + box support in MPR)


# MEASURED (round 2, B200, 48 matched agent steps per case): steps in which the fp32 trajectory leaves the fp64 one by > 1e-4 or takes a
# different number of substeps ("flipped by contact chaos": a grasping step is a discontinuous map).  WHICH steps flip depends on
# the last bit of the fp32 arithmetic: the two builds measured this round (before / after the solver helpers were inlined, which
# changes FMA contraction) gave per-case counts (10 6 2 1 1 5 7 5 4) and (11 5 4 4 0 6 6 9 6), totals 41 and 51 of 432.  Asserted:
# per case the larger of the two measurements + 50 % + 2 (small-number noise), and over all cases a total of at most 60 (14 % of the
# steps; 1.5 sigma above the larger total).
MEASURED_FLIPS = {"sugar_cube-0-default": 11, "sugar_cube-45-default": 6, "sand_ball-0-default": 4, "bread_crumb-0-default": 4, "acorn-0-default": 1,
                  "sand_ball-45-include_roll": 6, "sugar_cube-0-her_buffer-time_horizon": 7, "gripper_two_fingers-0-default": 9, "sugar_cube-120-default": 6}
MAX_TOTAL_FLIPS = 60
_FLIPS_SEEN = {}


def make(scene, direction, kw, n, auto_reset=False):
    from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
    cfg = make_config(sim_env="/xmls/%s.xml" % (scene if scene == "gripper_two_fingers" else scene + "_env"), direction=direction, **kw)
    return GripperSim(cfg, num_envs=n, auto_reset=auto_reset)


def tape(seed, n):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, (n, 6))
    a[:, 0] = np.abs(a[:, 0])
    a[:, 5] = np.where((np.arange(n) // 3) % 2 == 0, -np.abs(a[:, 5]), np.abs(a[:, 5]))
    return a.astype(np.float32)


@pytest.mark.parametrize("scene,direction,kw,seed", CASES, ids=["%s-%d-%s" % (c[0], c[1], "-".join(c[2]) or "default") for c in CASES])
def test_agent_step_at_matched_states(scene, direction, kw, seed):
    import torch
    from oracle import engine
    N = 48
    sim = make(scene, direction, kw, N)
    om = oracle_model(sim)
    okw = {k: int(v) for k, v in kw.items()}
    env = engine.Env(om, direction=direction, **okw)
    out0 = env.reset()
    sim.reset()
    info0 = sim.info.cpu().numpy()[0]
    assert [int(info0[I["GRASP"]]), int(info0[I["PHEROMONE"]])] == [out0.grasp, out0.pheromone]
    np.testing.assert_allclose(sim.achieved_goal.cpu().numpy()[0], out0.achieved_goal[:], atol=1e-6)
    np.testing.assert_array_equal(sim.desired_goal.cpu().numpy()[0], np.float32(out0.desired_goal[:]))
    acts = tape(seed, N)
    if not kw.get("include_roll", True):
        acts = acts[:, [0, 1, 2, 4, 5]]
    pre, ref = [], []
    for a in acts:
        # pre-step state rounded to fp32 on both sides
        q, v, c, w = [np.float32(x) for x in (env.qpos, env.data.qvel, env.ctrl, env.data.qacc_warmstart)]
        env.qpos[:], env.data.qvel[:], env.ctrl[:], env.data.qacc_warmstart[:] = q, v, c, w
        env.data.forward_position()
        pre.append((q.copy(), v.copy(), c.copy(), w.copy(), [env.p.contents.gripper_open, env.p.contents.episode_step, 0]))
        o = env.step(a.astype(np.float64))
        ref.append(dict(nsub=(o.nsub_a, o.nsub_b, o.nsub_c), reward=o.reward, done=o.done, status=o.status, pad=(o.grasp, o.pheromone),
                        grasped=o.object_grasped, open=o.gripper_open, reached=(o.reached_target, o.reached_initial, o.fail),
                        ag=np.array(o.achieved_goal[:]), dg=np.array(o.desired_goal[:]), qpos=env.qpos.copy(), target=np.array(o.target_qpos[:]),
                        total=o.total_distance, line=o.line_distance))
        if o.done:
            env.reset()
    sim.set_state(qpos=np.array([p[0] for p in pre]), qvel=np.array([p[1] for p in pre]), ctrl=np.array([p[2] for p in pre]),
                  warmstart=np.array([p[3] for p in pre]), flags=np.array([p[4] for p in pre], np.int32))
    sim.step(torch.tensor(acts, device=sim.device))
    info = sim.info.cpu().numpy()
    st = sim.get_state()
    rew, done = sim.reward.cpu().numpy(), sim.done.cpu().numpy()
    ag, dg = sim.achieved_goal.cpu().numpy(), sim.desired_goal.cpu().numpy()
    flipped = 0
    rew_err, qerrs = [], []
    L = engine.lib()
    from mujoco_rl_manipulate_unknown_objects_b200.vec_env import target_direction
    tdir = np.ascontiguousarray(target_direction(direction), np.float64)
    zero2 = np.zeros(2)
    for i, r in enumerate(ref):
        # (1) the reward FUNCTION at matched inputs (reward.py:18-41 [+ robot_env.py:268-271]): always within 1e-4
        io = np.ascontiguousarray(info[i, I["INIT_OBJ_POS"]:I["INIT_OBJ_POS"] + 3], np.float64)
        fo = np.ascontiguousarray(info[i, I["FINAL_OBJ_POS"]:I["FINAL_OBJ_POS"] + 3], np.float64)
        fr = L.orc_agent_reward(io.ctypes.data, fo.ctypes.data, tdir.ctypes.data, int(info[i, I["GRIPPER_OPEN"]]), zero2.ctypes.data, int(info[i, I["OBJECT_GRASPED"]]))
        if kw.get("her_buffer"):
            fr += 1.0 / np.exp(np.linalg.norm(np.float64(dg[i]) - np.float64(ag[i])))
        assert abs(rew[i] - fr) <= 1e-4, "reward function, env %d: %r vs %r" % (i, rew[i], fr)
        np.testing.assert_allclose(info[i, I["TARGET_QPOS"]:I["TARGET_QPOS"] + 5], r["target"], atol=2e-6, err_msg="IK target, env %d" % i)
        qerr = np.abs(st["qpos"][i] - r["qpos"]).max()
        nsub = tuple(int(x) for x in info[i, I["NSUB_A"]:I["NSUB_A"] + 3])
        if qerr > 1e-4 or nsub != r["nsub"]:
            flipped += 1
            continue
        # (2) end to end over the whole agent step (~100 substeps): the reward is 30 x the object's travel, so it inherits
        # 30 x the trajectory error that the north-star allows over such a horizon
        qerrs.append(qerr)
        assert abs(rew[i] - r["reward"]) <= 1e-4 + 30 * qerr and info[i, I["REWARD"]] == rew[i]
        rew_err.append(abs(rew[i] - r["reward"]))
        assert bool(done[i]) == bool(r["done"]) and int(info[i, I["STATUS"]]) == r["status"]
        assert (int(info[i, I["GRASP"]]), int(info[i, I["PHEROMONE"]])) == r["pad"]
        assert int(info[i, I["OBJECT_GRASPED"]]) == r["grasped"] and int(info[i, I["GRIPPER_OPEN"]]) == r["open"]
        assert tuple(int(x) for x in info[i, I["REACHED_TARGET"]:I["REACHED_TARGET"] + 3]) == r["reached"]
        np.testing.assert_allclose(ag[i], r["ag"], atol=1e-4 + qerr)
        np.testing.assert_allclose(dg[i], r["dg"], atol=1e-4 + qerr)
        assert abs(info[i, I["TOTAL_DISTANCE"]] - r["total"]) < 1e-4 and abs(info[i, I["LINE_DISTANCE"]] - r["line"]) < 1e-4
        assert int(info[i, I["EPISODE_STEP"]]) == pre[i][4][1] + 1
    print("\n[%s dir %d %s] %d matched agent steps, %d flipped by contact chaos, substeps %d, end-state qpos err median %.1e max %.1e, reward err median %.1e max %.1e, rewards>0: %d" % (
        scene, direction, kw, N, flipped, int(info[:, I["NSUB_A"]:I["NSUB_A"] + 3].sum()), np.median(qerrs), max(qerrs), np.median(rew_err), max(rew_err),
        sum(1 for r in ref if r["reward"] > 0)))
    case = "%s-%d-%s" % (scene, direction, "-".join(kw) or "default")
    _FLIPS_SEEN[case] = flipped
    assert flipped <= int(np.ceil(1.5 * MEASURED_FLIPS[case])) + 2, "flip count %d regressed against the measured %d" % (flipped, MEASURED_FLIPS[case])
    assert sum(1 for r in ref if r["reward"] > 0) > 0
    sim.close()


@pytest.mark.gpu
def test_total_flip_count_over_all_cases():
    """Runs after the parametrised cases above (file order): the sum of their flip counts against the measured total."""
    if len(_FLIPS_SEEN) < len(MEASURED_FLIPS):
        pytest.skip("needs every case of test_agent_step_at_matched_states in the same session")
    total = sum(_FLIPS_SEEN.values())
    print("\nflipped agent steps over all cases: %d of %d (measured 41 and 51)" % (total, 48 * len(MEASURED_FLIPS)))
    assert total <= MAX_TOTAL_FLIPS, _FLIPS_SEEN


def test_free_rollout_follows_the_golden_reference_until_contact():
    """End to end without any state injection: the contact-free prefix of the reference python's rollout (the gripper
    approaching the object) is reproduced step by step — substep counts exactly, state to fp32 accuracy."""
    import torch
    for name, scene, direction in [("rollout_sugar_cube_dir0_seed0.npz", "sugar_cube", 0), ("rollout_sand_ball_dir0_seed2.npz", "sand_ball", 0)]:
        g = golden(name)
        sim = make(scene, direction, {}, 4)
        sim.reset()
        np.testing.assert_array_equal(sim.desired_goal.cpu().numpy()[0], g["reset_desired"])
        np.testing.assert_allclose(sim.achieved_goal.cpu().numpy()[0], g["reset_achieved"], atol=1e-6)
        checked = 0
        for i, a in enumerate(g["actions"]):
            # first gripper/object contact (seen in the end-of-step contact list, or as a push reward): chaos starts here
            if (g["contact_geoms"][i][:, 0] > 0).any() or g["reward"][i] > 0.05:
                break
            sim.step(torch.tensor(np.tile(a, (4, 1)), device=sim.device))
            info = sim.info.cpu().numpy()
            assert int(info[0, I["NSUB_A"]:I["NSUB_A"] + 3].sum()) == g["nsub"][i], "step %d" % i
            qp = sim.get_state()["qpos"][0]
            np.testing.assert_allclose(qp[:7], g["qpos"][i][:7], atol=5e-5)   # gripper joints: contact-free, fp32 accuracy
            np.testing.assert_allclose(qp[7:], g["qpos"][i][7:], atol=5e-3)   # object resting / rocking on the floor for ~1000 substeps
            assert abs(info[0, I["REWARD"]] - g["reward"][i]) <= 1e-4 + 30 * np.abs(sim.get_state()["qpos"][0] - g["qpos"][i]).max()
            assert info[0, I["GRASP"]:I["GRASP"] + 2].astype(int).tolist() == g["pad"][i].tolist()
            assert np.all(info == info[0])  # identical environments stay identical
            checked += 1
        assert checked >= 5, checked
        sim.close()


def test_auto_reset_and_episode_bookkeeping():
    """SB3 VecEnv semantics: time-limit termination, reset inside the terminating step, Monitor-style return/length."""
    import torch
    from mujoco_rl_manipulate_unknown_objects_b200 import GripperSim, make_config
    N, H = 64, 5
    sim = GripperSim(make_config(sim_env="/xmls/sand_ball_env.xml", time_horizon=H), num_envs=N, auto_reset=True)
    gen = torch.Generator(device="cuda").manual_seed(0)
    total = torch.zeros(N, device="cuda")
    for t in range(H):
        a = torch.rand((N, 6), device="cuda", generator=gen) * 0.4 - 0.2
        sim.step(a)
        total += sim.reward
        d = sim.done.cpu().numpy()
        if t < H - 1:
            assert not d.any()
    info = sim.info.cpu().numpy()
    assert d.all() and (info[:, I["STATUS"]] == 2).all() and (info[:, I["EPISODE_STEP"]] == H).all()
    np.testing.assert_allclose(info[:, I["EPISODE_RETURN"]], total.cpu().numpy(), atol=1e-5)
    st = sim.get_state()
    assert (st["flags"] == np.array([1, 0, 0])).all()  # fresh episode: gripper open, step 0, RUNNING
    assert np.all(st["qpos"] == st["qpos"][0])
    # returned observation is the reset one, terminal one is kept aside
    assert torch.equal(sim.obs[0], sim.reset_obs) and torch.equal(sim.obs[N - 1], sim.reset_obs)
    assert not torch.equal(sim.terminal_obs[0], sim.reset_obs)
    np.testing.assert_array_equal(sim.desired_goal.cpu().numpy()[0], np.float32([1, 0]))  # robot_env.py:72
    sim.close()


def test_free_running_batch_statistics_match_the_oracle():
    """No state injection at all: 128 environments x 60 agent steps of U(-1,1)^6 actions from reset, auto-reset on, on the GPU and
    in the fp64 oracle (the bench's CPU arm).  Individual trajectories diverge once contacts start (chaotic map), the
    aggregates must not: this pins the workload of bench.py's two arms to each other."""
    import torch
    from oracle import engine
    n, steps = 128, 60
    sim = make("sugar_cube", 0, {}, n, auto_reset=True)
    om = oracle_model(sim)
    rng = np.random.default_rng(7)
    acts = rng.uniform(-1, 1, (n, steps, 6))
    sub_ref, tr_ref, rsum_ref = engine.rollout_threads(om, acts, 8, direction=0)
    sim.reset()
    sub = rew = done = 0.0
    for t in range(steps):
        sim.step(torch.tensor(acts[:, t].astype(np.float32), device=sim.device))
        info = sim.info.cpu().numpy()
        sub += info[:, I["NSUB_A"]:I["NSUB_A"] + 3].sum()
        rew += float(sim.reward.sum())
        done += float(sim.done.sum())
    print("\n[free-running batch] substeps gpu %d oracle %d (%.2f %%), reward sum gpu %.2f oracle %.2f, transitions %d / %d, episodes ended on gpu %d" % (
        sub, sub_ref, 100 * (sub - sub_ref) / sub_ref, rew, rsum_ref, n * steps, tr_ref, done))
    assert tr_ref == n * steps
    assert abs(sub - sub_ref) / sub_ref < 0.03
    assert abs(rew - rsum_ref) <= 0.15 * max(abs(rsum_ref), 1.0) + 1.0
    sim.close()
