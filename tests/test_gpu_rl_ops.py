"""Device RL plumbing (csrc/tu_rl_ops.cu) against the NumPy restatements of SB3 / the reference's
metric functions (oracle/sb3_ref.py)."""
import numpy as np
import pytest

import helpers as H
from oracle import sb3_ref as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,N", [(128, 4096), (2048, 33), (16, 65536)])
def test_gae_bit_exact_vs_sb3_restatement(T, N):
    import torch
    from gym_lorenz_b200 import rl_ops
    rng = np.random.default_rng(T)
    r = rng.normal(size=(T, N)).astype(np.float32)
    v = rng.normal(size=(T, N)).astype(np.float32)
    starts = (rng.random((T, N)) < 0.02).astype(np.float32)
    lv = rng.normal(size=N).astype(np.float32)
    dones = (rng.random(N) < 0.1).astype(np.float32)
    adv_ref, ret_ref = S.gae(r, v, starts, lv, dones, 0.99, 0.95)
    dev = "cuda:0"
    adv, ret = rl_ops.gae(*(torch.as_tensor(x, device=dev) for x in (r, v, starts, lv, dones)), 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy(), adv_ref)
    assert np.array_equal(ret.cpu().numpy(), ret_ref)


def test_running_moments_and_normalize():
    import torch
    from gym_lorenz_b200 import rl_ops
    rng = np.random.default_rng(3)
    dev = torch.device("cuda:0")
    rms_d = rl_ops.RunningMeanStd((6,), dev)
    rms_o = S.RunningMeanStd(shape=(6,))
    for k in range(5):
        x = (rng.normal(2.0 * k, 1.0 + k, size=(8192, 6))).astype(np.float32)
        rms_d.update(torch.as_tensor(x, device=dev))
        rms_o.update(x)
        assert np.allclose(rms_d.mean.cpu().numpy(), rms_o.mean, rtol=2e-6, atol=1e-6)
        assert np.allclose(rms_d.var.cpu().numpy(), rms_o.var, rtol=2e-5)
        assert np.isclose(rms_d.count, rms_o.count)
    # normalisation itself is exact given the same statistics (float64 math, one rounding to f32)
    class V:  # minimal venv stand-in
        num_envs = 8192
        class batch:
            obs_dim, device = 6, dev
    vn = rl_ops.DeviceVecNormalize(V, clip_obs=2.0)
    vn.obs_rms = rms_d
    z = vn.normalize_obs(torch.as_tensor(x, device=dev)).cpu().numpy()
    rms_same = S.RunningMeanStd(shape=(6,))
    rms_same.mean, rms_same.var = rms_d.mean.cpu().numpy(), rms_d.var.cpu().numpy()
    assert np.array_equal(z, S.normalize_obs(x, rms_same, clip_obs=2.0))
    # SoA (strided) input view gives the same result
    xs = torch.as_tensor(np.ascontiguousarray(x.T), device=dev).t()
    assert np.array_equal(vn.normalize_obs(xs).cpu().numpy(), z)


def test_frame_stack_exact():
    import torch
    from gym_lorenz_b200 import rl_ops
    rng = np.random.default_rng(4)
    N, dim, k = 5000, 6, 4
    dev = torch.device("cuda:0")

    class V:
        num_envs = N
        class batch:
            obs_dim, device = dim, dev
    fs = rl_ops.DeviceVecFrameStack(V, k)
    ref = np.zeros((N, dim * k), np.float32)
    for t in range(9):
        obs = rng.normal(size=(N, dim)).astype(np.float32)
        done = (rng.random(N) < 0.2).astype(np.uint8)
        out = fs._push(torch.as_tensor(obs, device=dev), torch.as_tensor(done, device=dev))
        ref = S.frame_stack_update(ref, obs, done)
        assert np.array_equal(out.cpu().numpy(), ref)


def test_eval_metrics_vs_reference_functions():
    import torch
    from gym_lorenz_b200 import rl_ops
    rng = np.random.default_rng(5)
    T, N = 2400, 64
    t = np.arange(T)[:, None, None]
    tau = rng.uniform(20, 400, size=(1, 3, N))
    e = rng.uniform(0.5, 30, size=(1, 3, N)) * np.exp(-t / tau) * np.cos(t / 7.0) + 0.01 * rng.normal(size=(T, 3, N))
    e[-1, 0, 0] = 1.0          # env 0: component 0 never settles
    e[-1, :, 1] = 1.0          # env 1: nothing settles -> NaN
    e[:, :, 2] = 0.001         # env 2: never leaves the band -> 0
    u = rng.normal(size=(T, 2, N)) * 50
    out = rl_ops.eval_metrics(torch.as_tensor(e, device="cuda:0"), torch.as_tensor(u, device="cuda:0"), dt=0.001)
    for i in range(N):
        mae, rmse, ts, en = S.steady_state_metrics(e[:, :, i], u[:, :, i], dt=0.001)
        assert np.isclose(out["mae"][i].item(), mae, rtol=1e-12)
        assert np.isclose(out["rmse"][i].item(), rmse, rtol=1e-12)
        assert np.isclose(out["energy"][i].item(), en, rtol=1e-12)
        got = out["settling_time"][i].item()
        assert (np.isnan(ts) and np.isnan(got)) or got == ts, (i, got, ts)
    assert np.isnan(out["settling_time"][1].item()) and out["settling_time"][2].item() == 0.0


def test_device_vecnormalize_framestack_on_real_env():
    """PMSM pipeline of code/lorenz_pmsm/train.py:166-170 (VecNormalize, clip_obs=10) and the
    filter pipeline of code/lorenz_filter/train.py:115 (VecFrameStack 4) on the tensor path."""
    import torch
    from gym_lorenz_b200 import rl_ops
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n = 4096
    env = BatchedChaosVecEnv("pmsm_sync", n, alpha=0.5, seed=2, max_episode_steps=7)
    vn = rl_ops.DeviceVecNormalize(env, norm_obs=True, norm_reward=False, clip_obs=10.0)
    fs = rl_ops.DeviceVecFrameStack(vn, 4)
    rms = S.RunningMeanStd(shape=(6,))
    stack = np.zeros((n, 24), np.float32)
    obs = fs.reset_tensor()
    raw = vn.get_original_obs().cpu().numpy()
    rms.update(raw)
    stack = S.frame_stack_update(stack, S.normalize_obs(raw, rms), np.zeros(n, bool))
    assert np.allclose(obs.cpu().numpy(), stack, rtol=1e-5, atol=1e-5)
    g = torch.Generator(device="cpu").manual_seed(0)
    for t in range(10):
        a = (torch.rand((n, 2), generator=g) * 2 - 1).to("cuda:0")
        obs, rew, done = fs.step_tensor(a)
        raw = vn.get_original_obs().cpu().numpy()
        rms.update(raw)
        stack = S.frame_stack_update(stack, S.normalize_obs(raw, rms), done.cpu().numpy() != 0)
        assert np.allclose(obs.cpu().numpy(), stack, rtol=2e-5, atol=2e-5), t
        assert rew.dtype == torch.float32
    assert np.allclose(vn.obs_rms.mean.cpu().numpy(), rms.mean, rtol=1e-5, atol=1e-5)
    env.close()


def test_device_rollout_collector_matches_sb3_bookkeeping():
    import torch
    from gym_lorenz_b200 import rl_ops
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n, T = 2048, 24
    env = BatchedChaosVecEnv("hr_sync", n, seed=5, max_episode_steps=10)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 32), torch.nn.Tanh(), torch.nn.Linear(32, 3)).to("cuda:0")

    def policy(obs):
        y = net(obs)
        return torch.tanh(y[:, :2]), y[:, 2], -0.5 * (y[:, :2] ** 2).sum(1)

    col = rl_ops.DeviceRolloutCollector(env, policy, n_steps=T, gamma=0.99, gae_lambda=0.95)
    out = col.collect()
    for k in ("obs", "actions", "rewards", "values", "log_probs", "episode_starts", "advantages", "returns"):
        assert out[k].is_cuda and out[k].shape[0] == T
    starts = out["episode_starts"].cpu().numpy()
    assert starts[0].all()                                   # SB3: _last_episode_starts = ones at the start
    assert starts[10].all() and starts[20].all() and not starts[5].any()   # TimeLimit 10, no early termination expected
    # GAE of the collected buffers equals the SB3 restatement bit-for-bit
    with torch.no_grad():
        lv = policy(col._last_obs)[1].cpu().numpy()
    adv, ret = S.gae(out["rewards"].cpu().numpy(), out["values"].cpu().numpy(), starts, lv,
                     col._last_starts.cpu().numpy(), 0.99, 0.95)
    assert np.array_equal(out["advantages"].cpu().numpy(), adv)
    assert np.array_equal(out["returns"].cpu().numpy(), ret)
    # the truncated step's reward carries the bootstrap gamma * V(terminal_obs)
    env2 = BatchedChaosVecEnv("hr_sync", n, seed=5, max_episode_steps=10)
    obs = env2.reset_tensor().clone()
    with torch.no_grad():
        for t in range(10):
            a, _, _ = policy(obs)
            o, r, d = env2.step_tensor(a)
            raw_r = r.float().clone()
            obs = o.clone()
        tv = policy(env2.batch.terminal_obs())[1]
    assert torch.allclose(out["rewards"][9], raw_r + 0.99 * tv, rtol=1e-6, atol=1e-6)
    env.close(); env2.close()


def test_cuda_graph_collector_equals_eager_collector():
    """The whole n_steps loop captured in one CUDA graph (env in graph mode: device-resident
    Philox step index) reproduces the eager loop bit-for-bit, replay after replay -- including
    auto-resets, whose random initial conditions depend on the advancing step index."""
    import torch
    from gym_lorenz_b200 import rl_ops
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n, T = 1024, 12
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3)).to("cuda:0")

    def policy(obs):                      # deterministic: the two runs must match exactly
        y = net(obs)
        return torch.tanh(y[:, :2]), y[:, 2], -(y[:, :2] ** 2).sum(1)

    outs = {}
    for mode in (False, True):
        env = BatchedChaosVecEnv("hr_sync", n, seed=9, max_episode_steps=5)
        col = rl_ops.DeviceRolloutCollector(env, policy, n_steps=T, use_cuda_graph=mode)
        rolls = []
        for _ in range(3):
            o = col.collect()
            torch.cuda.synchronize()
            rolls.append({k: v.clone() for k, v in o.items()})
        outs[mode] = (rolls, env.batch.step_index, env.stats())
        env.close()
    for r_e, r_g in zip(outs[False][0], outs[True][0]):
        for k in r_e:
            assert torch.equal(r_e[k], r_g[k]), k
    assert outs[False][1] == outs[True][1]             # same number of Philox steps consumed
    assert outs[False][2]["episodes"] == outs[True][2]["episodes"] > 0
    # and the rollouts differ from one another (the streams really advance across replays)
    assert not torch.equal(outs[True][0][0]["obs"], outs[True][0][1]["obs"])


def test_vecnormalize_pipeline_replays_as_cuda_graph():
    """VecNormalize(env) -- the PMSM pipeline of code/lorenz_pmsm/train.py:115-118 -- captured once as a
    CUDA graph and replayed: running statistics (device-resident count), discounted returns, env
    state and outputs must equal eager stepping bit for bit."""
    import torch
    from gym_lorenz_b200 import rl_ops
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n, per_graph, replays = 4096, 5, 4
    g0 = torch.Generator(device="cpu").manual_seed(7)
    acts = (torch.rand((per_graph, n, 2), generator=g0) * 2 - 1).to("cuda:0")

    def make():
        env = BatchedChaosVecEnv("pmsm_sync", n, seed=5, max_episode_steps=7)
        vn = rl_ops.DeviceVecNormalize(env, clip_obs=10.0, gamma=0.99)
        vn.reset_tensor()
        return env, vn

    env_e, vn_e = make()
    for r in range(replays):
        for t in range(per_graph):
            out_e = vn_e.step_tensor(acts[t])
    torch.cuda.synchronize()

    env_w, vn_w = make()                      # throw-away twin: warms the kernels up on a side stream
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        vn_w.step_tensor(acts[0])
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    env_w.close()

    env_g, vn_g = make()
    env_g.batch.set_graph_mode(True)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for t in range(per_graph):
            out_g = vn_g.step_tensor(acts[t])
    for r in range(replays):
        graph.replay()
    torch.cuda.synchronize()
    # the batch sums are accumulated with floating-point atomics: their order, hence the last bits of
    # the statistics, differs from run to run (eager or not) -- everything else is exact
    assert torch.allclose(out_g[0], out_e[0], rtol=1e-6, atol=1e-6) and torch.allclose(out_g[1], out_e[1], rtol=1e-6)
    assert torch.equal(out_g[2], out_e[2])
    for rg, re_ in ((vn_g.obs_rms, vn_e.obs_rms), (vn_g.ret_rms, vn_e.ret_rms)):
        assert torch.allclose(rg.mean, re_.mean, rtol=1e-12, atol=1e-13) and torch.allclose(rg.var, re_.var, rtol=1e-12)
        assert rg.count == re_.count
    assert vn_e.obs_rms.count == pytest.approx(1e-4 + n * (1 + per_graph * replays))
    assert torch.equal(vn_g.returns, vn_e.returns)      # raw rewards: independent of the statistics
    assert torch.equal(env_g.batch.state, env_e.batch.state)
    assert env_g.batch.stats()["episodes"] == env_e.batch.stats()["episodes"] > 0
    env_g.batch.set_graph_mode(False)
    env_g.close(); env_e.close()


def test_frame_stack_terminal_observation_follows_sb3_stacked_observations():
    """SB3 StackedObservations.update: for a finished env infos["terminal_observation"] becomes the
    previous stack rolled by one frame with the env's terminal observation as last frame -- here with
    VecNormalize underneath (the normalised terminal observation), as in FrameStack(VecNormalize(env))."""
    import torch
    from gym_lorenz_b200 import rl_ops
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n, k = 3000 + 5, 4
    env = BatchedChaosVecEnv("hr_sync", n, seed=4, max_episode_steps=6)
    vn = rl_ops.DeviceVecNormalize(env, norm_obs=True, norm_reward=False, clip_obs=10.0)
    fs = rl_ops.DeviceVecFrameStack(vn, k)
    prev = fs.reset_tensor().clone()
    g = torch.Generator(device="cpu").manual_seed(0)
    seen = 0
    for t in range(14):
        a = (torch.rand((n, 2), generator=g) * 2 - 1).to("cuda:0")
        obs, rew, done = fs.step_tensor(a)
        d = done != 0
        if bool(d.any()):
            want = torch.cat([prev[:, 6:], vn.terminal_obs()], dim=1)
            assert torch.equal(fs.terminal_obs()[d], want[d]), t
            assert torch.equal(obs[d][:, :18], torch.zeros_like(obs[d][:, :18]))     # stack restarts
            seen += int(d.sum())
        nd = ~d
        assert torch.equal(obs[nd][:, :18], prev[nd][:, 6:])
        prev = obs.clone()
    assert seen >= 2 * n
    env.close()


@pytest.mark.parametrize("graph", [False, True])
def test_collector_over_framestack_of_vecnormalize(graph):
    """The reference's lorenz_filter pipeline shape (VecFrameStack(4) around the env, PPO on top) with
    VecNormalize in between and a short TimeLimit, so the TimeLimit bootstrap evaluates the policy on
    stacked, normalised terminal observations every few steps; eager and as one CUDA graph."""
    import torch
    from gym_lorenz_b200 import rl_ops
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n, T, k = 1024, 12, 4
    torch.manual_seed(2)
    net = torch.nn.Sequential(torch.nn.Linear(6 * k, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3)).to("cuda:0")

    def policy(obs):
        assert obs.shape == (n, 6 * k)
        y = net(obs)
        return torch.tanh(y[:, :2]), y[:, 2], -(y[:, :2] ** 2).sum(1)

    env = BatchedChaosVecEnv("hr_sync", n, seed=9, max_episode_steps=5)
    fs = rl_ops.DeviceVecFrameStack(rl_ops.DeviceVecNormalize(env, norm_reward=False), k)
    col = rl_ops.DeviceRolloutCollector(fs, policy, n_steps=T, use_cuda_graph=graph)
    assert col.batch is env.batch and col.obs_dim == 6 * k
    for _ in range(3):
        out = col.collect()
        torch.cuda.synchronize()
    assert out["obs"].shape == (T, n, 6 * k) and bool(torch.isfinite(out["advantages"]).all())
    starts = out["episode_starts"].cpu().numpy()
    assert starts.any() and not starts.all()
    env.close()


def test_vecnormalize_return_statistics_follow_sb3():
    """SB3 VecNormalize.step_wait: the discounted returns and ret_rms are updated whenever `training`
    is set (norm_reward only decides whether rewards are rescaled), from float64 returns."""
    import torch
    from gym_lorenz_b200 import rl_ops
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n = 2048
    env = BatchedChaosVecEnv("pmsm_sync", n, seed=6, max_episode_steps=9)
    vn = rl_ops.DeviceVecNormalize(env, norm_obs=True, norm_reward=False, gamma=0.99)
    vn.reset_tensor()
    rms = S.RunningMeanStd(shape=())
    returns = np.zeros(n, np.float64)
    g = torch.Generator(device="cpu").manual_seed(1)
    for t in range(12):
        a = (torch.rand((n, 2), generator=g) * 2 - 1).to("cuda:0")
        _, rew, done = vn.step_tensor(a)
        assert torch.equal(rew, vn.old_reward)                      # norm_reward=False: rewards untouched
        returns = returns * 0.99 + vn.old_reward.double().cpu().numpy()
        rms.update(returns)
        returns[done.cpu().numpy() != 0] = 0.0
        assert np.array_equal(vn.returns.cpu().numpy(), returns), t
    assert np.isclose(vn.ret_rms.mean.item(), rms.mean, rtol=1e-12)
    assert np.isclose(vn.ret_rms.var.item(), rms.var, rtol=1e-11)
    assert np.isclose(vn.ret_rms.count, rms.count, rtol=1e-15)
    env.close()


def test_caller_supplied_buffers_are_validated():
    import torch
    b = H.gpu_batch("lorenz_rk4", 1000, seed=1)
    b.reset()
    T, NP = 4, b.n_pad
    acts = torch.zeros((T, 1000, 3), device=b.device)
    good = {"obs": torch.empty((T, 6, NP), dtype=torch.float32, device=b.device),
            "reward": torch.empty((T, NP), dtype=torch.float64, device=b.device),
            "done": torch.empty((T, NP), dtype=torch.uint8, device=b.device)}
    b.rollout(T, acts, out=dict(good))
    for key, bad in (("obs", torch.empty((T, 6, NP - 128), dtype=torch.float32, device=b.device)),
                     ("obs", torch.empty((T, 6, NP), dtype=torch.float64, device=b.device)),
                     ("reward", torch.empty((T, NP), dtype=torch.float32, device=b.device)),
                     ("reward", torch.empty((T, 2 * NP), dtype=torch.float64, device=b.device)[:, ::2]),
                     ("done", torch.empty((T, NP), dtype=torch.uint8)),
                     ("done", torch.empty((T - 1, NP), dtype=torch.uint8, device=b.device))):
        o = dict(good); o[key] = bad
        with pytest.raises(ValueError):
            b.rollout(T, acts, out=o)
    with pytest.raises(ValueError):
        b.rollout(T, acts, out={"obs": torch.empty((T, 999, 6), dtype=torch.float32, device=b.device)}, obs_layout="rows")
    h = H.gpu_batch("hr_sync", 256, add_noise=True)
    h.reset()
    a = torch.zeros((256, 2), device=h.device)
    h.step(a, noise=torch.zeros((3, h.n_pad), dtype=torch.float64, device=h.device))
    for bad in (torch.zeros((3, 256 - 1), dtype=torch.float64, device=h.device),
                torch.zeros((3, h.n_pad), dtype=torch.float32, device=h.device),
                torch.zeros((2, h.n_pad), dtype=torch.float64, device=h.device)):
        with pytest.raises(ValueError):
            h.step(a, noise=bad)
    b.close(); h.close()
