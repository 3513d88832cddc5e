"""The SB3 VecEnv contract (stable_baselines3 2.7.1 DummyVecEnv semantics) on the GPU env,
plus the single-env facades, attribute access used by the reference's scripts, DLPack
hand-off and checkpoint/resume."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def test_vecenv_numpy_contract_and_autoreset_infos():
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n = 300
    hr = BatchedChaosVecEnv("hr_sync", n, seed=1)
    assert hr.observation_space.shape == (6,) and hr.action_space.shape == (2,)
    assert hr.observation_space.dtype == np.float32
    assert float(hr.action_space.low.min()) == -1.0 and float(hr.action_space.high.max()) == 1.0
    obs = hr.reset()
    assert obs.shape == (n, 6) and obs.dtype == np.float32
    assert np.all(np.abs(obs) <= 1.0)  # reset clips both halves (lorenz_env_try.py:73-75)
    hr.close()
    env = BatchedChaosVecEnv("lorenz3", n, seed=1, max_episode_steps=5)  # no early termination
    assert env.num_envs == n and env.action_space.shape == (3,)
    obs = env.reset()
    prev = []
    for t in range(1, 11):
        a = np.random.default_rng(t).uniform(-0.05, 0.05, (n, 3)).astype(np.float32)
        env.step_async(a)
        obs, rew, dones, infos = env.step_wait()
        prev.append(obs)
        assert obs.shape == (n, 6) and rew.shape == (n,) and rew.dtype == np.float32
        assert dones.dtype == np.bool_ and len(infos) == n
        if t % 5 == 0:
            assert dones.all()
            for i in (0, n - 1):
                assert infos[i]["TimeLimit.truncated"] is True
                assert infos[i]["terminal_observation"].shape == (6,)
                assert infos[i]["episode"]["l"] == 5
                assert np.isfinite(infos[i]["episode"]["r"])
            assert np.all(np.abs(obs[:, :3]) <= 30.0)   # fresh reset obs (dynamic.py:37)
        else:
            assert not dones.any() and all(not d for d in infos)
    st = env.stats()
    assert st["episodes"] == 2 * n and st["truncated"] == 2 * n
    env.close()


def test_step_wait_output_lifetime():
    """copy_outputs=False: obs / reward are views of a ring of 3 pinned slots -- the arrays returned at
    step t are untouched by steps t+1 and t+2 (SB3's collect_rollouts reads self._last_obs after the
    next env.step) and are overwritten by step t+3.  copy_outputs=True (the default at this size):
    private copies, never overwritten, like DummyVecEnv."""
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n = 512
    rng = np.random.default_rng(0)

    def act():
        return rng.uniform(-0.05, 0.05, (n, 3)).astype(np.float32)

    env = BatchedChaosVecEnv("lorenz3", n, seed=1, copy_outputs=False)
    env.reset()
    obs_t, rew_t, _, _ = env.step(act())
    snap_o, snap_r = obs_t.copy(), rew_t.copy()
    for k in (1, 2):
        env.step(act())
        assert np.array_equal(obs_t, snap_o) and np.array_equal(rew_t, snap_r), f"overwritten after {k} more steps"
    obs_3, _, d_3, _ = env.step(act())
    assert d_3.dtype == np.bool_ and d_3.shape == (n,) and not d_3.any()   # nobody finished: all-False view of the slot
    assert np.shares_memory(obs_3, obs_t)                      # the ring wrapped around ...
    assert not np.array_equal(obs_t, snap_o)                   # ... so the old view now shows step t+3
    env.close()

    env = BatchedChaosVecEnv("lorenz3", n, seed=1)
    assert env._copy_outputs                                   # 512 envs: copies by default
    env.reset()
    kept = [env.step(act())[0] for _ in range(6)]
    snaps = [k.copy() for k in kept]
    for _ in range(4):
        env.step(act())
    assert all(np.array_equal(a, b) for a, b in zip(kept, snaps))
    assert not any(np.shares_memory(kept[0], k) for k in kept[1:])
    env.close()
    big = BatchedChaosVecEnv("lorenz3", 65536, seed=1)
    assert not big._copy_outputs                               # 65,536 envs: zero-copy views by default
    big.close()


def test_env_method_reset_honours_indices():
    """SB3: env_method("reset", indices=[...]) resets only those envs."""
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n = 64
    env = BatchedChaosVecEnv("lorenz3", n, seed=2, max_episode_steps=50)
    env.reset()
    for _ in range(3):
        env.step(np.zeros((n, 3), np.float32))
    before = np.stack(env.get_attr("state1"))
    lens = env.get_attr("current_step")
    assert all(v == 3 for v in lens)
    idx = [1, 7, 40]
    out = env.env_method("reset", indices=idx)
    assert len(out) == 3 and out[0][0].shape == (6,) and out[0][1] == {}
    after = np.stack(env.get_attr("state1"))
    lens = env.get_attr("current_step")
    others = [i for i in range(n) if i not in idx]
    assert np.array_equal(after[others], before[others]) and all(lens[i] == 3 for i in others)
    assert all(lens[i] == 0 for i in idx) and not np.array_equal(after[idx], before[idx])
    assert np.allclose(np.stack([o for o, _ in out])[:, :3], after[idx].astype(np.float32))
    assert len(env.env_method("reset")) == n and all(v == 0 for v in env.get_attr("current_step"))
    env.close()


def test_vecenv_get_set_attr_and_env_method(oracle_api):
    """code/lorenz_pmsm/test_evaluate.py:100-108,123-125 overwrite state1/state2 by attribute,
    call _get_derivatives and read the states back every step."""
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    env = BatchedChaosVecEnv("pmsm_sync", 4, alpha=0.5)
    env.reset()
    env.set_attr("state1", np.array([10.0, -10.0, 15.0], np.float32))
    env.set_attr("state2", np.array([0.0, 0.0, 0.0], np.float32))
    assert np.array_equal(env.get_attr("state1")[2], np.array([10.0, -10.0, 15.0], np.float32))
    d = env.env_method("_get_derivatives", np.array([10.0, -10.0, 15.0], np.float32), [0, 0])[0]
    # lorenz_env_try_pmsm.py:55-57 in float32
    x1, x2, x3 = np.float32(10), np.float32(-10), np.float32(15)
    ref = np.array([-x1 + x2 * x3, -x2 - x1 * x3 + np.float32(20.0) * x3, np.float32(5.46) * (x2 - x3)], np.float32)
    assert np.array_equal(d, ref)
    obs, rew, dones, infos = env.step(np.tile(np.array([[-1.0, 1.0]], np.float32), (4, 1)))
    e = np.asarray(env.get_attr("state1")[0]) - np.asarray(env.get_attr("state2")[0])
    assert e.tolist() == [9.890000343322754, -9.890000343322754, 14.863499641418457]
    assert env.get_attr("sigma")[0] == 5.46 and env.get_attr("current_step")[0] == 1
    assert env.get_attr("render_mode") == [None] * 4
    assert env.env_is_wrapped(object) == [False] * 4
    env.close()


def test_tensor_path_dlpack_no_host_round_trip():
    import torch
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    env = BatchedChaosVecEnv("lorenz_rk4", 4096, substeps=4)
    obs = env.reset_tensor()
    assert obs.is_cuda and obs.shape == (4096, 6) and obs.dtype == torch.float32
    cap = env.obs_dlpack()
    t2 = torch.utils.dlpack.from_dlpack(cap)
    assert t2.data_ptr() == env.batch.obs_planes.data_ptr()  # zero copy
    w = torch.randn(6, 3, device=obs.device)
    for _ in range(3):
        act = torch.tanh(obs @ w)            # policy consumes the strided view directly
        obs, rew, done = env.step_tensor(act)
        assert rew.is_cuda and done.dtype == torch.uint8
    env.close()


def test_single_env_facades_gymnasium_and_old_gym():
    from gym_lorenz_b200 import envs as E
    hr = E.HRSyncEnv(add_noise=False)
    obs, info = hr.reset(seed=3)
    assert obs.shape == (6,) and obs.dtype == np.float32 and info == {}
    out = hr.step(np.array([0.3, -0.2], np.float32))
    assert len(out) == 5 and isinstance(out[1], float) and out[2] is False and out[3] is False
    hr.state_master = np.array([100.0, 0.0, 0.0]); hr.state_slave = np.zeros(3)
    obs, r, term, trunc, _ = hr.step(np.zeros(2, np.float32))
    assert term is True and r == -2000.0          # lorenz_env_try.py:174-176
    hr.close()
    pm = E.PMSM_Sync_Env(alpha=0.5)
    obs, _ = pm.reset(seed=0)
    assert obs.dtype == np.float32
    lam0 = pm.lambda_coef
    pm.step(np.array([1.0, 1.0], np.float32))
    a1 = pm.adam_step
    pm.reset()
    assert pm.adam_step == a1 == 1 and pm.current_step == 0   # Adam state survives reset (:59-75)
    assert lam0 == 0.0
    pm.close()
    lo = E.lorenzEnv_transient()
    o = lo.reset()
    assert o.dtype == np.float64 and o.shape == (6,) and lo.t == 0
    o, r, d, info = lo.step(np.array([600.0, -600.0, 0.0], np.float32))  # clipped to +-500 (dynamic.py:63-65)
    assert isinstance(d, bool) and info == {}
    assert lo._get_current()[1] == 0.0
    pair = E.lorenzEnv_transient.lorenzEnv_transient()
    o = pair.reset()
    tgt = np.asarray(pair.state2).copy()
    pair.step(np.zeros(3, np.float32))
    assert np.array_equal(np.asarray(pair.state2), tgt)  # target never advanced (dynamic.py:194-219)
    lo.close(); pair.close()


def test_checkpoint_resume_is_bit_exact():
    import torch
    n = 2048
    b = H.gpu_batch("hr_sync", n, seed=4, add_noise=True, max_episode_steps=9)
    b.reset()
    a = torch.rand((30, n, 2), device=b.device) * 2 - 1
    for t in range(10):
        b.step(a[t])
    sd = b.state_dict()
    ref = [tuple(x.clone() for x in b.step(a[t])) for t in range(10, 20)]
    b2 = H.gpu_batch("hr_sync", n, seed=4, add_noise=True, max_episode_steps=9)
    b2.load_state_dict(sd)
    for t in range(10, 20):
        o, r, d = b2.step(a[t])
        assert torch.equal(o, ref[t - 10][0]) and torch.equal(r, ref[t - 10][1]) and torch.equal(d, ref[t - 10][2])
    b.close(); b2.close()


def test_reset_distribution_and_mask():
    import torch
    n = 100000
    b = H.gpu_batch("lorenz3", n, seed=123)
    b.reset()
    s = b.state.cpu().numpy()[:3, :n]
    assert s.min() >= -30 and s.max() < 30
    assert abs(s.mean()) < 0.2 and abs(s.std() - 60 / np.sqrt(12)) < 0.1
    # Kolmogorov-Smirnov against U(-30,30)
    from scipy import stats
    assert stats.kstest(s[0], "uniform", args=(-30, 60)).pvalue > 1e-3
    assert abs(np.corrcoef(s[0], s[1])[0, 1]) < 0.02
    before = b.state.clone()
    mask = torch.zeros(n, dtype=torch.uint8, device=b.device); mask[::2] = 1
    b.reset(mask)
    after = b.state
    assert torch.equal(after[:, 1:n:2], before[:, 1:n:2])
    assert not torch.equal(after[:3, 0:n:2], before[:3, 0:n:2])
    b.close()


@pytest.mark.parametrize("kind", ["lorenz3", "hr_sync", "pmsm_sync", "lorenz_rk4"])
def test_graph_mode_replay_equals_eager_stepping(kind):
    """cl_set_graph_mode: the Philox step index lives on the device, so a captured sequence of
    step launches can be replayed and still advances the random streams (auto-resets, noise)."""
    import torch
    n, per_graph, replays = 3000, 6, 4
    kw = dict(seed=31, autoreset=True, max_episode_steps=5)
    if kind in ("hr_sync", "pmsm_sync"):
        kw["add_noise"] = True
    amp = 0.05 if kind == "lorenz3" else 1.0
    g0 = torch.Generator(device="cpu").manual_seed(2)
    acts = ((torch.rand((per_graph, n, H.gpu_batch(kind, 1).act_dim), generator=g0) * 2 - 1) * amp).to("cuda:0")
    # eager reference: per_graph * replays steps, cycling through the same action tensors
    e = H.gpu_batch(kind, n, **kw)
    e.reset()
    for r in range(replays):
        for t in range(per_graph):
            e.step(acts[t])
    torch.cuda.synchronize()
    # graph: capture per_graph steps once, replay
    b = H.gpu_batch(kind, n, **kw)
    b.reset()
    b.set_graph_mode(True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):          # warm-up outside capture (does not advance: we restore below)
        sd = b.state_dict()
        b.step(acts[0])
        b.load_state_dict(sd)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for t in range(per_graph):
            b.step(acts[t])
    for r in range(replays):
        graph.replay()
    torch.cuda.synchronize()
    assert b.step_index == e.step_index == 1 + per_graph * replays
    assert torch.equal(torch.nan_to_num(b.state), torch.nan_to_num(e.state))
    assert torch.equal(b.ep_len, e.ep_len) and torch.equal(b.ep_return, e.ep_return)
    assert b.stats()["episodes"] == e.stats()["episodes"] > 0
    b.set_graph_mode(False)
    assert b.step_index == e.step_index
    b.close(); e.close()


@pytest.mark.parametrize("kind", ["lorenz_rk4", "hr_sync", "pmsm_sync", "memristive4_pair"])   # the last one: per-block episode statistics
def test_host_step_modes_give_identical_results(kind):
    """DMA chain, zero-copy, the sliced two-stream pipeline and the streamed mode (kernel launched before
    the caller's array is staged, blocks wait for their slice's generation flag) are the same computation: every
    obs / reward / done / info must agree bit for bit, at a batch size that is not a multiple of
    the slice granularity and with episodes ending inside the window."""
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n, T = 20000 + 37, 9
    modes = [("dma", 1), ("zerocopy", 1), ("pipelined", 2), ("pipelined", 3), ("pipelined", 7), ("pipelined", 64),
             ("streamed", 1), ("streamed", 3), ("streamed", 16), ("streamed", 64),
             # staging lanes (copy threads): 1 .. 4, slice counts that do and do not divide by the lanes
             ("streamed", 16, 1), ("streamed", 5, 2), ("streamed", 16, 3), ("streamed", 2, 4), ("streamed", 37, 4),
             # negative: the relay as its own kernel on the side stream instead of block 0 of the step kernel
             ("streamed", 16, -1), ("streamed", 37, -3)]
    rng = np.random.default_rng(5)
    ref = None
    for mode, k, *threads in modes:
        os.environ.pop("CHAOS_B200_COPY_THREADS", None)
        os.environ.pop("CHAOS_B200_RELAY", None)
        if threads:
            os.environ["CHAOS_B200_COPY_THREADS"] = str(abs(threads[0]))
            if threads[0] < 0:
                os.environ["CHAOS_B200_RELAY"] = "kernel"
        env = BatchedChaosVecEnv(kind, n, seed=3, max_episode_steps=4)
        env.batch.set_host_mode(mode, k)        # creates the staging context (reads the variables)
        os.environ.pop("CHAOS_B200_COPY_THREADS", None)
        os.environ.pop("CHAOS_B200_RELAY", None)
        a_rng = np.random.default_rng(11)
        lo, hi = env.action_space.low, env.action_space.high
        trace = [env.reset().copy()]
        for t in range(T):
            a = a_rng.uniform(-1, 1, (n, env.action_space.shape[0])).astype(np.float32) * np.minimum(hi, 1.0)
            if t % 2:   # alternate: user ndarray (staged slice by slice) / pinned staging buffer
                env.batch.host_action_buffer()[:] = a
                env.batch.step_host_async(None); env._waiting = True
                obs, rew, done, infos = env.step_wait()
            else:
                obs, rew, done, infos = env.step(a)
            trace += [obs.copy(), rew.copy(), done.copy()]
            if done.any():
                i = int(np.flatnonzero(done)[-1])
                trace += [np.asarray(infos[i]["terminal_observation"]).copy(),
                          np.float64(infos[i]["episode"]["r"]), np.int64(infos[i]["episode"]["l"])]
        trace.append(env.batch.state.cpu().numpy().copy()[:, :n])
        env.close()
        if ref is None:
            ref = trace
            assert any(np.asarray(x).dtype == bool and np.asarray(x).any() for x in trace)
        else:
            assert len(trace) == len(ref)
            for k_, (x, y) in enumerate(zip(trace, ref)):
                assert np.array_equal(np.asarray(x), np.asarray(y), equal_nan=True), (mode, k, k_)


def test_tensor_path_has_no_host_round_trip():
    """SURVEY 8d cfg 5: on the tensor / DLPack path observations and rewards reach the policy as device
    tensors with NO host round trip.  64 policy -> step_tensor iterations under torch.profiler (CUPTI
    activity records): the loop contains our step kernel 64 times and not a single host<->device memcpy."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n = 4096
    env = BatchedChaosVecEnv("hr_sync", n, seed=3)
    net = torch.nn.Sequential(torch.nn.Linear(6, 32), torch.nn.Tanh(), torch.nn.Linear(32, 2), torch.nn.Tanh()).to("cuda:0")
    obs = env.reset_tensor()
    with torch.no_grad():
        for _ in range(4):                         # warm-up outside the profiled region
            obs, rew, done = env.step_tensor(net(obs))
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            ret = torch.zeros(n, dtype=torch.float64, device="cuda:0")
            for _ in range(64):
                obs, rew, done = env.step_tensor(net(obs))
                ret += rew
            cap = torch.utils.dlpack.from_dlpack(env.obs_dlpack())     # the DLPack hand-off itself
            torch.cuda.synchronize()
    assert cap.data_ptr() == env.batch.obs_planes.data_ptr()
    names = [e.name for e in prof.events()]
    ours = [x for x in names if "k_step" in x]
    if not ours:
        pytest.skip("no CUDA activity records (CUPTI unavailable on this box)")
    assert len(ours) == 64, len(ours)
    copies = [x for x in names if "memcpy" in x.lower() and ("htod" in x.lower() or "dtoh" in x.lower())]
    assert copies == [], copies
    assert bool(torch.isfinite(ret).all())
    env.close()


def test_streamed_host_mode_survives_synchronous_launches(tmp_path):
    """With CUDA_LAUNCH_BLOCKING=1 (or under a profiler) a launch returns only when the kernel has finished,
    so the CPU cannot stage the action array while the step kernel runs.  The relay kernel notices that
    nothing is published, calls the step off before any block has stored anything, and step_wait redoes it
    as a zero-copy step: same results as zero-copy mode, one fallback recorded, later steps not streamed."""
    import json
    import subprocess
    import sys
    script = tmp_path / "run.py"
    script.write_text(
        "import sys, json, hashlib\n"
        f"sys.path.insert(0, {str(H.os.path.dirname(H.os.path.dirname(H.os.path.abspath(__file__))))!r})\n"
        "import numpy as np\n"
        "from gym_lorenz_b200.vec_env import BatchedChaosVecEnv\n"
        "mode = sys.argv[1]\n"
        "env = BatchedChaosVecEnv('hr_sync', 20000, seed=3, max_episode_steps=4)\n"
        "env.batch.set_host_mode(mode, 8)\n"
        "env.reset()\n"
        "rng = np.random.default_rng(1)\n"
        "h = hashlib.sha256()\n"
        "for t in range(6):\n"
        "    obs, rew, done, infos = env.step(rng.uniform(-1, 1, (20000, 2)).astype(np.float32))\n"
        "    h.update(obs.tobytes()); h.update(rew.tobytes()); h.update(done.tobytes())\n"
        "print(json.dumps({'hash': h.hexdigest(), 'fallbacks': int(env.batch.streamed_fallbacks)}))\n")
    out = {}
    for mode, blocking in (("zerocopy", "0"), ("streamed", "0"), ("streamed", "1")):
        env = dict(H.os.environ, CUDA_LAUNCH_BLOCKING=blocking)
        r = subprocess.run([sys.executable, str(script), mode], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[(mode, blocking)] = json.loads(r.stdout.strip().splitlines()[-1])
    assert out[("streamed", "0")]["hash"] == out[("zerocopy", "0")]["hash"]
    assert out[("streamed", "1")]["hash"] == out[("zerocopy", "0")]["hash"]
    assert out[("streamed", "0")]["fallbacks"] == 0
    assert out[("streamed", "1")]["fallbacks"] == 1


def test_nvtx_ranges_option(tmp_path):
    """CHAOS_B200_NVTX=1 wraps the launching entry points in NVTX ranges; results do not change."""
    import subprocess
    import sys
    script = tmp_path / "run.py"
    script.write_text(
        "import sys, hashlib\n"
        f"sys.path.insert(0, {str(H.os.path.dirname(H.os.path.dirname(H.os.path.abspath(__file__))))!r})\n"
        "import numpy as np, torch\n"
        "from gym_lorenz_b200.core import ChaosBatch\n"
        "from gym_lorenz_b200.vec_env import BatchedChaosVecEnv\n"
        "env = BatchedChaosVecEnv('hr_sync', 3000, seed=3)\n"
        "env.reset()\n"
        "rng = np.random.default_rng(1)\n"
        "h = hashlib.sha256()\n"
        "for t in range(4):\n"
        "    obs, rew, done, infos = env.step(rng.uniform(-1, 1, (3000, 2)).astype(np.float32))\n"
        "    h.update(obs.tobytes()); h.update(rew.tobytes())\n"
        "out = env.batch.rollout(3)\n"
        "h.update(out['reward'].cpu().numpy().tobytes())\n"
        "print(h.hexdigest(), ChaosBatch.step.__name__, hasattr(ChaosBatch.step, '__wrapped__') or 'wrapper' in repr(ChaosBatch.step))\n")
    outs = []
    for flag in ("0", "1"):
        env = dict(H.os.environ, CHAOS_B200_NVTX=flag)
        r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1].split())
    assert outs[0][0] == outs[1][0]                     # same results
    assert outs[0][2] == "False" and outs[1][2] == "True"   # ranges only when asked for
