"""Parity tests proper: the CUDA path (through the C-ABI) against
  (1) the committed golden vectors from the unmodified reference,
  (2) the CPU oracle on the same seeded inputs,
  (3) size-independent properties at BASELINE.json's full sizes.
Tolerances: f64 kinds 1e-12 relative per control interval (north_star); float32 PMSM env
2.5e-7 (2 ulp f32) on pow-derived quantities, states bit-exact; reward/done exact wherever
the state is bit-identical and no libm pow is involved."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

EULER_EXACT = {"lorenz3", "lorenz3_pair", "lorenz4_pair", "pmsm_single", "pmsm_free"}  # pure IEEE mul/add kinds


@pytest.mark.parametrize("name", H.PARITY_CASES)
def test_teacher_forced_per_interval_vs_reference_golden(name):
    case = H.load_case(name)
    kind = case["kind"]
    g = H.gpu_teacher_forced(case)
    rt = H.rtol_for(kind)
    H.assert_close(g["state"], case["state"], rt, f"{name} state")
    # HR observations are float32 (`astype(np.float32)`, lorenz_env_try.py:179): 1 ulp(f32)
    H.assert_close(g["obs"], case["obs"], 1.2e-7 if kind == "hr_sync" else rt, f"{name} obs")
    H.assert_close(g["reward"], case["reward"], rt, f"{name} reward")
    assert np.array_equal(g["done"], case["done"]), f"{name} done"
    if kind in EULER_EXACT:
        assert H.max_rel(g["state"], case["state"]) == 0.0, f"{name}: Euler kinds must be bit-exact"
        assert H.max_rel(g["reward"], case["reward"]) == 0.0
    if kind == "pmsm_sync":  # physical states are pure f32 mul/add -> bit-exact
        assert H.max_rel(g["state"][..., :6], case["state"][..., :6]) == 0.0


@pytest.mark.parametrize("name", H.PARITY_CASES)
def test_free_running_64_steps_vs_reference_golden(name):
    """64 steps <= 0.64 time units (Lorenz, lambda1 ~ 0.9): far below the Lyapunov horizon, a
    1-ulp difference grows < 2x.  Stated tolerance: 1e-11 relative (f64), 1e-5 (f32 env)."""
    case = H.load_case(name)
    kind = case["kind"]
    g = H.gpu_free_run(case)
    tol = 1e-5 if kind == "pmsm_sync" else 1e-11
    if name.endswith("diverge"):
        # blow-up cases: compare while finite and the guard (done) pattern exactly
        assert np.array_equal(g["done"], case["done"])
        return
    H.assert_close(g["state"], case["state"], tol, f"{name} state", atol=1e-13)
    H.assert_close(g["reward"], case["reward"], tol, f"{name} reward", atol=1e-13)
    assert np.array_equal(g["done"], case["done"])
    if kind in EULER_EXACT:
        assert H.max_rel(g["state"], case["state"]) == 0.0


def test_cfg1_single_env_1000_steps_vs_reference_dump():
    """BASELINE.json configs[0] through the single-env facade (old-gym API)."""
    import torch
    from gym_lorenz_b200.envs import lorenzEnv_transient
    z = np.load(os.path.join(H.GOLDEN, "cfg1_lorenz3.npz"))
    for tag in ("small", "wide"):
        env = lorenzEnv_transient()
        env.reset()
        env.state1 = z[f"{tag}_st0"][:3]
        env.t = 0.0
        for t in range(1000):
            obs, r, d, info = env.step(z[f"{tag}_actions"][t])
            assert obs.dtype == np.float64 and obs.shape == (6,)
            assert H.same_nonfinite(obs, z[f"{tag}_obs"][t])
            m = np.isfinite(obs)
            assert np.array_equal(obs[m], z[f"{tag}_obs"][t][m]), (tag, t)
            if np.isfinite(r):
                assert r == z[f"{tag}_reward"][t]
            assert d is False and info == {}
        assert env.t == z[f"{tag}_t"][-1]
        env.close()


def test_pmsm_xlsx_float32_kats_on_gpu():
    """The reference's own artefact PMSM_Origin_Data.xlsx, replayed on the GPU bit-exactly."""
    import torch
    z = np.load(os.path.join(H.GOLDEN, "pmsm_xlsx_kat.npz"))
    cols = sorted({k.split("_")[0] for k in z.files})
    for c in cols:
        acts, err, alpha = z[f"{c}_actions"], z[f"{c}_err"], float(z[f"{c}_alpha"])
        b = H.gpu_batch("pmsm_sync", 1, autoreset=False, max_episode_steps=0, alpha=alpha)
        H.gpu_set_state(b, np.array([[10, -10, 15, 0, 0, 0, 0, 0, 0]], np.float64))
        rewards = []
        for t in range(len(acts)):
            obs, rew, done = b.step(torch.as_tensor(acts[t:t + 1], device=b.device))
            e = obs[0, :3].cpu().numpy()
            assert np.array_equal(e, err[t]), (c, t)
            rewards.append(float(rew[0].item()))
        if c == "a0":
            ref = [-44.790496826171875, -44.388118743896484, -43.990455627441406]
            assert np.allclose(rewards[:3], ref, rtol=2.5e-7)
        b.close()


@pytest.mark.parametrize("kind,kw,amp", [
    ("lorenz3", {}, 0.05), ("lorenz3_pair", {}, 0.05), ("lorenz4_pair", {}, 1.0),
    ("hr_sync", {}, 1.0), ("hr_sync", {"add_filter": True, "add_noise": True}, 1.0),
    ("pmsm_sync", {"alpha": 0.25}, 1.0), ("pmsm_sync", {"alpha": 0.5, "add_noise": True}, 1.0),
    ("pmsm_classic", {}, 2.0), ("pmsm_single", {}, 0.5), ("memristive4_pair", {}, 2.0), ("pmsm_free", {}, 0.0),
])
def test_gpu_vs_oracle_seeded_batch_with_autoreset(oracle_api, kind, kw, amp):
    """4096 envs x 40 steps free-running, Philox resets + Philox noise on both sides, TimeLimit
    12 so that auto-reset, terminal obs and episode stats are exercised."""
    import torch
    O = oracle_api
    n, T, lim = 4096, 40, 12
    b = H.gpu_batch(kind, n, seed=42, autoreset=True, max_episode_steps=lim, env_id_base=1000,
                    **H.flags_from_kwargs(None, kw))
    fl = O.F_AUTORESET | (O.F_ADD_NOISE if kw.get("add_noise") else 0) | (O.F_ADD_FILTER if kw.get("add_filter") else 0)
    o = O.Oracle(kind, n, flags=fl, max_episode_steps=lim, seed=42, env_id_base=1000, alpha=kw.get("alpha", 0.5))
    obs_g = b.reset().double().cpu().numpy()
    obs_o = o.reset()[:, :n].T
    assert np.array_equal(b.state.double().cpu().numpy()[:, :n], o.state[:, :n].astype(np.float64)), "reset state"
    H.assert_close(obs_g, obs_o, 6e-8 if b.real == torch.float64 else 0.0, "reset obs", atol=1e-30)
    rng = np.random.default_rng(3)
    rt = 1e-9 if kind != "pmsm_sync" else 1e-4   # free-running 40 steps incl. libm-vs-device normals
    for t in range(T):
        a = rng.uniform(-amp, amp, (n, b.act_dim)).astype(np.float32)
        obs, rew, done = b.step(torch.as_tensor(a, device=b.device))
        ao = np.zeros((o.act_dim, o.n_pad), np.float32); ao[:, :n] = a.T
        oo, ro, do, extra = o.step(ao)
        assert np.array_equal(done.cpu().numpy(), do[:n]), f"done @ {t}"
        H.assert_close(b.state.double().cpu().numpy()[:, :n], o.state[:, :n].astype(np.float64), rt, f"state @ {t}", atol=1e-12)
        H.assert_close(rew.double().cpu().numpy(), ro[:n], rt, f"reward @ {t}", atol=1e-12)
        if do[:n].any():
            m = do[:n] != 0
            tg = b.terminal_obs().double().cpu().numpy()[m]
            H.assert_close(tg, extra["term_obs"][0][:, :n].T[m], max(rt, 6e-8), "terminal obs", atol=1e-9)
            assert np.array_equal(b.last_ep_len.cpu().numpy()[:n][m], extra["last_ep_len"][:n][m])
            H.assert_close(b.last_ep_ret.cpu().numpy()[:n][m], extra["last_ep_ret"][:n][m], rt, "episode return", atol=1e-9)
    sg = b.stats()
    for k, name in enumerate(("episodes", "return_sum", "return_sq_sum", "length_sum", "nonfinite_events",
                              "terminated", "truncated")):
        assert np.isclose(sg[name], o.stats[k], rtol=1e-6, atol=1e-9), (name, sg[name], o.stats[k])
    assert sg["episodes"] >= 3 * n
    if kind == "lorenz3":
        assert sg["episodes"] == 3 * n and sg["length_sum"] == 3 * n * lim and sg["truncated"] == 3 * n
    b.close()


def test_full_size_65536_lorenz_parity_properties(oracle_api):
    """BASELINE configs[1] size: (a) the oracle on all 65,536 envs for 5 steps, bit-exact;
    (b) size-independent properties: determinism, invariance to how the batch is split into
    slabs (env_id_base), and sum-of-rewards == episode return."""
    import torch
    O = oracle_api
    n = 65536
    b = H.gpu_batch("lorenz3", n, seed=7, autoreset=True, max_episode_steps=1000)
    o = O.Oracle("lorenz3", n, flags=O.F_AUTORESET, max_episode_steps=1000, seed=7)
    b.reset(); o.reset()
    rng = np.random.default_rng(0)
    ret = np.zeros(n)
    for t in range(5):
        a = rng.uniform(-0.05, 0.05, (n, 3)).astype(np.float32)
        obs, rew, done = b.step(torch.as_tensor(a, device=b.device))
        oo, ro, do, _ = o.step(np.ascontiguousarray(a.T))
        assert np.array_equal(b.state.cpu().numpy(), o.state)
        assert np.array_equal(rew.cpu().numpy(), ro)
        ret += ro
    assert np.array_equal(b.ep_return.cpu().numpy(), ret)
    # slab invariance: two half-size slabs with env_id_base 0 / n/2 reproduce the full batch
    halves = []
    for r in range(2):
        h = H.gpu_batch("lorenz3", n // 2, seed=7, autoreset=True, max_episode_steps=1000, env_id_base=r * n // 2)
        h.reset()
        halves.append(h.state.cpu().numpy())
        h.close()
    full = H.gpu_batch("lorenz3", n, seed=7, autoreset=True, max_episode_steps=1000)
    full.reset()
    assert np.array_equal(np.concatenate(halves, axis=1), full.state.cpu().numpy())
    full.close(); b.close()


def test_hr_rk4_full_size_vs_oracle(oracle_api):
    import torch
    O = oracle_api
    n = 65536
    b = H.gpu_batch("hr_sync", n, seed=9)
    o = O.Oracle("hr_sync", n, flags=O.F_AUTORESET, max_episode_steps=5000, seed=9)
    b.reset(); o.reset()
    rng = np.random.default_rng(1)
    for t in range(3):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        obs, rew, done = b.step(torch.as_tensor(a, device=b.device))
        oo, ro, do, _ = o.step(np.ascontiguousarray(a.T))
        H.assert_close(b.state.cpu().numpy(), o.state, 1e-12, "hr state", atol=1e-15)
        H.assert_close(rew.cpu().numpy(), ro, 1e-12, "hr reward")
        assert np.array_equal(done.cpu().numpy(), do)
    # bit-exact fraction should be overwhelming (x**3 via FMA-corrected cube vs libm pow)
    frac = np.mean(b.state.cpu().numpy()[:6] == o.state[:6])
    assert frac > 0.99, frac
    b.close()


@pytest.mark.parametrize("n", [1, 31, 32, 33, 127, 129, 1000])
@pytest.mark.parametrize("kind,amp", [("lorenz3", 0.05), ("hr_sync", 1.0), ("pmsm_sync", 1.0)])
def test_ragged_and_tiny_batches_vs_oracle(oracle_api, kind, amp, n):
    """Edge sizes: single env, partial warps, one-past-a-block; padding lanes stay untouched."""
    import torch
    O = oracle_api
    b = H.gpu_batch(kind, n, seed=3, autoreset=True, max_episode_steps=2)
    o = O.Oracle(kind, n, flags=O.F_AUTORESET, max_episode_steps=2, seed=3)
    b.reset(); o.reset()
    rng = np.random.default_rng(n)
    rt = H.rtol_for(kind)
    for t in range(5):
        a = rng.uniform(-amp, amp, (n, b.act_dim)).astype(np.float32)
        obs, rew, done = b.step(torch.as_tensor(a, device=b.device))
        ao = np.zeros((o.act_dim, o.n_pad), np.float32); ao[:, :n] = a.T
        oo, ro, do, _ = o.step(ao)
        H.assert_close(b.state.double().cpu().numpy()[:, :n], o.state[:, :n].astype(np.float64), rt, f"state n={n}", atol=1e-300)
        H.assert_close(rew.double().cpu().numpy(), ro[:n], rt, "reward")
        assert np.array_equal(done.cpu().numpy(), do[:n])
    assert float(b.state[:, n:].abs().sum().item()) == 0.0 and int(b.ep_len[n:].sum().item()) == 0
    b.close()


def test_nonfinite_and_out_of_range_actions_follow_np_clip(oracle_api):
    """np.clip propagates NaN and saturates +-inf (dynamic.py:63-65, lorenz_env_try.py:92-93)."""
    import torch
    O = oracle_api
    specials = np.array([np.nan, np.inf, -np.inf, 1e30, -1e30, 0.0, -0.0, 499.99997, 500.00003], np.float32)
    for kind in ("lorenz3", "hr_sync", "pmsm_sync", "lorenz4_pair"):
        n = len(specials)
        b = H.gpu_batch(kind, n, seed=1, autoreset=False, max_episode_steps=0)
        o = O.Oracle(kind, n, seed=1)
        b.reset(); o.reset()
        a = np.zeros((n, b.act_dim), np.float32); a[:, 0] = specials; a[:, -1] = specials[::-1]
        obs, rew, done = b.step(torch.as_tensor(a, device=b.device))
        ao = np.zeros((o.act_dim, o.n_pad), np.float32); ao[:, :n] = a.T
        with np.errstate(all="ignore"):
            oo, ro, do, _ = o.step(ao)
        sg, so = b.state.double().cpu().numpy()[:, :n], o.state[:, :n].astype(np.float64)
        assert H.same_nonfinite(sg, so), kind
        H.assert_close(sg, so, H.rtol_for(kind), kind)
        assert H.same_nonfinite(rew.double().cpu().numpy(), ro[:n])
        assert np.array_equal(done.cpu().numpy(), do[:n])
        b.close()


def test_invalid_arguments_raise():
    import torch
    from gym_lorenz_b200 import ChaosLibError
    with pytest.raises(ChaosLibError):
        H.gpu_batch("lorenz3", 0)
    with pytest.raises(KeyError):
        H.gpu_batch("no_such_env", 4)
    b = H.gpu_batch("lorenz3", 8)
    b.reset()
    with pytest.raises(ValueError):
        b.step(torch.zeros((8, 2), device=b.device))
    with pytest.raises(ValueError):
        b.rollout(4, torch.zeros((4, 7, 3), device=b.device))
    b.close()


def test_maximum_size_8M_envs_slab_invariance():
    """BASELINE configs[3] total size on ONE GPU (8 Mi envs, f32 RK4): the batch equals the
    concatenation of 8 x 1 Mi-env slabs keyed by env_id_base (GPU-count invariance)."""
    import torch
    n, parts = 8 * 1048576, 8
    full = H.gpu_batch("lorenz_rk4_f32", n, seed=11, substeps=4)
    full.reset()
    full.rollout(3, want=("reward",))
    ref_state = full.state[:3].clone()
    ref_ret = full.ep_return.clone()
    full.close()
    for r in (0, 5, 7):
        m = n // parts
        s = H.gpu_batch("lorenz_rk4_f32", m, seed=11, substeps=4, env_id_base=r * m)
        s.reset()
        s.rollout(3, want=("reward",))
        assert torch.equal(s.state[:3, :m], ref_state[:, r * m:(r + 1) * m])
        assert torch.equal(s.ep_return[:m], ref_ret[r * m:(r + 1) * m])
        s.close()


def test_divergence_events_are_counted_once(oracle_api):
    """+-500 impulse actions blow dynamic.py's Lorenz up: NaN/inf propagate silently as
    in the reference; stats[4] counts each env's finite -> non-finite transition exactly once."""
    import torch
    O = oracle_api
    n, T = 4096, 400
    b = H.gpu_batch("lorenz3", n, seed=3, autoreset=True, max_episode_steps=0)
    o = O.Oracle("lorenz3", n, flags=O.F_AUTORESET, seed=3)
    b.reset(); o.reset()
    g = torch.Generator(device="cpu").manual_seed(0)
    acts = ((torch.rand((T, n, 3), generator=g) * 2 - 1) * 500).to(b.device)
    b.rollout(T, acts, want=("reward",))
    ao = np.ascontiguousarray(np.pad(acts.cpu().numpy().transpose(0, 2, 1), ((0, 0), (0, 0), (0, o.n_pad - n))))
    with np.errstate(all="ignore"):
        o.rollout(T, ao)
    st = b.stats()
    nonfinite_now = int((~torch.isfinite(b.state[:3, :n]).all(0)).sum().item())
    assert st["nonfinite_events"] == o.stats[4] == nonfinite_now > 0
    assert H.same_nonfinite(b.state.cpu().numpy()[:3, :n], o.state[:3, :n])
    b.close()
