"""North-star kinds on the GPU: RK4 x S vs the oracle (1e-12 per interval), vs scipy DOP853
(1e-9, the tolerance BASELINE.json states), fused rollout == repeated single steps, f32
variant, per-env parameter randomisation."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def test_lorenz_rk4_per_interval_vs_oracle(oracle_api):
    import torch
    O = oracle_api
    n = 8192
    for S in (1, 4, 16):
        b = H.gpu_batch("lorenz_rk4", n, seed=5, substeps=S, autoreset=False, max_episode_steps=0)
        o = O.Oracle("lorenz_rk4", n, seed=5, substeps=S, dt=0.01, act_limit=1.0, act_gain=50.0)
        b.reset(); o.reset()
        assert np.array_equal(b.state.cpu().numpy(), o.state)
        rng = np.random.default_rng(S)
        for t in range(4):
            o.state[...] = b.state.cpu().numpy()  # teacher-forced: same pre-state every interval
            a = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
            obs, rew, done = b.step(torch.as_tensor(a, device=b.device))
            oo, ro, do, _ = o.step(np.ascontiguousarray(a.T))
            H.assert_close(b.state.cpu().numpy()[:3], o.state[:3], 1e-12, f"S={S} state", atol=1e-13)
            H.assert_close(rew.cpu().numpy(), ro, 1e-12, "reward", atol=1e-13)
            assert np.array_equal(done.cpu().numpy(), do)
        b.close()


def test_lorenz_rk4_vs_scipy_dop853():
    import torch
    from scipy.integrate import solve_ivp
    n = 256
    b = H.gpu_batch("lorenz_rk4", n, seed=3, substeps=16, autoreset=False, max_episode_steps=0)
    b.reset()
    # move onto the attractor first (200 uncontrolled intervals)
    z = torch.zeros((n, 3), device=b.device)
    for _ in range(200):
        b.step(z)
    st = b.state.cpu().numpy()[:3, :n].T.copy()
    a = np.random.default_rng(0).uniform(-1, 1, (n, 3)).astype(np.float32)
    b.step(torch.as_tensor(a, device=b.device))
    got = b.state.cpu().numpy()[:3, :n].T

    def f(t, s, u):
        x, y, zz = s
        return [10 * (y - x) + u[0], x * (28 - zz) - y + u[1], x * y - (8 / 3) * zz + u[2]]
    worst = 0.0
    for i in range(n):
        ref = solve_ivp(f, (0, 0.01), st[i], args=(a[i].astype(np.float64) * 50.0,), method="DOP853",
                        rtol=1e-13, atol=1e-13).y[:, -1]
        worst = max(worst, float(np.max(np.abs(got[i] - ref) / np.maximum(np.abs(ref), 1.0))))
    assert worst < 1e-9, worst
    b.close()


@pytest.mark.parametrize("kind", ["lorenz_rk4", "lorenz3", "hr_sync", "pmsm_sync", "pmsm_rk4"])
def test_fused_rollout_equals_single_steps_bit_exactly(kind):
    import torch
    n, T = 3000, 17
    kw = dict(seed=11, autoreset=True, max_episode_steps=7)
    b1 = H.gpu_batch(kind, n, **kw)
    b2 = H.gpu_batch(kind, n, **kw)
    b1.reset(); b2.reset()
    g = torch.Generator(device="cpu").manual_seed(0)
    amp = 0.05 if kind == "lorenz3" else float(b1.layout.act_high)  # +-500 impulses overflow to NaN
    acts = (torch.rand((T, n, b1.act_dim), generator=g) * 2 - 1).to(b1.device) * amp
    out = b1.rollout(T, acts)
    for t in range(T):
        obs, rew, done = b2.step(acts[t])
        assert torch.equal(out["obs"][t, :, :n].t(), obs), (kind, t)
        assert torch.equal(out["reward"][t, :n], rew)
        assert torch.equal(out["done"][t, :n], done)
    assert torch.equal(b1.state, b2.state) and torch.equal(b1.ep_len, b2.ep_len)
    assert torch.equal(b1.ep_return, b2.ep_return)
    assert b1.step_index == b2.step_index
    b1.close(); b2.close()


def test_synthetic_action_rollout_vs_oracle(oracle_api):
    O = oracle_api
    n, T = 4096, 6
    b = H.gpu_batch("lorenz_rk4", n, seed=21, substeps=4, env_id_base=77)
    o = O.Oracle("lorenz_rk4", n, flags=O.F_AUTORESET, seed=21, substeps=4, dt=0.01, act_limit=1.0,
                 act_gain=50.0, max_episode_steps=1000, env_id_base=77)
    b.reset(); o.reset()
    out = b.rollout(T)
    ref = o.rollout(T, None, synth_amp=1.0)
    H.assert_close(out["obs"].double().cpu().numpy(), ref["obs"], 1e-6, "obs (f32 store)", atol=1e-6)
    H.assert_close(out["reward"].cpu().numpy(), ref["reward"], 1e-11, "reward", atol=1e-12)
    H.assert_close(b.state.cpu().numpy()[:3], o.state[:3], 1e-11, "state", atol=1e-12)
    b.close()


def test_lorenz_rk4_f32_vs_oracle(oracle_api):
    import torch
    O = oracle_api
    n = 4096
    b = H.gpu_batch("lorenz_rk4_f32", n, seed=5, substeps=8, autoreset=False, max_episode_steps=0)
    o = O.Oracle("lorenz_rk4_f32", n, seed=5, substeps=8, dt=0.01, act_limit=1.0, act_gain=50.0)
    b.reset(); o.reset()
    assert np.array_equal(b.state.cpu().numpy(), o.state)
    a = np.random.default_rng(0).uniform(-1, 1, (n, 3)).astype(np.float32)
    b.step(torch.as_tensor(a, device=b.device))
    o.step(np.ascontiguousarray(a.T))
    # float32 round-off only (the oracle integrates the same scheme in float32, in the textbook form; the
    # kernel on the shifted coordinate z - rho): a few ulp(32) = 3.8e-6 per substep on states of scale 30
    H.assert_close(b.state.cpu().numpy()[:3], o.state[:3], 2e-5, "f32 rk4 state", atol=4e-5)
    b.close()


def test_per_env_parameter_randomisation(oracle_api):
    O = oracle_api
    n = 4096
    b = H.gpu_batch("pmsm_rk4", n, seed=0, param_jitter=0.1)
    o = O.Oracle("pmsm_rk4", n, seed=0, param_jitter=0.1, substeps=4, dt=0.001, act_gain=50.0)
    sg = b.state.cpu().numpy()
    assert np.array_equal(sg[6:8], o.state[6:8])
    assert 5.46 * 0.9 <= sg[6, :n].min() and sg[6, :n].max() <= 5.46 * 1.1
    assert 20 * 0.9 <= sg[7, :n].min() and sg[7, :n].max() <= 20 * 1.1
    assert sg[6, :n].std() > 0.1
    b.reset()
    assert np.array_equal(b.state.cpu().numpy()[6:8], sg[6:8])  # parameters survive reset
    b.close()


@pytest.mark.parametrize("kind,n,layout", [
    ("lorenz_rk4", 65536, "soa"),      # bench shape: bulk-copy (TMA) action staging
    ("lorenz_rk4", 40000, "soa"),      # partial last warp, padded planes
    ("lorenz_rk4", 40000, "aos"),      # policy-shaped [T, N, A] actions: LDG path
    ("lorenz_rk4", 40000, "synth"),    # in-kernel Philox actions
    ("hr_sync", 33333, "aos"),
    ("pmsm_sync", 35000, "soa"),
    ("lorenz3", 34567, "soa"),
])
def test_dynamic_rollout_is_bit_identical_to_static(kind, n, layout, monkeypatch):
    """k_rollout_dyn (env-warp x interval-chunk tasks pulled from an atomic queue, chunks of one
    env-warp handed between SMs through release/acquire) must reproduce the static kernel."""
    import torch
    T = 37   # not a multiple of the chunk (8): exercises the short last chunk
    kw = dict(seed=13, autoreset=True, max_episode_steps=11)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CHAOS_B200_DYN", mode)
        b = H.gpu_batch(kind, n, **kw)
        b.reset()
        g = torch.Generator(device="cpu").manual_seed(5)
        amp = 0.05 if kind == "lorenz3" else float(b.layout.act_high)
        if layout == "soa":
            soa = ((torch.rand((T, b.act_dim, b.n_pad), generator=g) * 2 - 1) * amp).to(b.device)
            acts = soa[:, :, :n].permute(0, 2, 1)
        elif layout == "aos":
            acts = ((torch.rand((T, n, b.act_dim), generator=g) * 2 - 1) * amp).to(b.device)
        else:
            acts = None
        out = b.rollout(T, acts)
        torch.cuda.synchronize()
        assert b.dyn_launch_count == (1 if mode == "1" else 0)
        res[mode] = (out["obs"].clone(), out["reward"].clone(), out["done"].clone(), b.state.clone(),
                     b.ep_len.clone(), b.ep_return.clone(), b.stats())
        b.close()
    for x, y in zip(res["0"][:6], res["1"][:6]):
        assert torch.equal(torch.nan_to_num(x[..., :n].double()), torch.nan_to_num(y[..., :n].double()))
    s0, s1 = res["0"][6], res["1"][6]
    assert s0["episodes"] == s1["episodes"] and s0["length_sum"] == s1["length_sum"]
    assert np.isclose(s0["return_sum"], s1["return_sum"], rtol=1e-9)


def test_dynamic_rollout_is_selected_automatically_at_65536(monkeypatch):
    monkeypatch.delenv("CHAOS_B200_DYN", raising=False)
    b = H.gpu_batch("lorenz_rk4", 65536, seed=1)
    b.reset()
    b.rollout(64, want=("reward",))
    assert b.dyn_launch_count == 1          # 2048 env-warps on 592 schedulers: 86.5 % static balance
    b2 = H.gpu_batch("lorenz_rk4", 1048576, seed=1)
    b2.reset()
    b2.rollout(16, want=("reward",))
    assert b2.dyn_launch_count == 0         # 32768 env-warps: 98.8 % static balance
    b.close(); b2.close()


@pytest.mark.parametrize("kind", ["pmsm_sync", "hr_sync", "lorenz_rk4"])
def test_dynamic_rollout_many_handoffs_stress(kind, monkeypatch):
    """26 chunks per env-warp, repeated: every chunk boundary hands an env-warp's planes (and,
    for PMSM, its int32 Adam counter) from one SM to another through L2."""
    import torch
    n, T = 70000, 203
    ref = None
    for rep, mode in enumerate(("0", "1", "1", "1")):
        monkeypatch.setenv("CHAOS_B200_DYN", mode)
        b = H.gpu_batch(kind, n, seed=21, autoreset=True, max_episode_steps=50)
        b.reset()
        out = b.rollout(T, None, want=("reward", "done"))      # in-kernel Philox actions
        torch.cuda.synchronize()
        cur = (out["reward"][:, :n].clone(), out["done"][:, :n].clone(), b.state.clone(), b.aux_int.clone(),
               b.ep_len.clone(), b.ep_return.clone())
        if ref is None:
            ref = cur
        else:
            for x, y in zip(ref, cur):
                assert torch.equal(torch.nan_to_num(x.double()), torch.nan_to_num(y.double())), (kind, rep)
        b.close()


@pytest.mark.parametrize("kind,n,T,workers,chunk", [
    ("lorenz_rk4", 65536, 37, None, None),    # bench shape: 12 residents + 1-2 guests per SM, short last chunk
    ("lorenz_rk4", 65536, 64, "16", "1"),     # clamped to 13 residents: 0-1 guests, single-interval chunks
    ("lorenz_rk4", 40000, 29, "12", "3"),     # ragged: partial last warp; SMs own 8 or 9 env-warps (8 residents)
    ("lorenz_rk4", 70001, 33, "8", "8"),      # 8 residents + 6-7 guests: guests nearly alternate with own chunks
    ("lorenz_rk4", 70001, 40, "3", "5"),      # 3 residents + 11-12 guests
    ("lorenz_rk4_f32", 65536, 26, None, "2"),
    ("pmsm_rk4", 50000, 21, None, None),
    ("lorenz_rk4", 3000, 40, None, None),     # fewer env-warps (94) than SMs: one env-warp per block, no guests
])
def test_sm_local_rollout_is_bit_identical_to_static(kind, n, T, workers, chunk, monkeypatch):
    """k_rollout_sm (per SM: resident env-warps in registers for the whole launch, guest env-warps chunked
    through a shared-memory queue, actions by tensor copies) must reproduce the static one-thread-per-env
    kernel bit for bit, with episodes ending (TimeLimit 11) inside the window, with and without the tensor
    map, and so must the global-queue kernel it replaces."""
    import torch
    kw = dict(seed=13, autoreset=True, max_episode_steps=11)
    res = {}
    for mode in ("static", "sm", "sm_rows", "dyn"):
        monkeypatch.setenv("CHAOS_B200_DYN", "0" if mode == "static" else "1")
        monkeypatch.setenv("CHAOS_B200_SM", "0" if mode == "dyn" else "1")
        monkeypatch.setenv("CHAOS_B200_SM_TMAP", "0" if mode == "sm_rows" else "1")
        for k, v in (("CHAOS_B200_SM_WORKERS", workers), ("CHAOS_B200_SM_CHUNK", chunk)):
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, v)
        b = H.gpu_batch(kind, n, **kw)
        b.reset()
        g = torch.Generator(device="cpu").manual_seed(5)
        soa = ((torch.rand((T, b.act_dim, b.n_pad), generator=g) * 2 - 1) * float(b.layout.act_high)).to(b.device)
        out = b.rollout(T, soa[:, :, :n].permute(0, 2, 1))
        torch.cuda.synchronize()
        assert b.sm_launch_count == (1 if mode.startswith("sm") else 0), mode
        assert b.dyn_launch_count == (0 if mode == "static" else 1), mode
        res[mode] = (out["obs"].clone(), out["reward"].clone(), out["done"].clone(), b.state.clone(),
                     b.ep_len.clone(), b.ep_return.clone(), b.stats())
        b.close()
    for other in ("sm", "sm_rows", "dyn"):
        for x, y in zip(res["static"][:6], res[other][:6]):
            assert torch.equal(torch.nan_to_num(x[..., :n].double()), torch.nan_to_num(y[..., :n].double())), other
        s0, s1 = res["static"][6], res[other][6]
        for key in ("episodes", "length_sum", "terminated", "truncated", "nonfinite_events"):
            assert s0[key] == s1[key], (other, key)
        assert s0["episodes"] > 0
        assert np.isclose(s0["return_sum"], s1["return_sum"], rtol=1e-9)


@pytest.mark.parametrize("gdiv", ["2", "4", "16"])
def test_sm_local_rollout_guest_chunk_divisor(gdiv, monkeypatch):
    """Guests may use shorter chunks than residents (CHAOS_B200_SM_GDIV): same bits as the static kernel."""
    import torch
    n, T = 65536, 53
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CHAOS_B200_DYN", mode)
        monkeypatch.setenv("CHAOS_B200_SM_GDIV", gdiv)
        b = H.gpu_batch("lorenz_rk4", n, seed=17, autoreset=True, max_episode_steps=13)
        b.reset()
        g = torch.Generator(device="cpu").manual_seed(8)
        soa = (torch.rand((T, b.act_dim, b.n_pad), generator=g) * 2 - 1).to(b.device)
        out = b.rollout(T, soa[:, :, :n].permute(0, 2, 1))
        torch.cuda.synchronize()
        assert b.sm_launch_count == (1 if mode == "1" else 0)
        res[mode] = (out["obs"].clone(), out["reward"].clone(), out["done"].clone(), b.state.clone(), b.ep_len.clone(),
                     b.ep_return.clone())
        b.close()
    for x, y in zip(res["0"], res["1"]):
        assert torch.equal(x, y)


def test_sm_local_rollout_repeated_launches_and_jitter(monkeypatch):
    """Back-to-back launches (state leaves and re-enters shared memory every launch) with per-env
    parameters (the generic, register-parameter interval loop) against the static kernel."""
    import torch
    n, T = 65536, 24
    ref = None
    for mode in ("0", "1"):
        monkeypatch.setenv("CHAOS_B200_DYN", mode)
        b = H.gpu_batch("lorenz_rk4", n, seed=3, autoreset=True, max_episode_steps=17, param_jitter=0.1)
        b.reset()
        g = torch.Generator(device="cpu").manual_seed(2)
        soa = (torch.rand((T, b.act_dim, b.n_pad), generator=g) * 2 - 1).to(b.device)
        outs = []
        for rep in range(3):
            o = b.rollout(T, soa[:, :, :n].permute(0, 2, 1))
            outs.append((o["obs"].clone(), o["reward"].clone(), o["done"].clone()))
        torch.cuda.synchronize()
        assert b.sm_launch_count == (3 if mode == "1" else 0)
        cur = (outs, b.state.clone(), b.ep_len.clone(), b.ep_return.clone())
        if ref is None:
            ref = cur
        else:
            for (a0, a1, a2), (c0, c1, c2) in zip(ref[0], cur[0]):
                assert torch.equal(a0, c0) and torch.equal(a1, c1) and torch.equal(a2, c2)
            assert torch.equal(ref[1], cur[1]) and torch.equal(ref[2], cur[2]) and torch.equal(ref[3], cur[3])
        b.close()


def test_bench_configuration_against_the_oracle_directly(oracle_api, monkeypatch):
    """BASELINE configs[1] on the instantiation bench.py times: Lorenz RK4 x 16, FP64, 65,536 envs,
    given actions, plain rollout shape -> k_rollout_sm.  State, reward and done flags of a 4,096-env
    slab (same global env ids) against oracle.rollout, T = 32 control intervals with TimeLimit resets
    inside the window."""
    import torch
    O = oracle_api
    for k in ("CHAOS_B200_DYN", "CHAOS_B200_SM", "CHAOS_B200_SM_WORKERS", "CHAOS_B200_SM_CHUNK", "CHAOS_B200_PLAIN"):
        monkeypatch.delenv(k, raising=False)
    n, T, S = 65536, 32, 16
    kw = dict(seed=7, autoreset=True, max_episode_steps=20, substeps=S)
    b = H.gpu_batch("lorenz_rk4", n, **kw)
    b.reset()
    g = torch.Generator(device="cpu").manual_seed(11)
    soa = (torch.rand((T, b.act_dim, b.n_pad), generator=g) * 2 - 1).to(b.device)
    out = b.rollout(T, soa[:, :, :n].permute(0, 2, 1))
    torch.cuda.synchronize()
    assert b.dyn_launch_count == 1 and b.plain_launch_count == 1 and b.sm_launch_count == 1
    for lo, m in ((0, 2048), (41_000, 4096), (65536 - 1024, 1024)):
        o = O.Oracle("lorenz_rk4", m, flags=O.F_AUTORESET, seed=7, substeps=S, dt=0.01, act_limit=1.0, act_gain=50.0,
                     max_episode_steps=20, env_id_base=lo)
        o.reset()
        a_np = np.zeros((T, o.act_dim, o.n_pad), np.float32)
        a_np[:, :, :m] = soa[:, :, lo:lo + m].cpu().numpy()
        ref = o.rollout(T, a_np)
        H.assert_close(out["reward"][:, lo:lo + m].cpu().numpy(), ref["reward"][:, :m], 1e-11, "reward vs oracle", atol=1e-11)
        assert np.array_equal(out["done"][:, lo:lo + m].cpu().numpy(), ref["done"][:, :m])
        H.assert_close(out["obs"][:, :, lo:lo + m].double().cpu().numpy(), ref["obs"][:, :, :m], 1e-6, "obs (f32)", atol=1e-5)
        H.assert_close(b.state[:3, lo:lo + m].cpu().numpy(), o.state[:3, :m], 1e-11, "state vs oracle", atol=1e-11)
        assert np.array_equal(b.ep_len[lo:lo + m].cpu().numpy(), o.ep_len[:m])
    b.close()


def test_pmsm_rk4_with_parameter_jitter_vs_oracle_and_scipy(oracle_api):
    """BASELINE configs[2]: chaotic PMSM pair, 65,536 envs, FP64, per-env sigma/gamma ~ U(0.9,1.1) x
    nominal.  Per control interval vs the oracle (1e-12) and, for a sample, vs DOP853 (1e-9)."""
    import torch
    from scipy.integrate import solve_ivp
    O = oracle_api
    n = 65536
    b = H.gpu_batch("pmsm_rk4", n, seed=17, param_jitter=0.1, substeps=4, autoreset=False, max_episode_steps=0)
    o = O.Oracle("pmsm_rk4", n, seed=17, param_jitter=0.1, substeps=4, dt=0.001, act_limit=1.0, act_gain=50.0, alpha=0.5)
    b.reset(); o.reset()
    assert np.array_equal(b.state.cpu().numpy(), o.state)
    rng = np.random.default_rng(2)
    for t in range(3):
        o.state[...] = b.state.cpu().numpy()
        pre = o.state.copy()
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        obs, rew, done = b.step(torch.as_tensor(a, device=b.device))
        oo, ro, do, _ = o.step(np.ascontiguousarray(a.T))
        H.assert_close(b.state.cpu().numpy()[:6], o.state[:6], 1e-12, "pmsm_rk4 state", atol=1e-13)
        H.assert_close(rew.cpu().numpy(), ro, 1e-12, "pmsm_rk4 reward", atol=1e-12)
        assert np.array_equal(done.cpu().numpy(), do)
    got = b.state.cpu().numpy()

    def f(_t, s, u, q):
        x, y, z = s
        return [-x + y * z + u[0], -y - x * z + q[1] * z + u[1], q[0] * (y - z)]
    worst = 0.0
    for i in range(0, n, n // 16):
        q = pre[6:8, i]
        u = (float(a[i, 0]) * 50.0, float(a[i, 1]) * 50.0)
        r1 = solve_ivp(f, (0, 0.001), pre[0:3, i], args=((0.0, 0.0), q), method="DOP853", rtol=1e-13, atol=1e-13).y[:, -1]
        r2 = solve_ivp(f, (0, 0.001), pre[3:6, i], args=(u, q), method="DOP853", rtol=1e-13, atol=1e-13).y[:, -1]
        worst = max(worst, float(np.max(np.abs(got[0:3, i] - r1) / np.maximum(np.abs(r1), 1.0))),
                    float(np.max(np.abs(got[3:6, i] - r2) / np.maximum(np.abs(r2), 1.0))))
    assert worst < 1e-9, worst
    b.close()


@pytest.mark.parametrize("kind,n", [("lorenz_rk4", 65536), ("hr_sync", 4096), ("lorenz3", 33), ("pmsm_sync", 1001)])
def test_row_major_observation_rollout_equals_plane_layout(kind, n, monkeypatch):
    """obs_layout='rows' ([T, N, obs_dim], warp-transposed 16-byte stores when every row block is
    16-byte aligned, generic stores otherwise -- n=33 and n=1001 make the per-interval stride
    unaligned) must carry exactly the values of the SoA plane layout, static and dynamic kernel."""
    import torch
    T = 9
    for dyn in ("0", "1"):
        monkeypatch.setenv("CHAOS_B200_DYN", dyn)
        outs = []
        for layout in ("planes", "rows"):
            b = H.gpu_batch(kind, n, seed=3, max_episode_steps=4)
            b.reset()
            o = b.rollout(T, None, obs_layout=layout)
            torch.cuda.synchronize()
            outs.append(o["obs"].clone())
            b.close()
        planes, rows = outs
        assert rows.shape == (T, n, planes.shape[1])
        assert torch.equal(torch.nan_to_num(planes[:, :, :n].permute(0, 2, 1)), torch.nan_to_num(rows))


@pytest.mark.parametrize("kind,n,T", [("lorenz_rk4", 65536 + 40, 40), ("lorenz_rk4", 65536 + 40, 19),   # 19: partial last chunk
                                      ("pmsm_rk4", 9000, 19), ("lorenz_rk4", 3000, 21),
                                      ("lorenz_rk4_f32", 65536 + 40, 24), ("lorenz_rk4_f32", 5000, 11)])
def test_plain_rollout_instantiation_equals_generic_bitwise(kind, n, T, monkeypatch):
    """The rollout kernels exist twice for the FP64-bound kinds: generic, and with the plain I/O
    shape as compile-time facts (kernels_common.cuh PlainRollout).  Same inputs -> same bits, on
    the dynamic (65,576 envs: env-warps do not divide over the schedulers) and the static kernel,
    with episodes ending inside the window."""
    import torch
    outs = []
    for plain in ("1", "0"):
        monkeypatch.setenv("CHAOS_B200_PLAIN", plain)
        b = H.gpu_batch(kind, n, seed=9, autoreset=True, max_episode_steps=6, substeps=3 if kind == "pmsm_rk4" else 16)
        b.reset()
        g = torch.Generator(device="cpu").manual_seed(1)
        acts = (torch.rand((T, b.act_dim, b.n_pad), generator=g) * 2 - 1).to(b.device)
        acts[:, :, 5] = 0.0
        o = b.rollout(T, acts[:, :, :n].permute(0, 2, 1))
        outs.append((o["obs"].clone(), o["reward"].clone(), o["done"].clone(), b.state.clone(), b.ep_len.clone(),
                     b.ep_return.clone(), dict(b.stats()), b.dyn_launch_count))
        assert b.plain_launch_count == (1 if plain == "1" else 0)
        b.close()
    a, c = outs
    assert a[7] == c[7]
    if "CHAOS_B200_DYN" not in os.environ:      # automatic choice: dynamic kernel only when env-warps do not divide evenly
        assert (a[7] > 0) == (n > 65536)
    for k in range(6):
        assert torch.equal(a[k][..., :n], c[k][..., :n]), k
    for key in ("episodes", "length_sum", "terminated", "truncated", "nonfinite_events"):
        assert a[6][key] == c[6][key], key
    assert a[6]["episodes"] > 0
    assert np.isclose(a[6]["return_sum"], c[6]["return_sum"], rtol=1e-12)   # atomics: order differs


@pytest.mark.parametrize("kind", ["lorenz_rk4", "lorenz_rk4_f32"])
def test_one_mebi_envs_per_gpu_properties(kind, oracle_api):
    """BASELINE configs[3] size (1,048,576 envs per GPU, FP64 and FP32): size-independent properties.
    (a) sharding invariance: 8 slabs of 131,072 envs with env_id_base = r * 131,072 (what 8 ranks
        own) reproduce the single 1 Mi batch bit for bit -- reset states and a fused rollout with
        auto-resets inside; (b) the fused rollout equals single steps; (c) sum of rewards equals the
        Monitor return of finished episodes; (d) a 4,096-env sample of the big batch equals the
        oracle run on just those envs (same global env ids)."""
    import torch
    O = oracle_api
    n, shards, T = 1 << 20, 8, 6
    kw = dict(seed=13, autoreset=True, max_episode_steps=4, substeps=16)
    full = H.gpu_batch(kind, n, **kw)
    full.reset()
    g = torch.Generator(device="cpu").manual_seed(5)
    acts = (torch.rand((T, full.act_dim, n), generator=g) * 2 - 1).to(full.device)        # SoA, env stride 1
    st0 = full.state.clone()
    out = full.rollout(T, acts.permute(0, 2, 1))
    assert full.plain_launch_count == 1
    per = n // shards
    for r in range(shards):
        sl = slice(r * per, (r + 1) * per)
        h = H.gpu_batch(kind, per, env_id_base=r * per, **kw)
        h.reset()
        assert torch.equal(h.state[:, :per], st0[:, sl])
        if r in (0, 5):
            oh = h.rollout(T, acts[:, :, sl].permute(0, 2, 1))
            assert torch.equal(oh["obs"][:, :, :per], out["obs"][:, :, sl])
            assert torch.equal(oh["reward"][:, :per], out["reward"][:, sl])
            assert torch.equal(oh["done"][:, :per], out["done"][:, sl])
            assert torch.equal(h.state[:, :per], full.state[:, sl])
        h.close()
    # (b) single steps on a fresh batch
    b2 = H.gpu_batch(kind, n, **kw)
    b2.reset()
    ret = torch.zeros(n, dtype=torch.float64, device=full.device)
    for t in range(T):
        obs, rew, done = b2.step(acts[t].t())
        assert torch.equal(rew, out["reward"][t, :n]) and torch.equal(done, out["done"][t, :n])
        ret += rew.double()
        if t == 3:   # (c) every env hits the 4-step TimeLimit here (unless it blew up earlier)
            fin = done != 0
            assert fin.all()
            trunc_only = done == 2
            assert torch.equal(b2.last_ep_ret[:n][trunc_only], ret[trunc_only])
            ret.zero_()
    assert torch.equal(b2.state, full.state)
    # (d) oracle on a sample: global env ids 777,000 .. 781,095
    lo, m = 777_000, 4096
    o = O.Oracle(kind, m, flags=O.F_AUTORESET, seed=13, substeps=16, dt=0.01, act_limit=1.0, act_gain=50.0,
                 max_episode_steps=4, env_id_base=lo)
    o.reset()
    a_np = np.zeros((T, o.act_dim, o.n_pad), np.float32)
    a_np[:, :, :m] = acts[:, :, lo:lo + m].cpu().numpy()
    ref = o.rollout(T, a_np)
    tol = 1e-11 if kind == "lorenz_rk4" else 2e-4
    H.assert_close(out["reward"][:, lo:lo + m].double().cpu().numpy(), ref["reward"][:, :m], tol, "reward vs oracle", atol=tol)
    assert np.array_equal(out["done"][:, lo:lo + m].cpu().numpy(), ref["done"][:, :m])
    H.assert_close(full.state[:3, lo:lo + m].double().cpu().numpy(), o.state[:3, :m], tol, "state vs oracle", atol=tol)
    full.close(); b2.close()


@pytest.mark.parametrize("kind", ["lorenz_rk4", "hr_sync", "lorenz_rk4_f32"])
@pytest.mark.parametrize("layout", ["planes", "rows"])
def test_rollout_outputs_stay_inside_their_buffers(kind, layout):
    """Guard bands around every rollout output (ragged batch: partial last warp, N not a multiple of
    the 128-env padding): the vectorised row stores / plane stores must not touch a byte outside
    [T, ...] -- the check compute-sanitizer would do, written as a test."""
    import torch
    n, T, G = 1000 + 13, 7, 4096
    b = H.gpu_batch(kind, n, seed=3, autoreset=True, max_episode_steps=3)
    b.reset()
    NP, O = b.n_pad, b.obs_dim
    shape = (T, n, O) if layout == "rows" else (T, O, NP)
    numel = int(np.prod(shape))
    raw = {"obs": torch.full((numel + 2 * G,), 12345.0, dtype=torch.float32, device=b.device),
           "reward": torch.full((T * NP + 2 * G,), 12345.0, dtype=b.real, device=b.device),
           "done": torch.full((T * NP + 2 * G,), 77, dtype=torch.uint8, device=b.device)}
    out = {"obs": raw["obs"][G:G + numel].view(shape), "reward": raw["reward"][G:G + T * NP].view(T, NP),
           "done": raw["done"][G:G + T * NP].view(T, NP)}
    g = torch.Generator(device="cpu").manual_seed(1)
    acts = (torch.rand((T, b.act_dim, NP), generator=g) * 2 - 1).to(b.device)
    for dyn in ("0", "1"):
        import os
        os.environ["CHAOS_B200_DYN"] = dyn
        try:
            b.rollout(T, acts[:, :, :n].permute(0, 2, 1), out=out, obs_layout=layout)
        finally:
            os.environ.pop("CHAOS_B200_DYN", None)
        torch.cuda.synchronize()
        for k, fill in (("obs", 12345.0), ("reward", 12345.0), ("done", 77)):
            assert bool((raw[k][:G] == fill).all()) and bool((raw[k][-G:] == fill).all()), (k, dyn)
        if layout == "planes":   # padding lanes n .. n_pad of every plane are never written either
            assert bool((out["obs"][:, :, n:] == 12345.0).all()) and bool((out["done"][:, n:] == 77).all())
        assert bool(torch.isfinite(out["obs"][..., :n] if layout == "planes" else out["obs"]).all())
    b.close()
