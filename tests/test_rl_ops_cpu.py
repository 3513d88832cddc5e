"""CPU checks of the SB3 / metric restatements in oracle/sb3_ref.py against independent
formulations (the device kernels are checked against these in tests/test_gpu_rl_ops.py)."""
import numpy as np

from oracle import sb3_ref as S


def test_gae_matches_closed_form_without_episode_boundaries():
    rng = np.random.default_rng(0)
    T, N, g, lam = 12, 5, 0.99, 0.95
    r, v = rng.normal(size=(T, N)).astype(np.float32), rng.normal(size=(T, N)).astype(np.float32)
    lv = rng.normal(size=N).astype(np.float32)
    adv, ret = S.gae(r, v, np.zeros((T, N), np.float32), lv, np.zeros(N), g, lam)
    vv = np.vstack([v, lv[None]]).astype(np.float64)
    delta = r + g * vv[1:] - vv[:-1]
    ref = np.zeros((T, N))
    for t in range(T):
        ref[t] = sum((g * lam) ** k * delta[t + k] for k in range(T - t))
    assert np.allclose(adv, ref, rtol=2e-5, atol=2e-5) and np.allclose(ret, adv + v)


def test_gae_episode_start_cuts_the_bootstrap():
    r = np.ones((3, 1), np.float32); v = np.zeros((3, 1), np.float32)
    starts = np.array([[0], [0], [1]], np.float32)       # a new episode starts at t=2
    adv, _ = S.gae(r, v, starts, np.array([5.0]), np.array([1.0]), 0.9, 1.0)
    assert np.allclose(adv[:, 0], [1 + 0.9 * 1, 1.0, 1.0])   # t=1 does not see t=2; t=2 does not see V_last


def test_running_mean_std_equals_batch_statistics():
    rng = np.random.default_rng(1)
    x = rng.normal(3.0, 2.0, size=(4000, 6))
    rms = S.RunningMeanStd(shape=(6,))
    for k in range(0, 4000, 500):
        rms.update(x[k:k + 500])
    assert np.allclose(rms.mean, x.mean(0), atol=1e-3) and np.allclose(rms.var, x.var(0), rtol=1e-3)
    z = S.normalize_obs(x.astype(np.float32), rms, clip_obs=1.5)
    assert z.dtype == np.float32 and z.max() <= 1.5 and z.min() >= -1.5


def test_frame_stack_update():
    st = np.arange(12, dtype=np.float32).reshape(2, 6)     # 2 envs, 3 frames of 2
    out = S.frame_stack_update(st, np.array([[100, 101], [200, 201]], np.float32), [False, True])
    assert out[0].tolist() == [2, 3, 4, 5, 100, 101] and out[1].tolist() == [0, 0, 0, 0, 200, 201]


def test_metrics_reference_semantics():
    T = 2400
    t = np.arange(T)
    e = np.stack([np.exp(-t / 100.0), 0.5 * np.exp(-t / 50.0), np.full(T, 0.01)], 1)
    u = np.ones((T, 2))
    mae, rmse, ts, en = S.steady_state_metrics(e, u, dt=0.001)
    assert np.isclose(en, 2 * T * 0.001)
    last = np.max(np.where(np.abs(e[:, 0]) > 0.05)[0]) + 1
    assert np.isclose(ts, last * 0.001) and mae > 0 and rmse >= mae * 0.5
    e[-1, 1] = 1.0                                              # component 1 never settles -> NaN ignored by nanmax
    _, _, ts2, _ = S.steady_state_metrics(e, u, dt=0.001)
    assert np.isclose(ts2, ts)
    e[-1, :] = 1.0
    assert np.isnan(S.steady_state_metrics(e, u, dt=0.001)[2])
