"""North-star kinds have no reference class: the RK4 x S scheme (oracle restatement) is pinned
against adaptive integration (scipy DOP853, rtol=atol=1e-13) per control interval at the
tolerance BASELINE.json states (rtol 1e-9), with zero-order-hold control."""
import numpy as np
import pytest
from scipy.integrate import solve_ivp


def lorenz(t, s, u, q=(10.0, 28.0, 8.0 / 3.0)):
    x, y, z = s
    return [q[0] * (y - x) + u[0], x * (q[1] - z) - y + u[1], x * y - q[2] * z + u[2]]


def pmsm(t, s, u, q=(5.46, 20.0)):
    x, y, z = s
    return [-x + y * z + u[0], -y - x * z + q[1] * z + u[1], q[0] * (y - z)]


def on_attractor(f, n, rng, u0=(0, 0, 0)):
    out = []
    for _ in range(n):
        s0 = rng.uniform(-10, 10, 3)
        sol = solve_ivp(f, (0, 5.0), s0, args=(u0,), method="DOP853", rtol=1e-10, atol=1e-10)
        out.append(sol.y[:, -1])
    return np.array(out)


@pytest.mark.parametrize("S,gain,tol", [(16, 0.0, 1e-9), (16, 50.0, 1e-9), (32, 50.0, 1e-10)])
def test_lorenz_rk4_substeps_vs_dop853(oracle_api, S, gain, tol):
    O = oracle_api
    rng = np.random.default_rng(7)
    n = 24
    st = on_attractor(lorenz, n, rng)
    orc = O.Oracle("lorenz_rk4", n, substeps=S, dt=0.01, act_limit=1.0, act_gain=gain)
    orc.state[:3, :n] = st.T
    a = np.zeros((3, orc.n_pad), np.float32)
    a[:, :n] = rng.uniform(-1, 1, (3, n)).astype(np.float32)
    orc.step(a)
    worst = 0.0
    for i in range(n):
        u = a[:, i].astype(np.float64) * gain
        sol = solve_ivp(lorenz, (0, 0.01), st[i], args=(u,), method="DOP853", rtol=1e-13, atol=1e-13)
        ref = sol.y[:, -1]
        worst = max(worst, float(np.max(np.abs(orc.state[:3, i] - ref) / np.maximum(np.abs(ref), 1.0))))
    assert worst < tol, worst


def test_pmsm_rk4_substeps_vs_dop853(oracle_api):
    O = oracle_api
    rng = np.random.default_rng(8)
    n = 16
    st = on_attractor(pmsm, n, rng)
    orc = O.Oracle("pmsm_rk4", n, substeps=4, dt=0.001, act_limit=1.0, act_gain=50.0, alpha=0.5)
    orc.state[:3, :n] = st.T
    orc.state[3:6, :n] = st.T + rng.uniform(-1, 1, (3, n))
    s2 = orc.state[3:6, :n].copy()
    a = np.zeros((2, orc.n_pad), np.float32)
    a[:, :n] = rng.uniform(-1, 1, (2, n)).astype(np.float32)
    orc.step(a)
    worst = 0.0
    for i in range(n):
        u = [float(a[0, i]) * 50.0, float(a[1, i]) * 50.0, 0.0]
        ref1 = solve_ivp(pmsm, (0, 0.001), st[i], args=((0, 0, 0),), method="DOP853", rtol=1e-13, atol=1e-13).y[:, -1]
        ref2 = solve_ivp(pmsm, (0, 0.001), s2[:, i], args=(u,), method="DOP853", rtol=1e-13, atol=1e-13).y[:, -1]
        worst = max(worst, float(np.max(np.abs(orc.state[:3, i] - ref1) / np.maximum(np.abs(ref1), 1.0))),
                    float(np.max(np.abs(orc.state[3:6, i] - ref2) / np.maximum(np.abs(ref2), 1.0))))
    assert worst < 1e-9, worst


def test_euler_reference_scheme_is_not_rk4_accurate(oracle_api):
    """Documents SURVEY D1: the reference's Euler x 1 is ~1e-3 off the true flow, which is why
    parity kinds reproduce the scheme, not the ODE."""
    O = oracle_api
    rng = np.random.default_rng(9)
    st = on_attractor(lorenz, 4, rng)
    orc = O.Oracle("lorenz3", 4)
    orc.state[:3, :4] = st.T
    orc.step(np.zeros((3, orc.n_pad), np.float32))
    err = 0.0
    for i in range(4):
        ref = solve_ivp(lorenz, (0, 0.01), st[i], args=((0, 0, 0),), method="DOP853", rtol=1e-13, atol=1e-13).y[:, -1]
        err = max(err, float(np.max(np.abs(orc.state[:3, i] - ref))))
    assert 1e-5 < err < 1.0
