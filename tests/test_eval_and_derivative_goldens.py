"""Reference-pinned fixtures for SURVEY 8f rank 3 (evaluation metrics) and 8a row a6 (right-hand sides).

tests/golden/eval_metrics.npz and derivatives.npz are produced by tests/golden/make_golden.py from the
reference's own code: calculate_advanced_metrics and the steady-state block of
code/lorenz_pmsm/test_evaluate.py (:25-59, :239-250, executed from the file), hr_derivatives
(lorenz_env_try.py:7-12) and the derivative halves of the classic envs' state0 after step().
CPU part: the NumPy restatement (oracle/sb3_ref.py) against the fixtures and -- where the reference
tree exists -- against the reference executed on fresh inputs.  GPU part: cl_eval_metrics and
cl_derivatives through the C-ABI against the fixtures."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import ref_loader as R
from oracle import sb3_ref as S

GOLD = np.load(os.path.join(H.GOLDEN, "eval_metrics.npz"))
DER = np.load(os.path.join(H.GOLDEN, "derivatives.npz"))


def _same(a, b):
    return (np.isnan(a) and np.isnan(b)) or a == b


@pytest.mark.parametrize("tag", ["long", "short"])
def test_restatement_reproduces_the_reference_generated_metrics(tag):
    e, u, dt = GOLD[f"{tag}_err"].astype(np.float64), GOLD[f"{tag}_ctrl"].astype(np.float64), float(GOLD[f"{tag}_dt"])
    ref = GOLD[f"{tag}_metrics"]
    for k in range(e.shape[2]):
        got = S.steady_state_metrics(e[:, :, k], u[:, :, k], dt=dt)
        for c in range(4):
            assert _same(float(got[c]), float(ref[k, c])), (tag, k, c, got[c], ref[k, c])
    assert np.isnan(ref[1, 2]) and ref[2, 2] == 0.0      # the never-settles / never-leaves cases are in the fixture


@pytest.mark.skipif(not R.available(), reason="reference tree not present")
def test_restatement_equals_the_reference_executed_here():
    import sys
    sys.path.insert(0, H.GOLDEN)
    import make_golden as MG
    block = R.eval_steady_block()
    calc = R.load_eval_script().calculate_advanced_metrics
    rng = np.random.default_rng(123)
    for T, dt in ((2500, 0.001), (900, 0.01), (2000, 0.001)):
        e, u = MG.eval_inputs(rng, T, 6)
        for k in range(6):
            ref = block(e[:, 0, k], e[:, 1, k], e[:, 2, k], u[:, 0, k], u[:, 1, k], dt)
            got = S.steady_state_metrics(e[:, :, k], u[:, :, k], dt=dt)
            assert all(_same(float(g), float(r)) for g, r in zip(got, ref)), (T, k, got, ref)
            for c in range(3):
                a = calc(e[:, c, k], u[:, 0, k], u[:, 1, k], dt=dt)
                b = S.calculate_advanced_metrics(e[:, c, k], u[:, 0, k], u[:, 1, k], dt=dt)
                assert _same(float(a[0]), float(b[0])) and a[1] == b[1]


def test_derivative_fixture_is_self_consistent():
    """The harvested (state, derivative) pairs obey the published right-hand sides (dynamic.py:70-72,
    lorenz_env_transient_pmsm.py:83-85) -- guards the harvesting in make_golden.py."""
    s, d = DER["lorenz3_state"], DER["lorenz3_deriv"]
    x, y, z = s.T
    assert np.array_equal(d[:, 0], 10 * (y - x)) and np.array_equal(d[:, 1], 28 * x - y - x * z)
    assert np.array_equal(d[:, 2], x * y - (8 / 3) * z)
    s, d = DER["pmsm_classic_state"], DER["pmsm_classic_deriv"]
    x, y, z = s.T
    assert np.array_equal(d[:, 0], -x + y * z) and np.array_equal(d[:, 2], 5.46 * (y - z))


# ---- GPU --------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["long", "short"])
def test_cl_eval_metrics_against_reference_generated_fixture(tag):
    import torch
    from gym_lorenz_b200 import rl_ops
    e, u, dt = GOLD[f"{tag}_err"].astype(np.float64), GOLD[f"{tag}_ctrl"].astype(np.float64), float(GOLD[f"{tag}_dt"])
    ref = GOLD[f"{tag}_metrics"]
    out = rl_ops.eval_metrics(torch.as_tensor(e, device="cuda:0"), torch.as_tensor(u, device="cuda:0"), dt=dt)
    for k in range(e.shape[2]):
        assert np.isclose(out["mae"][k].item(), ref[k, 0], rtol=1e-12, atol=0)
        assert np.isclose(out["rmse"][k].item(), ref[k, 1], rtol=1e-12, atol=0)
        assert _same(out["settling_time"][k].item(), float(ref[k, 2])), (k, out["settling_time"][k].item(), ref[k, 2])
        assert np.isclose(out["energy"][k].item(), ref[k, 3], rtol=1e-12, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["lorenz3", "lorenz3_pair", "lorenz4_pair", "pmsm_classic", "pmsm_single"])
def test_cl_derivatives_bit_exact_against_the_reference(kind):
    """Euler kinds are compiled without FMA contraction in the reference's expression order: bit-exact."""
    import torch
    b = H.gpu_batch(kind, 128)
    st = torch.as_tensor(np.ascontiguousarray(DER[f"{kind}_state"].T), device=b.device)
    got = b.derivatives(st).cpu().numpy().T
    assert np.array_equal(got, DER[f"{kind}_deriv"]), float(np.max(np.abs(got - DER[f"{kind}_deriv"])))
    b.close()


@pytest.mark.gpu
def test_cl_derivatives_hr_against_hr_derivatives():
    """hr_derivatives (lorenz_env_try.py:7-12) with the float32 controls step() forms (clip(a)*100).
    `x1**3` is libm pow in the reference; the kernel's FMA-corrected cube agrees to the last bit almost
    everywhere (<= 1 ulp of the largest term otherwise) -- tolerance 4e-16 of the term scale."""
    import torch
    b = H.gpu_batch("hr_sync", 128)
    x = DER["hr_sync_state"]
    st = torch.as_tensor(np.ascontiguousarray(x.T), device=b.device)
    ctl = torch.as_tensor(np.ascontiguousarray(DER["hr_sync_control"].T), device=b.device)
    scale = np.maximum(np.abs(x[:, :1]) ** 3, 1.0) * np.ones((1, 3))
    for action, ref in ((ctl, DER["hr_sync_deriv"]), (None, DER["hr_sync_deriv_free"])):
        got = b.derivatives(st, action).cpu().numpy().T
        err = np.abs(got - ref) / scale
        assert err.max() <= 4e-16, err.max()
        assert (got == ref).mean() > 0.95
    b.close()
