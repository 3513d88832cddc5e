"""Batched key-stream trajectories (code/chaos_apl/main.py:1270-1309 `generate`, quantisation :230,:479-480)."""
import numpy as np
import pytest

import helpers as H


def test_quantize_equals_the_numpy_formula():
    import torch
    from gym_lorenz_b200 import keystream
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-40, 40, 5000), np.array([0.00005, -0.00005, 0.00015, 1.23455, -7.0, 0.0])])
    for scale, mod, off in ((1e4, 8, 1), (1e1, 8, 1), (1.0, 8, 1), (1e4, 256, 0)):
        ref = (np.mod(np.round(x * scale), mod) + off).astype(np.uint8)
        got = keystream.quantize(torch.as_tensor(x), scale, mod, off).numpy()
        assert np.array_equal(got, ref), (scale, mod, off)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["lorenz4_pair", "memristive4_pair"])
def test_keystream_replays_the_reference_trajectories(kind):
    """state1 / state2 sequences equal the states the UNMODIFIED reference env went through (golden fixture:
    8 initial conditions x 64 steps), bit for bit -- graph-chunked with a partial tail, and eager with the
    fixture's actions (the memristive pair is driven, the 4-D Lorenz pair ignores its action)."""
    import torch
    from gym_lorenz_b200 import keystream
    case = H.load_case(kind)
    K, T = case["actions"].shape[:2]
    acts = torch.as_tensor(case["actions"], device="cuda:0")
    got = keystream.generate(T, n_streams=K, kind=kind, burn_in=0, state0=case["st0"][:, :8],
                             actions=lambda i: acts[:, i].contiguous())
    ref = np.transpose(case["state"][:, :, :8], (2, 1, 0))        # [8, T, K]
    assert np.array_equal(got.cpu().numpy(), ref, equal_nan=True)
    if kind == "lorenz4_pair":                                     # action ignored: the zero-action graph path must agree
        for burn, chunk in ((0, 7), (10, 64), (13, 5)):
            g = keystream.generate(T, n_streams=K, kind=kind, burn_in=burn, state0=case["st0"][:, :8], chunk=chunk)
            assert g.shape == (8, T - burn, K)
            assert np.array_equal(g.cpu().numpy(), ref[:, burn:], equal_nan=True), (burn, chunk)


@pytest.mark.gpu
def test_keystream_streams_are_independent_and_seeded():
    from gym_lorenz_b200 import keystream
    a = keystream.generate(1700, n_streams=300, burn_in=1500, seed=5)
    b = keystream.generate(1700, n_streams=300, burn_in=1500, seed=5, use_cuda_graph=False)
    c = keystream.generate(1700, n_streams=300, burn_in=1500, seed=6)
    assert a.shape == (8, 200, 300)
    assert bool((a == b).all()) and not bool((a == c).all())
    assert bool(a.isfinite().all())
    q = keystream.quantize(a)
    assert int(q.min()) >= 1 and int(q.max()) <= 8
    # every DNA rule index shows up with a roughly even share over 8 x 200 x 300 samples
    share = np.bincount(q.cpu().numpy().ravel(), minlength=9)[1:] / q.numel()
    assert np.all(np.abs(share - 0.125) < 0.02), share
