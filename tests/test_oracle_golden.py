"""Pins oracle/chaos_oracle.c against the committed golden vectors, which were produced by
executing the UNMODIFIED reference classes (tests/golden/make_golden.py) or read from the
reference's own artefact PMSM_Origin_Data.xlsx.  Runs anywhere (no reference tree, no GPU)."""
import os

import numpy as np
import pytest

import helpers as H
import refcheck as RC


@pytest.mark.parametrize("name", H.PARITY_CASES)
def test_oracle_free_running_matches_reference_bit_exactly(oracle_api, name):
    case = H.load_case(name)
    kind, kw = case["kind"], case["kwargs"]
    K, T = case["actions"].shape[:2]
    noisy = bool(np.any(case["noise"] != 0))
    for k in range(K):
        with np.errstate(all="ignore"):
            o = RC.drive_oracle(kind, case["st0"][k], case["actions"][k], case["noise"][k] if noisy else None,
                                adam_step=int(case["adam0"][k]), **kw)
        for key in ("state", "obs", "reward"):
            assert H.same_nonfinite(o[key], case[key][k]), (name, key)
            assert H.max_rel(o[key], case[key][k]) == 0.0, (name, key, H.max_rel(o[key], case[key][k]))
        assert np.array_equal(o["done"], case["done"][k]), name


def test_oracle_cfg1_1000_step_dump(oracle_api):
    """BASELINE.json configs[0]: single dynamic.py env, 1000 random-action steps."""
    z = np.load(os.path.join(H.GOLDEN, "cfg1_lorenz3.npz"))
    for tag in ("small", "wide"):
        with np.errstate(all="ignore"):
            o = RC.drive_oracle("lorenz3", z[f"{tag}_st0"], z[f"{tag}_actions"])
        assert H.same_nonfinite(o["state"][:, :3], z[f"{tag}_state1"])
        assert H.max_rel(o["state"][:, :3], z[f"{tag}_state1"]) == 0.0
        assert H.max_rel(o["obs"], z[f"{tag}_obs"]) == 0.0
        assert H.max_rel(o["reward"], z[f"{tag}_reward"]) == 0.0
        assert H.max_rel(o["state"][:, 3], z[f"{tag}_t"]) == 0.0
        assert not o["done"].any() and not z[f"{tag}_done"].any()  # `t == 10` never fires
    assert z["small_t"][-1] != 10.0
    assert not np.isfinite(z["wide_state1"][-1]).all()  # +-500 impulses blow up


def test_oracle_reproduces_reference_xlsx_float32_kats(oracle_api):
    """PMSM_Origin_Data.xlsx (code/lorenz_pmsm/test_evaluate.py:61-166), IC (10,-10,15),(0,0,0)."""
    z = np.load(os.path.join(H.GOLDEN, "pmsm_xlsx_kat.npz"))
    cols = sorted({k.split("_")[0] for k in z.files})
    assert len(cols) >= 6
    total = 0
    for c in cols:
        acts, err, alpha = z[f"{c}_actions"], z[f"{c}_err"], float(z[f"{c}_alpha"])
        st0 = np.array([10, -10, 15, 0, 0, 0, 0, 0, 0], np.float64)
        o = RC.drive_oracle("pmsm_sync", st0, acts, alpha=alpha)
        e = (o["state"][:, :3].astype(np.float32) - o["state"][:, 3:6].astype(np.float32))
        assert np.array_equal(e, err), c
        total += len(acts)
    assert total == 51 + 123 + 344 + 191 + 124 + 56
    # the SURVEY 8c spot values
    a0 = z["a0_err"]
    assert a0[0].tolist() == [9.890000343322754, -9.890000343322754, 14.863499641418457]
    assert a0[1].tolist() == [9.783853530883789, -9.779096603393555, 14.72834587097168]
    o = RC.drive_oracle("pmsm_sync", np.array([10, -10, 15, 0, 0, 0, 0, 0, 0.0]), z["a0_actions"][:3], alpha=0.5)
    assert o["reward"].tolist() == [-44.790496826171875, -44.388118743896484, -43.990455627441406]


def test_oracle_autoreset_and_timelimit_contract(oracle_api):
    """SB3 DummyVecEnv.step_wait + gymnasium TimeLimit semantics, pinned by construction:
    truncated at max_episode_steps, terminal obs kept, returned obs is the reset obs,
    episode return/length accumulated (Monitor)."""
    O = oracle_api
    n, T, lim = 5, 25, 10
    orc = O.Oracle("lorenz3", n, flags=O.F_AUTORESET, max_episode_steps=lim, seed=3)
    orc.reset()
    rng = np.random.default_rng(0)
    acts = rng.uniform(-0.05, 0.05, (T, 3, orc.n_pad)).astype(np.float32)
    out = orc.rollout(T, acts)
    done = out["done"][:, :n]
    assert np.array_equal(np.flatnonzero(done[:, 0]), [9, 19])
    assert np.all(done[9] == 2) and np.all(done[19] == 2)  # truncated, not terminated
    assert np.all(out["last_ep_len"][:n] == lim)
    assert orc.stats[0] == 2 * n and orc.stats[3] == 2 * n * lim and orc.stats[6] == 2 * n
    # return accounting: sum of the 10 rewards of the last finished episode
    assert np.allclose(out["last_ep_ret"][:n], out["reward"][10:20, :n].sum(0), rtol=1e-15)
    # obs at a done step is a fresh reset obs: |x|,|y|,|z| <= 30 and differs from terminal obs
    assert np.all(np.abs(out["obs"][9, :3, :n]) <= 30.0)
    assert not np.array_equal(out["obs"][9, :, :n], out["term_obs"][9, :, :n])
    assert np.all(orc.ep_len[:n] == T - 20)
