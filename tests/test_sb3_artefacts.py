"""What the reference's own stable_baselines3 artefacts pin about the wrappers around the env step.

tests/golden/sb3_artefacts.json is extracted (tests/golden/make_sb3_artefact_golden.py) from the eight finished
training runs the reference ships -- `pmsm_a2c_alpha_*_clean_model` (SB3 2.7.1 zips) and
`pmsm_a2c_alpha_*_clean_vecnorm.pkl` (pickled VecNormalize), written by code/lorenz_pmsm/train.py:152-190:
DummyVecEnv([Monitor(gym.make("lorenz_pmsm-v0"))]) -> VecNormalize(norm_obs=True, norm_reward=False,
clip_obs=10.0) -> A2C.learn(1_000_000).  They are the only SB3-produced numbers in the reference tree.

Pinned here (rows a13 and f1 of DESIGN.md section 0):
  * RunningMeanStd: count starts at 1e-4 and grows by the batch size, in float64, once per reset() and once
    per step for obs_rms, once per step for ret_rms -- bit pattern of both counts after 1,000,000 steps;
  * VecNormalize: ret_rms is updated although norm_reward=False (training=True); defaults clip_reward=10,
    gamma=0.99, epsilon=1e-8; old_obs / old_reward are the float32 unnormalised values;
  * TimeLimit / DummyVecEnv auto-reset: all 100 recorded episodes are exactly max_episode_steps = 2000 long,
    and after 1,000,000 = 500 x 2000 steps `_last_episode_starts` is [True] -- the step counter restarts with
    the auto-reset and the 2000th step of an episode is the done one;
  * the observation returned on a done step is the RESET observation (error state within the reset box
    (-60, 60), e'_3 = sigma (e_2 - e_3) as lorenz_env_try_pmsm.py:64-75 builds it);
  * Monitor: `r` rounded to 6 decimals, `l` an int, `t` rounded to 6 decimals and increasing, a window of 100.
Not pinned by any artefact: the mean / variance merge of update_from_moments beyond its count, GAE.
"""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

import helpers as H
from oracle import sb3_ref as S

with open(os.path.join(H.GOLDEN, "sb3_artefacts.json")) as _f:
    ART = json.load(_f)
RUNS = ART["runs"]
SIGMA = 5.46        # lorenz_env_try_pmsm.py:15


def test_artefacts_are_the_eight_alpha_runs_of_sb3_2_7_1():
    assert ART["stable_baselines3"] == "2.7.1"
    assert sorted(RUNS) == ["0.10", "0.11", "0.12", "0.14", "0.17", "0.25", "0.33", "0.50"]
    for r in RUNS.values():
        assert (r["num_timesteps"], r["n_envs"], r["num_envs"]) == (1_000_000, 1, 1)
        assert (r["norm_obs"], r["norm_reward"], r["training"]) == (True, False, True)


def test_running_mean_std_count_bit_pattern_after_a_million_steps():
    rms = S.RunningMeanStd(shape=())
    for _ in range(1_000_000):
        rms.update_from_moments(0.0, 0.0, 1)
    ret_count = float(rms.count)
    rms.update_from_moments(0.0, 0.0, 1)
    obs_count = float(rms.count)
    for r in RUNS.values():
        assert float.fromhex(r["ret_count"]) == ret_count       # one update per step, norm_reward=False
        assert float.fromhex(r["obs_count"]) == obs_count       # + the one in reset()


def _oracle_pmsm_env(oracle_api, n, seed, max_episode_steps=2000):
    o = oracle_api.Oracle("pmsm_sync", n, flags=oracle_api.F_AUTORESET, max_episode_steps=max_episode_steps,
                          dt=0.001, alpha=0.25, seed=seed)

    def reset_fn():
        return o.reset()[:, :n].T.astype(np.float32)

    def step_fn(actions):
        a = np.zeros((o.act_dim, o.n_pad), np.float32)
        a[:, :n] = np.asarray(actions, np.float32).T
        obs, rew, done, out = o.step(a)
        step_fn.last = out
        return obs[:, :n].T.astype(np.float32), rew[:n].astype(np.float32), done[:n] != 0

    return o, reset_fn, step_fn


def test_vecnormalize_restatement_reproduces_the_artefact_schedule_and_defaults(oracle_api):
    run = RUNS["0.25"]
    o, reset_fn, step_fn = _oracle_pmsm_env(oracle_api, 1, seed=3)
    vn = S.VecNormalizeRef(reset_fn, step_fn, 1, 6, norm_obs=True, norm_reward=False, clip_obs=10.0)
    assert (vn.clip_reward, vn.gamma, vn.epsilon) == (run["clip_reward"], run["gamma"], run["epsilon"])
    obs = vn.reset()
    assert obs.dtype == np.float32
    rng = np.random.default_rng(0)
    k = 4100
    ref = S.RunningMeanStd(shape=())
    ref.update_from_moments(0.0, 0.0, 1)
    for t in range(k):
        nobs, nrew, done = vn.step(rng.uniform(-1, 1, (1, 2)).astype(np.float32))
        ref.update_from_moments(0.0, 0.0, 1)
        assert np.array_equal(nrew, vn.old_reward)              # norm_reward=False: rewards pass through
        assert np.all(np.abs(nobs) <= 10.0)
    assert float(vn.obs_rms.count) == float(ref.count)
    assert np.isclose(vn.ret_rms.count, 1e-4 + k, rtol=0, atol=1e-6)
    assert vn.old_obs.dtype == np.float32 and vn.old_reward.dtype == np.float32
    assert len(run["old_obs"]) == 6 and len(run["old_reward"]) == 1


def test_timelimit_and_autoreset_match_the_recorded_episodes(oracle_api):
    """500 episodes of exactly 2000 steps fill 1,000,000 steps and leave `_last_episode_starts` = [True]; the
    oracle's TimeLimit / auto-reset accounting must produce the same pattern."""
    for r in RUNS.values():
        assert set(r["ep_l"]) == {2000} and len(r["ep_l"]) == r["stats_window_size"] == 100
        assert r["num_timesteps"] % 2000 == 0 and r["last_episode_starts"] == [True]
    n = 4
    o, reset_fn, step_fn = _oracle_pmsm_env(oracle_api, n, seed=11)
    reset_fn()
    T = 4001
    rng = np.random.default_rng(2)
    acts = np.zeros((T, o.act_dim, o.n_pad), np.float32)
    acts[:, :, :n] = rng.uniform(-0.2, 0.2, (T, o.act_dim, n))
    out = o.rollout(T, acts)
    done = out["done"][:, :n]
    when = np.flatnonzero(done.any(axis=1))
    assert when.tolist() == [1999, 3999]                        # the 2000th step of each episode, counter restarts
    assert np.all(done[when] & 2) and not np.any(done[when] & 1)  # truncated, not terminated
    assert np.all(out["last_ep_len"][:n] == 2000)
    # the observation of a done step is the reset observation, the terminal one goes to term_obs
    for t in when:
        ob, tob = out["obs"][t][:, :n], out["term_obs"][t][:, :n]
        assert np.all(np.abs(ob[:3]) <= 60.0) and not np.array_equal(ob, tob)
        assert np.allclose(ob[5], SIGMA * (ob[1] - ob[2]), rtol=0, atol=2e-3)


def test_reset_observation_structure_of_the_recorded_last_observations():
    """`_last_original_obs` (model zip) and `old_obs` (VecNormalize pickle) were returned by the done step
    1,000,000: reset observations [e, f(s1) - f(s2)] with s1, s2 ~ U(-30, 30)^3 and zero action."""
    for r in RUNS.values():
        for key in ("last_original_obs", "old_obs"):
            ob = np.asarray(r[key], np.float32)
            assert ob.shape == (6,) and np.all(np.abs(ob[:3]) <= 60.0)
            assert np.isclose(ob[5], SIGMA * (ob[1] - ob[2]), rtol=0, atol=2e-3)
        assert np.all(np.abs(np.asarray(r["last_obs"])) <= r["clip_obs"])


def test_monitor_record_format():
    for r in RUNS.values():
        ep_r, ep_t = np.asarray(r["ep_r"]), np.asarray(r["ep_t"])
        assert all(round(x, 6) == x for x in r["ep_r"]) and all(round(x, 6) == x for x in r["ep_t"])
        assert np.all(np.diff(ep_t) > 0) and np.all(np.isfinite(ep_r))


# ---------------------------------------------------------------------------------- GPU side

@pytest.mark.gpu
def test_gpu_vecenv_monitor_and_timelimit_follow_the_artefacts():
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    n = 3
    env = BatchedChaosVecEnv("pmsm_sync", n, seed=5, alpha=0.25)       # default TimeLimit: the registered 2000
    env.reset()
    rng = np.random.default_rng(4)
    ends = []
    ret = np.zeros(n, np.float64)
    for t in range(4000):
        obs, rew, dones, infos = env.step(rng.uniform(-0.2, 0.2, (n, 2)).astype(np.float32))
        ret += rew.astype(np.float64)
        if dones.any():
            assert dones.all()
            ends.append(t)
            for i in range(n):
                info = infos[i]
                assert info["TimeLimit.truncated"] is True
                assert info["episode"]["l"] == 2000 and isinstance(info["episode"]["l"], int)
                assert info["episode"]["r"] == round(info["episode"]["r"], 6)
                assert abs(info["episode"]["r"] - ret[i]) <= 1e-6 * max(1.0, abs(ret[i]))
                assert info["episode"]["t"] == round(info["episode"]["t"], 6)
                tob = info["terminal_observation"]
                assert tob.shape == (6,) and not np.array_equal(tob, obs[i])
                assert np.all(np.abs(obs[i, :3]) <= 60.0)
                assert np.isclose(obs[i, 5], SIGMA * (obs[i, 1] - obs[i, 2]), rtol=0, atol=2e-3)
            ret[:] = 0.0
        else:
            assert all(not infos[i] for i in range(n))
    assert ends == [1999, 3999]
    env.close()


@pytest.mark.gpu
def test_gpu_vecnormalize_counts_follow_the_artefact_schedule():
    import torch
    from gym_lorenz_b200 import rl_ops
    from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
    run = RUNS["0.25"]
    env = BatchedChaosVecEnv("pmsm_sync", 1, seed=5, alpha=0.25)
    vn = rl_ops.DeviceVecNormalize(env, norm_obs=True, norm_reward=False, clip_obs=10.0)
    assert (vn.clip_reward, vn.gamma, vn.epsilon) == (run["clip_reward"], run["gamma"], run["epsilon"])
    vn.reset_tensor()
    ref = S.RunningMeanStd(shape=())
    ref.update_from_moments(0.0, 0.0, 1)
    g = torch.Generator(device="cpu").manual_seed(1)
    k = 2100
    for _ in range(k):
        nobs, nrew, done = vn.step_tensor((torch.rand((1, 2), generator=g) * 2 - 1).to("cuda:0"))
        ref.update_from_moments(0.0, 0.0, 1)
    assert vn.obs_rms.count == float(ref.count)                       # reset + k steps, bit for bit
    ref2 = S.RunningMeanStd(shape=())
    for _ in range(k):
        ref2.update_from_moments(0.0, 0.0, 1)
    assert vn.ret_rms.count == float(ref2.count)                      # updated although norm_reward=False
    assert torch.equal(nrew, vn.old_reward) and bool((nobs.abs() <= 10.0).all())
    env.close()
