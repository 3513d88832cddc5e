"""TEST INFRASTRUCTURE ONLY -- NumPy restatements of the stable_baselines3==2.7.1 pieces that the
reference's pipelines wrap around the env (the library is un-vendored and not installed here:
these follow SB3's published source) and of the reference's own metric functions.  What the
reference's own SB3 artefacts hold about these pieces (the pickled VecNormalize objects and model zips of
code/lorenz_pmsm/train.py's eight finished runs: update schedule and counts of obs_rms / ret_rms, defaults,
Monitor records, TimeLimit) is extracted by tests/golden/make_sb3_artefact_golden.py and checked in
tests/test_sb3_artefacts.py; the moment-merge arithmetic and GAE are pinned by no artefact.

  gae                       common/buffers.py::RolloutBuffer.compute_returns_and_advantage
  RunningMeanStd            common/running_mean_std.py
  normalize_obs             common/vec_env/vec_normalize.py::_normalize_obs / normalize_obs
  VecNormalizeRef           common/vec_env/vec_normalize.py::VecNormalize.reset / step_wait / _update_reward
  frame_stack_update        common/vec_env/stacked_observations.py::StackedObservations.update (1-D obs)
  calculate_advanced_metrics, steady_state_metrics
                            /root/reference/code/lorenz_pmsm/test_evaluate.py:25-59, :239-250
"""
from __future__ import annotations

import numpy as np


def gae(rewards, values, episode_starts, last_values, dones, gamma, gae_lambda):
    rewards, values, episode_starts = (np.asarray(x, np.float32) for x in (rewards, values, episode_starts))
    last_values = np.asarray(last_values, np.float32).flatten()
    T = rewards.shape[0]
    advantages = np.zeros_like(rewards)
    last_gae_lam = 0
    for step in reversed(range(T)):
        if step == T - 1:
            next_non_terminal = 1.0 - np.asarray(dones).astype(np.float32)
            next_values = last_values
        else:
            next_non_terminal = 1.0 - episode_starts[step + 1]
            next_values = values[step + 1]
        delta = rewards[step] + gamma * next_values * next_non_terminal - values[step]
        last_gae_lam = delta + gamma * gae_lambda * next_non_terminal * last_gae_lam
        advantages[step] = last_gae_lam
    return advantages, advantages + values


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, arr):
        self.update_from_moments(np.mean(arr, axis=0), np.var(arr, axis=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m_2 = m_a + m_b + np.square(delta) * self.count * batch_count / (self.count + batch_count)
        self.mean, self.var, self.count = new_mean, m_2 / (self.count + batch_count), batch_count + self.count


def normalize_obs(obs, rms, clip_obs=10.0, epsilon=1e-8):
    return np.clip((obs - rms.mean) / np.sqrt(rms.var + epsilon), -clip_obs, clip_obs).astype(np.float32)


class VecNormalizeRef:
    """VecNormalize over a callable env pair: `reset_fn() -> obs [N, D]`, `step_fn(actions) -> (obs, rewards,
    dones)` (auto-resetting, like DummyVecEnv).  Statement order of SB3's reset() / step_wait()."""

    def __init__(self, reset_fn, step_fn, num_envs, obs_dim, training=True, norm_obs=True, norm_reward=True,
                 clip_obs=10.0, clip_reward=10.0, gamma=0.99, epsilon=1e-8):
        self.reset_fn, self.step_fn = reset_fn, step_fn
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon
        self.obs_rms = RunningMeanStd(shape=(obs_dim,))
        self.ret_rms = RunningMeanStd(shape=())
        self.returns = np.zeros(num_envs)
        self.old_obs = self.old_reward = None

    def normalize_obs(self, obs):
        return normalize_obs(obs, self.obs_rms, self.clip_obs, self.epsilon) if self.norm_obs else obs

    def normalize_reward(self, reward):
        if not self.norm_reward:
            return reward
        return np.clip(reward / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward,
                       self.clip_reward).astype(np.float32)

    def reset(self):
        obs = self.reset_fn()
        self.old_obs = obs
        self.returns = np.zeros(len(self.returns))
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        return self.normalize_obs(obs)

    def step(self, actions):
        obs, rewards, dones = self.step_fn(actions)
        self.old_obs, self.old_reward = obs, rewards
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        nobs = self.normalize_obs(obs)
        if self.training:                                   # _update_reward: not gated on norm_reward
            self.returns = self.returns * self.gamma + rewards
            self.ret_rms.update(self.returns)
        nrew = self.normalize_reward(rewards)
        self.returns[np.asarray(dones, bool)] = 0
        return nobs, nrew, dones


def frame_stack_update(stacked, obs, dones):
    dim = obs.shape[-1]
    stacked = np.roll(stacked, shift=-dim, axis=-1)
    for i, d in enumerate(dones):
        if d:
            stacked[i] = 0
    stacked[..., -dim:] = obs
    return stacked


def calculate_advanced_metrics(arr_e, arr_a1, arr_a2, dt=0.01, error_band=0.05):
    exceed = np.where(np.abs(arr_e) > error_band)[0]
    if len(exceed) == 0:
        settling_time = 0.0
    else:
        stable_idx = exceed[-1] + 1
        settling_time = stable_idx * dt if stable_idx < len(arr_e) else np.nan
    energy_cost = np.sum(np.square(arr_a1) + np.square(arr_a2)) * dt
    return settling_time, energy_cost


def steady_state_metrics(e, u, dt):
    """e [T,3], u [T,2] -> (mae, rmse, max settling time, energy) as test_evaluate.py:239-250."""
    start = min(1000, len(e) // 2)
    s = e[start:]
    mae = np.mean([np.mean(np.abs(s[:, c])) for c in range(e.shape[1])])
    rmse = np.mean([np.sqrt(np.mean(s[:, c] ** 2)) for c in range(e.shape[1])])
    ts, energy = [], 0.0
    for c in range(e.shape[1]):
        t_c, energy = calculate_advanced_metrics(e[:, c], u[:, 0], u[:, 1], dt=dt)
        ts.append(t_c)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            max_ts = np.nanmax(ts)
    return mae, rmse, max_ts, energy
