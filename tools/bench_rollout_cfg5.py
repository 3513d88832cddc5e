#!/usr/bin/env python
"""BASELINE.json configs[4]: PPO-style rollout collection with train.py's hyper-parameters
(code/train.py:112-118: MlpPolicy, n_steps=2048, gae_lambda=0.95) on a 4,096-env HR
(`lorenz_try-v0`) GPU VecEnv.  stable_baselines3 is not installed in this image, so the policy is
a torch MLP with SB3's default MlpPolicy shape (pi=[64,64], vf=[64,64], tanh, diagonal Gaussian)
and the loop restates SB3's `collect_rollouts`:
  (a) numpy path  -- what stock SB3 does: policy on `device='cpu'` (train.py:118), obs/actions
                     as NumPy through VecEnv.step(), Python loop over dones;
  (b) tensor path -- DeviceRolloutCollector: GPU policy, env.step_tensor, DLPack-able views, GAE
                     on device; no host round trip in the loop.
Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_lorenz_b200 import rl_ops
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv

N = int(os.environ.get("CFG5_ENVS", 4096)); T = int(os.environ.get("CFG5_STEPS", 256))


class ActorCritic(torch.nn.Module):
    def __init__(self, obs_dim=6, act_dim=2):
        super().__init__()
        mlp = lambda: torch.nn.Sequential(torch.nn.Linear(obs_dim, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh())
        self.pi, self.vf = mlp(), mlp()
        self.mu, self.v = torch.nn.Linear(64, act_dim), torch.nn.Linear(64, 1)
        self.log_std = torch.nn.Parameter(torch.zeros(act_dim))

    def forward(self, obs):
        mu = self.mu(self.pi(obs))
        std = self.log_std.exp()
        a = mu + std * torch.randn_like(mu)
        logp = (-0.5 * ((a - mu) / std) ** 2 - self.log_std - 0.9189385).sum(-1)
        return a, self.v(self.vf(obs)).squeeze(-1), logp


def numpy_path():
    env = BatchedChaosVecEnv("hr_sync", N, seed=0)
    pol = ActorCritic()                      # CPU, as train.py does
    obs = env.reset(); starts = np.ones(N, bool)
    buf = {k: np.zeros((T, N) + s, np.float32) for k, s in (("obs", (6,)), ("actions", (2,)), ("rewards", ()), ("values", ()), ("starts", ()))}
    def run():
        nonlocal obs, starts
        for t in range(T):
            with torch.no_grad():
                a, v, lp = pol(torch.as_tensor(obs))
            a = a.numpy(); clipped = np.clip(a, -1, 1)
            new_obs, rew, dones, infos = env.step(clipped)
            for idx in np.flatnonzero(dones):           # SB3's per-env TimeLimit bootstrap loop
                if infos[idx].get("TimeLimit.truncated", False):
                    with torch.no_grad():
                        rew[idx] += 0.99 * pol(torch.as_tensor(infos[idx]["terminal_observation"][None]))[1].item()
            buf["obs"][t], buf["actions"][t], buf["rewards"][t], buf["values"][t], buf["starts"][t] = obs, a, rew, v.numpy(), starts
            obs, starts = new_obs, dones
    run(); t0 = time.perf_counter(); run(); el = time.perf_counter() - t0
    env.close()
    return N * T / el


def tensor_path(graph=False):
    env = BatchedChaosVecEnv("hr_sync", N, seed=0)
    pol = ActorCritic().to("cuda:0")
    col = rl_ops.DeviceRolloutCollector(env, pol, n_steps=T, gamma=0.99, gae_lambda=0.95, use_cuda_graph=graph)
    col.collect(); col.collect(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = col.collect(); torch.cuda.synchronize(); el = time.perf_counter() - t0
    env.close()
    return N * T / el


if __name__ == "__main__":
    a, b, c = numpy_path(), tensor_path(False), tensor_path(True)
    print(json.dumps({"config": f"cfg5: HR env x {N}, {T}-step rollout, MlpPolicy-shaped actor-critic",
                      "numpy_path_cpu_policy_fps": a, "tensor_path_gpu_policy_fps": b, "tensor_path_cuda_graph_fps": c,
                      "reference_ppo_fps_1env_windows_desktop": 1757}))
