"""Device-resident counterparts of the SB3 pieces that wrap the env step in the reference's
pipelines.  Everything stays in HBM as torch tensors; the arithmetic is
in csrc/tu_rl_ops.cu.

  gae(...)               SB3 RolloutBuffer.compute_returns_and_advantage   (code/train.py:112-120)
  DeviceVecNormalize     SB3 VecNormalize(norm_obs, norm_reward, clip_obs) (code/lorenz_pmsm/train.py:118,170)
  DeviceVecFrameStack    SB3 VecFrameStack(n_stack)                        (code/lorenz_filter/train.py:115)
  eval_metrics(...)      calculate_advanced_metrics + steady-state MAE/RMSE (code/lorenz_pmsm/test_evaluate.py:25-59,239-250)
  DeviceRolloutCollector SB3 OnPolicyAlgorithm.collect_rollouts without NumPy / host round trips
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional, Tuple

import torch

from . import _lib as L


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)   # absent in CPU-only builds


def _stream(dev: torch.device):
    """torch's current stream on `dev` as a raw cudaStream_t (see ChaosBatch._stream)."""
    if _RAW_STREAM is not None and dev.index is not None:
        return _RAW_STREAM(dev.index)
    return torch.cuda.current_stream(dev).cuda_stream


def gae(rewards: torch.Tensor, values: torch.Tensor, episode_starts: torch.Tensor, last_values: torch.Tensor,
        last_dones: torch.Tensor, gamma: float, gae_lambda: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """advantages, returns (float32 [T, N]) exactly as SB3 computes them (float32 recurrence)."""
    lib = L.load()
    T, N = rewards.shape
    dev = rewards.device
    args = [x.to(dev, torch.float32).contiguous() for x in (rewards, values, episode_starts)]
    lv = last_values.to(dev, torch.float32).contiguous().view(-1)
    ld = last_dones.to(dev, torch.float32).contiguous().view(-1)
    adv, ret = torch.empty_like(args[0]), torch.empty_like(args[0])
    with torch.cuda.device(dev):
        L.check(lib.cl_gae(_stream(dev), _p(args[0]), _p(args[1]), _p(args[2]), _p(lv), _p(ld), float(gamma),
                           float(gae_lambda), T, N, N, _p(adv), _p(ret)), None, "cl_gae")
    return adv, ret


def eval_metrics(err: torch.Tensor, ctrl: torch.Tensor, dt: float, error_band: float = 0.05,
                 steady_start: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """err f64 [T, n_err, N], ctrl f64 [T, n_ctrl, N] -> per-trajectory mae/rmse/settling_time/energy."""
    lib = L.load()
    T, ne, N = err.shape
    nc = ctrl.shape[1]
    dev = err.device
    err = err.to(dev, torch.float64).contiguous()
    ctrl = ctrl.to(dev, torch.float64).contiguous()
    if steady_start is None:
        steady_start = min(1000, T // 2)          # test_evaluate.py:239
    out = torch.empty((N, 4), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.cl_eval_metrics(_stream(dev), _p(err), _p(ctrl), T, ne, nc, N, N, int(steady_start), float(dt),
                                    float(error_band), _p(out)), None, "cl_eval_metrics")
    return {"mae": out[:, 0], "rmse": out[:, 1], "settling_time": out[:, 2], "energy": out[:, 3]}


def base_batch(env):
    """The ChaosBatch under any stack of device wrappers (walks `.venv` until `.batch` is found)."""
    e = env
    while not hasattr(e, "batch"):
        if not hasattr(e, "venv"):
            raise AttributeError(f"{type(env).__name__} wraps no BatchedChaosVecEnv")
        e = e.venv
    return e.batch


class RunningMeanStd:
    """SB3 RunningMeanStd (float64 mean/var/count, Chan et al. parallel update).  Batch moments and
    the merge both run on the device and update `mean` / `var` / `count_t` in place: no host
    synchronisation, and an `update` can be captured in a CUDA graph and replayed."""

    def __init__(self, shape, device, epsilon: float = 1e-4):
        self.mean = torch.zeros(shape, dtype=torch.float64, device=device)
        self.var = torch.ones(shape, dtype=torch.float64, device=device)
        self.count_t = torch.full((1,), float(epsilon), dtype=torch.float64, device=device)
        self._acc = torch.zeros((2,) + tuple(shape), dtype=torch.float64, device=device)

    @property
    def count(self) -> float:
        """Host copy of the running count (synchronises; for inspection / checkpoints)."""
        return float(self.count_t.item())

    @count.setter
    def count(self, v: float) -> None:
        self.count_t.fill_(float(v))

    def update(self, x: torch.Tensor) -> None:
        """x: f32 or f64 [N, dim] (any strides)."""
        lib = L.load()
        n, dim = x.shape
        dev = x.device
        if x.dtype not in (torch.float32, torch.float64):
            raise ValueError("RunningMeanStd.update expects float32 or float64")
        with torch.cuda.device(dev):
            st = _stream(dev)
            fn, name = (lib.cl_obs_moments, "cl_obs_moments") if x.dtype == torch.float32 else \
                (lib.cl_moments_f64, "cl_moments_f64")
            L.check(fn(st, _p(x), x.stride(0), x.stride(1), n, dim, _p(self.mean), _p(self._acc)), None, name)
            L.check(lib.cl_rms_update(st, _p(self._acc), n, dim, _p(self.mean), _p(self.var), _p(self.count_t)),
                    None, "cl_rms_update")


class DeviceVecNormalize:
    """VecNormalize on device tensors: `reset_tensor()` / `step_tensor(actions)` of the wrapped
    env, observations normalised with running statistics (updated while `training`), rewards
    optionally normalised by the running std of the discounted return."""

    def __init__(self, venv, training: bool = True, norm_obs: bool = True, norm_reward: bool = True,
                 clip_obs: float = 10.0, clip_reward: float = 10.0, gamma: float = 0.99, epsilon: float = 1e-8):
        self.venv, self.training = venv, training
        self.norm_obs, self.norm_reward = norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon
        self.num_envs = venv.num_envs
        b = base_batch(venv)
        self.device = b.device
        self.obs_rms = RunningMeanStd((b.obs_dim,), self.device)
        self.ret_rms = RunningMeanStd((1,), self.device)
        self.returns = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
        self._out = torch.empty((self.num_envs, b.obs_dim), dtype=torch.float32, device=self.device)
        self._term = torch.empty_like(self._out)
        self.old_obs = None
        self.old_reward = None

    def normalize_obs(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not self.norm_obs:
            return obs
        lib = L.load()
        out = torch.empty(obs.shape, dtype=torch.float32, device=obs.device) if out is None else out
        n, dim = obs.shape
        with torch.cuda.device(obs.device):
            L.check(lib.cl_obs_normalize(_stream(obs.device), _p(obs), obs.stride(0), obs.stride(1), _p(out),
                                         out.stride(0), out.stride(1), n, dim, _p(self.obs_rms.mean),
                                         _p(self.obs_rms.var), float(self.epsilon), float(self.clip_obs)),
                    None, "cl_obs_normalize")
        return out

    def normalize_reward(self, reward: torch.Tensor) -> torch.Tensor:
        if not self.norm_reward:
            return reward
        r = reward.double() / torch.sqrt(self.ret_rms.var[0] + self.epsilon)
        return torch.clamp(r, -self.clip_reward, self.clip_reward).float()

    def reset_tensor(self) -> torch.Tensor:
        obs = self.venv.reset_tensor()
        self.old_obs = obs
        self.returns.zero_()
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        return self.normalize_obs(obs, self._out)

    def step_tensor(self, actions: torch.Tensor):
        obs, rew, done = self.venv.step_tensor(actions)
        self.old_obs, self.old_reward = obs, rew
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        nobs = self.normalize_obs(obs, self._out)
        if self.training:
            # SB3 VecNormalize.step_wait -> _update_reward: whenever `training`, float64 returns
            self.returns.mul_(self.gamma).add_(rew)      # in place: replayable as a captured graph
            self.ret_rms.update(self.returns.view(-1, 1))
        nrew = self.normalize_reward(rew)
        self.returns.masked_fill_(done != 0, 0.0)
        return nobs, nrew, done

    def terminal_obs(self) -> torch.Tensor:
        inner = self.venv.terminal_obs() if hasattr(self.venv, "terminal_obs") else base_batch(self.venv).terminal_obs()
        return self.normalize_obs(inner, self._term)

    def get_original_obs(self) -> torch.Tensor:
        return self.old_obs

    def state_dict(self):
        return {"obs_mean": self.obs_rms.mean.clone(), "obs_var": self.obs_rms.var.clone(),
                "obs_count": self.obs_rms.count, "ret_mean": self.ret_rms.mean.clone(),
                "ret_var": self.ret_rms.var.clone(), "ret_count": self.ret_rms.count}

    def load_state_dict(self, sd):
        for rms, k in ((self.obs_rms, "obs"), (self.ret_rms, "ret")):   # in place: captured graphs keep pointing here
            rms.mean.copy_(sd[k + "_mean"]); rms.var.copy_(sd[k + "_var"]); rms.count = sd[k + "_count"]


class DeviceVecFrameStack:
    """VecFrameStack(n_stack) for 1-D observations, stacked along the last axis, on device.
    Wraps a BatchedChaosVecEnv or a DeviceVecNormalize (code/lorenz_filter/train.py:109-115)."""

    def __init__(self, venv, n_stack: int):
        self.venv, self.n_stack = venv, int(n_stack)
        self.num_envs = venv.num_envs
        src = base_batch(venv)
        self.dim = src.obs_dim
        self.device = src.device
        self.stacked = torch.zeros((self.num_envs, self.dim * self.n_stack), dtype=torch.float32, device=self.device)
        self._term = torch.zeros_like(self.stacked)

    def _inner_terminal_obs(self) -> torch.Tensor:
        return self.venv.terminal_obs() if hasattr(self.venv, "terminal_obs") else base_batch(self.venv).terminal_obs()

    def _push(self, obs: torch.Tensor, done: Optional[torch.Tensor], term: Optional[torch.Tensor] = None) -> torch.Tensor:
        lib = L.load()
        with torch.cuda.device(self.device):
            if term is None or done is None:
                L.check(lib.cl_frame_stack(_stream(self.device), _p(self.stacked), _p(obs), obs.stride(0), obs.stride(1),
                                           _p(done), self.num_envs, self.dim, self.n_stack), None, "cl_frame_stack")
            else:
                L.check(lib.cl_frame_stack_term(_stream(self.device), _p(self.stacked), _p(obs), obs.stride(0),
                                                obs.stride(1), _p(done), _p(term), term.stride(0), term.stride(1),
                                                _p(self._term), self.num_envs, self.dim, self.n_stack), None,
                        "cl_frame_stack_term")
        return self.stacked

    def reset_tensor(self) -> torch.Tensor:
        obs = self.venv.reset_tensor()
        self.stacked.zero_()
        return self._push(obs, None)

    def step_tensor(self, actions: torch.Tensor):
        obs, rew, done = self.venv.step_tensor(actions)
        return self._push(obs, done, self._inner_terminal_obs()), rew, done

    def terminal_obs(self) -> torch.Tensor:
        """[N, dim * n_stack]: for envs that finished an episode in the last step, the stacked terminal
        observation of SB3's StackedObservations.update (previous stack rolled by one frame + the inner
        env's -- normalised, if wrapped -- terminal observation).  Other rows are undefined."""
        return self._term


class DeviceRolloutCollector:
    """SB3 `OnPolicyAlgorithm.collect_rollouts` on device tensors: policy forward -> env step ->
    buffer write, then GAE, with no NumPy and no host round trip inside the loop.

    `policy(obs) -> (actions, values, log_probs)` is any torch callable (the learner's network);
    time-limit truncations bootstrap with the value of the terminal observation like SB3
    (`rewards[idx] += gamma * V(terminal_obs)`).

    `use_cuda_graph=True` captures the whole n_steps loop (policy kernels, env step kernels,
    buffer writes, GAE) into ONE CUDA graph and replays it per rollout: at a few thousand envs the
    loop is launch-latency bound (~30 small kernels per step), which is exactly what graphs remove.
    The env is switched to graph mode (device-resident Philox step index), so replays keep
    advancing the random streams.
    """

    def __init__(self, env, policy: Callable, n_steps: int, gamma: float = 0.99, gae_lambda: float = 0.95,
                 use_cuda_graph: bool = False):
        self.env, self.policy, self.n_steps = env, policy, int(n_steps)
        self.gamma, self.gae_lambda = gamma, gae_lambda
        b = base_batch(env)
        self.batch, dev, N, T = b, b.device, b.num_envs, self.n_steps
        self.obs_dim = env.stacked.shape[1] if hasattr(env, "stacked") else b.obs_dim
        self.buf = {
            "obs": torch.empty((T, N, self.obs_dim), dtype=torch.float32, device=dev),
            "actions": torch.empty((T, N, b.act_dim), dtype=torch.float32, device=dev),
            "rewards": torch.empty((T, N), dtype=torch.float32, device=dev),
            "values": torch.empty((T, N), dtype=torch.float32, device=dev),
            "log_probs": torch.empty((T, N), dtype=torch.float32, device=dev),
            "episode_starts": torch.empty((T, N), dtype=torch.float32, device=dev),
        }
        self._last_obs = None
        self._last_starts = torch.ones(N, dtype=torch.float32, device=dev)
        self._lo = float(b.layout.act_low)
        self._hi = float(b.layout.act_high)
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graph = None
        self._graph_out = None
        self._warm = False

    def _terminal_obs(self):
        return self.env.terminal_obs() if hasattr(self.env, "terminal_obs") else self.batch.terminal_obs()

    def _loop(self, sync_free: bool) -> Dict[str, torch.Tensor]:
        buf = self.buf
        for t in range(self.n_steps):
            obs = self._last_obs
            actions, values, log_probs = self.policy(obs)
            buf["obs"][t].copy_(obs)
            buf["actions"][t].copy_(actions)
            buf["values"][t].copy_(values.view(-1))
            buf["log_probs"][t].copy_(log_probs.view(-1))
            buf["episode_starts"][t].copy_(self._last_starts)
            new_obs, rew, done = self.env.step_tensor(torch.clamp(actions, self._lo, self._hi))
            rew = rew.float().clone()
            trunc = (done & L.DONE_TRUNCATED).bool() & ~(done & L.DONE_TERMINATED).bool()
            if sync_free or bool(trunc.any()):   # TimeLimit bootstrap (SB3 collect_rollouts)
                _, tv, _ = self.policy(self._terminal_obs())
                rew = torch.where(trunc, rew + self.gamma * tv.view(-1), rew)
            buf["rewards"][t].copy_(rew)
            self._last_starts.copy_((done != 0).float())
            self._last_obs.copy_(new_obs)
        _, last_values, _ = self.policy(self._last_obs)
        adv, ret = gae(buf["rewards"], buf["values"], buf["episode_starts"], last_values.view(-1), self._last_starts,
                       self.gamma, self.gae_lambda)
        out = dict(buf)
        out["advantages"], out["returns"] = adv, ret
        return out

    @torch.no_grad()
    def collect(self) -> Dict[str, torch.Tensor]:
        if self._last_obs is None:
            self._last_obs = self.env.reset_tensor().clone()
        if not self.use_cuda_graph:
            return self._loop(sync_free=False)
        if self._graph is None:
            dev = self.batch.device
            if not self._warm:
                # first rollout: eager, sync-free -- it is a real rollout AND the warm-up that
                # CUDA-graph capture needs (lazy initialisation, allocator, occupancy queries)
                self.batch.set_graph_mode(True)
                self._warm = True
                return self._loop(sync_free=True)
            torch.cuda.synchronize(dev)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):            # capture executes nothing
                self._graph_out = self._loop(sync_free=True)
        self._graph.replay()
        return self._graph_out
