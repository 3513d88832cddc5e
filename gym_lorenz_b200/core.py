"""ChaosBatch: a slab of N envs resident on one B200, driven through the C-ABI.

PyTorch is plumbing here: it owns the device memory (SoA planes) and the CUDA stream; all
arithmetic of the hot path runs in the hand-written kernels of libchaos_b200.so.  Device
tensors returned by `step` / `reset` / `rollout` are zero-copy views of those planes (hand
them to a policy directly, or export with `torch.utils.dlpack.to_dlpack`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib as L

_KIND_DEFAULTS = {
    # kind: (dt, substeps, act_limit, act_gain)
    L.LORENZ_RK4: (0.01, 16, 1.0, 50.0),
    L.LORENZ_RK4_F32: (0.01, 16, 1.0, 50.0),
    L.PMSM_RK4: (0.001, 4, 1.0, 50.0),
}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)   # absent in CPU-only builds


class ChaosBatch:
    """N envs of one kind on one GPU.

    Parameters mirror the reference constructors: `add_noise`, `eval_mode`, `add_filter`
    (HRSyncEnv, lorenz_env_try.py:19) and `alpha`, `add_noise` (PMSM_Sync_Env,
    lorenz_env_try_pmsm.py:9); `max_episode_steps` is the gymnasium TimeLimit of the
    registration (gym_lorenz/__init__.py:12,20).  New: num_envs, device, seed, env_id_base
    (global index of env 0: Philox subsequences are keyed by the global env index so that a
    slab behaves the same on any rank), autoreset, substeps / dt / act_limit / act_gain /
    param_jitter for the RK4 kinds.
    """

    def __init__(self, kind, num_envs: int, *, device="cuda:0", seed: int = 0, env_id_base: int = 0,
                 autoreset: bool = True, max_episode_steps: Optional[int] = None,
                 add_noise: bool = False, eval_mode: bool = False, add_filter: bool = False,
                 alpha: float = 0.5, substeps: Optional[int] = None, dt: Optional[float] = None,
                 act_limit: Optional[float] = None, act_gain: Optional[float] = None,
                 param_jitter: float = 0.0, obs_f64: bool = False):
        self.lib = L.load()
        self.kind = L.KIND_NAMES[kind] if isinstance(kind, str) else int(kind)
        self.kind_name = {v: k for k, v in L.KIND_NAMES.items()}[self.kind]
        self.layout = L.layout(self.kind)
        self.num_envs = int(num_envs)
        self.n_pad = ((self.num_envs + 127) // 128) * 128
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.ChaosLibError("ChaosBatch needs a CUDA device; there is no CPU fallback")
        if not torch.cuda.is_available():
            raise L.ChaosLibError("CUDA is not available; there is no CPU fallback")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self._dev_index = int(dev_index)
        d_dt, d_sub, d_lim, d_gain = _KIND_DEFAULTS.get(self.kind, (0.01, 1, 1.0, 1.0))
        flags = (L.F_ADD_NOISE if add_noise else 0) | (L.F_EVAL_MODE if eval_mode else 0) | \
                (L.F_ADD_FILTER if add_filter else 0) | (L.F_AUTORESET if autoreset else 0) | \
                (L.F_OBS_F64 if obs_f64 else 0)
        if max_episode_steps is None:
            max_episode_steps = self.layout.default_max_episode_steps
        self.cfg = L.Config(
            abi_version=L.CL_ABI_VERSION, kind=self.kind, device=dev_index, flags=flags,
            num_envs=self.num_envs, n_pad=self.n_pad, env_id_base=int(env_id_base), seed=int(seed),
            max_episode_steps=int(max_episode_steps),
            substeps=int(substeps if substeps is not None else d_sub),
            dt=float(dt if dt is not None else d_dt), alpha=float(alpha),
            act_limit=float(act_limit if act_limit is not None else d_lim),
            act_gain=float(act_gain if act_gain is not None else d_gain),
            param_jitter=float(param_jitter))
        self.obs_f64 = bool(obs_f64)
        self.autoreset = bool(autoreset)
        self.real = torch.float64 if self.layout.real_bytes == 8 else torch.float32
        self.obs_dim, self.act_dim = self.layout.obs_dim, self.layout.act_dim
        dev, NP = self.device, self.n_pad
        with torch.cuda.device(dev):
            self.state = torch.zeros((self.layout.n_state, NP), dtype=self.real, device=dev)
            self.aux_int = torch.zeros((max(self.layout.n_int, 1), NP), dtype=torch.int32, device=dev)
            self.ep_len = torch.zeros(NP, dtype=torch.int32, device=dev)
            self.ep_return = torch.zeros(NP, dtype=torch.float64, device=dev)
            self.stats_buf = torch.zeros(L.NSTATS, dtype=torch.float64, device=dev)
            odt = torch.float64 if obs_f64 else torch.float32
            self.obs_planes = torch.zeros((self.obs_dim, NP), dtype=odt, device=dev)
            self.term_obs_planes = torch.zeros((self.obs_dim, NP), dtype=odt, device=dev)
            self.reward_buf = torch.zeros(NP, dtype=self.real, device=dev)
            self.done_buf = torch.zeros(NP, dtype=torch.uint8, device=dev)
            self.last_ep_ret = torch.zeros(NP, dtype=torch.float64, device=dev)
            self.last_ep_len = torch.zeros(NP, dtype=torch.int32, device=dev)
        self._bufs = L.Buffers(_ptr(self.state), _ptr(self.aux_int), _ptr(self.ep_len),
                               _ptr(self.ep_return), _ptr(self.stats_buf))
        ctx = C.c_void_p()
        L.check(self.lib.cl_create(C.byref(self.cfg), C.byref(ctx)), None, "cl_create")
        self.ctx = ctx
        self._host_views = {}
        self._pin_view = None
        self._host_view = L.HostView()
        self._host_view_ref = C.byref(self._host_view)
        # cached per-call objects of the hot `step()` path (views alias the static planes)
        N = self.num_envs
        self._obs_view = self._view(self.obs_planes)
        self._rew_view = self.reward_buf[:N]
        self._done_view = self.done_buf[:N]
        self._io_step = self._io()
        self._bufs_ref = C.byref(self._bufs)
        self._io_step_ref = C.byref(self._io_step)
        L.check(self.lib.cl_init_persistent(self.ctx, self._stream(), C.byref(self._bufs)),
                self.ctx, "cl_init_persistent")

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        # torch's current stream on this device, as the raw cudaStream_t.  The private accessor returns the
        # integer directly (0.1 us); the public route builds a Stream object per call (2 us, twice per host step)
        if _RAW_STREAM is not None:
            return _RAW_STREAM(self._dev_index)
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        if getattr(self, "ctx", None):
            torch.cuda.synchronize(self.device)
            self.lib.cl_destroy(self.ctx)
            self.ctx = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def block_size(self) -> int:
        return self.lib.cl_block_size(self.ctx)

    @property
    def dyn_launch_count(self) -> int:
        return self.lib.cl_dyn_launch_count(self.ctx)

    @property
    def plain_launch_count(self) -> int:
        """Rollouts that ran on the plain-I/O kernel instantiation (FP64-bound kinds)."""
        return self.lib.cl_plain_launch_count(self.ctx)

    @property
    def sm_launch_count(self) -> int:
        """Rollouts that ran on the SM-local kernel (k_rollout_sm)."""
        return self.lib.cl_sm_launch_count(self.ctx)

    @property
    def launch_count(self) -> int:
        return self.lib.cl_launch_count(self.ctx)

    @property
    def step_index(self) -> int:
        v = C.c_uint64()
        self.lib.cl_get_step_index(self.ctx, C.byref(v))
        return v.value

    @step_index.setter
    def step_index(self, v: int):
        self.lib.cl_set_step_index(self.ctx, int(v))

    def set_graph_mode(self, enable: bool = True) -> None:
        """Keep the Philox step index on the device so that step/rollout/reset launches can be
        captured in a CUDA graph (torch.cuda.graph) and replayed."""
        L.check(self.lib.cl_set_graph_mode(self.ctx, int(bool(enable))), self.ctx, "cl_set_graph_mode")

    def _view(self, planes: torch.Tensor) -> torch.Tensor:
        """[C][n_pad] planes -> zero-copy [N, C] view (strides (1, n_pad))."""
        return planes[:, : self.num_envs].t()

    def _io(self, action=None, noise=None, mask=None, obs=None, reward=None, done=None,
            term_obs=None) -> L.IO:
        io = L.IO()
        if action is not None:
            io.action, io.act_es, io.act_cs = _ptr(action), action.stride(0), action.stride(1)
        io.noise = _ptr(noise)
        obs = self.obs_planes if obs is None else obs
        io.obs, io.obs_es, io.obs_cs = _ptr(obs), 1, self.n_pad
        io.reward = _ptr(self.reward_buf if reward is None else reward)
        io.done = _ptr(self.done_buf if done is None else done)
        io.term_obs = _ptr(self.term_obs_planes if term_obs is None else term_obs)
        io.last_ep_ret, io.last_ep_len = _ptr(self.last_ep_ret), _ptr(self.last_ep_len)
        io.mask = _ptr(mask)
        return io

    def _check_buffer(self, name: str, t, shape, dtype) -> None:
        """Caller-supplied device buffers reach the kernels as raw pointers: refuse anything whose
        shape / dtype / device / layout differs from what the kernel will address."""
        if not isinstance(t, torch.Tensor):
            raise ValueError(f"{name} must be a torch tensor")
        if t.device != self.device:
            raise ValueError(f"{name} must live on {self.device}, got {t.device}")
        if t.dtype != dtype:
            raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")

    def _check_action(self, actions: torch.Tensor) -> torch.Tensor:
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions, np.float32))
        if actions.device != self.device or actions.dtype != torch.float32:
            actions = actions.to(self.device, torch.float32, non_blocking=True)
        if actions.shape != (self.num_envs, self.act_dim):
            raise ValueError(f"actions must be [{self.num_envs}, {self.act_dim}], got {tuple(actions.shape)}")
        return actions

    # ------------------------------------------------------------------ device API
    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reset all envs (or those where mask != 0).  Returns obs [N, obs_dim] (device view)."""
        if mask is not None:
            m = torch.zeros(self.n_pad, dtype=torch.uint8, device=self.device)
            m[: self.num_envs] = mask.to(self.device).to(torch.uint8)
            mask = m
        io = self._io(mask=mask)
        L.check(self.lib.cl_reset(self.ctx, self._stream(), C.byref(self._bufs), C.byref(io)),
                self.ctx, "cl_reset")
        self._keep = mask
        return self._view(self.obs_planes)

    def step(self, actions: torch.Tensor, noise: Optional[torch.Tensor] = None):
        """One control interval.  actions: f32 [N, act_dim] device tensor, any strides.

        Returns (obs [N, obs_dim], reward [N], done_flags u8 [N]) -- device views of internal
        buffers, overwritten by the next call.  done_flags: bit0 terminated, bit1 truncated;
        where set (and autoreset is on) obs is already the reset observation and the terminal
        observation is in `terminal_obs()`, SB3 DummyVecEnv.step_wait style.
        `noise` (f64 [noise_dim, n_pad] standard normals) overrides the Philox process noise.
        """
        if not (isinstance(actions, torch.Tensor) and actions.dtype == torch.float32 and actions.device == self.device
                and actions.dim() == 2 and actions.shape[0] == self.num_envs and actions.shape[1] == self.act_dim):
            actions = self._check_action(actions)
        io = self._io_step
        io.action, io.act_es, io.act_cs = actions.data_ptr(), actions.stride(0), actions.stride(1)
        if noise is not None:
            self._check_buffer("noise", noise, (self.layout.noise_dim, self.n_pad), torch.float64)
        io.noise = None if noise is None else noise.data_ptr()
        rc = self.lib.cl_step(self.ctx, self._stream(), self._bufs_ref, self._io_step_ref)
        if rc != 0:
            L.check(rc, self.ctx, "cl_step")
        self._keep_step = (actions, noise)
        return self._obs_view, self._rew_view, self._done_view

    def terminal_obs(self) -> torch.Tensor:
        return self._view(self.term_obs_planes)

    def rollout(self, T: int, actions: Optional[torch.Tensor] = None, *, synth_amp: Optional[float] = None,
                out: Optional[Dict[str, torch.Tensor]] = None, want=("obs", "reward", "done"),
                obs_layout: str = "planes"):
        """T fused control intervals in ONE launch (state stays in registers).

        actions: f32 [T, N, act_dim] (any strides) or None -> in-kernel Philox
        U(-synth_amp, synth_amp) (default: the action-space bound).  Outputs are time-major:
        reward [T, n_pad], done [T, n_pad] and obs either as SoA planes [T, obs_dim, n_pad]
        (`obs_layout="planes"`) or as policy-shaped rows [T, N, obs_dim] (`obs_layout="rows"`,
        float32 only; written with warp-transposed 16-byte stores).  Pass `out` to reuse buffers,
        `want` to skip streams.
        """
        T = int(T)
        dev, NP = self.device, self.n_pad
        if out is None:
            out = {}
        rows = obs_layout == "rows"
        if rows and self.obs_f64:
            raise ValueError("obs_layout='rows' is float32 only")
        odt = torch.float64 if self.obs_f64 else torch.float32
        if "obs" in want and "obs" not in out:
            shape = (T, self.num_envs, self.obs_dim) if rows else (T, self.obs_dim, NP)
            out["obs"] = torch.empty(shape, dtype=odt, device=dev)
        if "reward" in want and "reward" not in out:
            out["reward"] = torch.empty((T, NP), dtype=self.real, device=dev)
        if "done" in want and "done" not in out:
            out["done"] = torch.empty((T, NP), dtype=torch.uint8, device=dev)
        io = L.IO()
        desc = L.RolloutDesc(T=T, synth_amp=float(self.layout.act_high if synth_amp is None else synth_amp))
        if actions is not None:
            if actions.device != dev or actions.dtype != torch.float32 or \
                    tuple(actions.shape) != (T, self.num_envs, self.act_dim):
                raise ValueError(f"actions must be f32 [{T}, {self.num_envs}, {self.act_dim}] on {dev}")
            io.action, desc.act_ts, io.act_es, io.act_cs = _ptr(actions), actions.stride(0), \
                actions.stride(1), actions.stride(2)
        if "obs" in want:
            o = out["obs"]
            if rows:
                self._check_buffer("out['obs'] (obs_layout='rows')", o, (T, self.num_envs, self.obs_dim), odt)
                io.obs, io.obs_es, io.obs_cs, desc.obs_ts = _ptr(o), self.obs_dim, 1, self.num_envs * self.obs_dim
            else:
                self._check_buffer("out['obs']", o, (T, self.obs_dim, NP), odt)
                io.obs, io.obs_es, io.obs_cs, desc.obs_ts = _ptr(o), 1, NP, self.obs_dim * NP
        if "reward" in want:
            self._check_buffer("out['reward']", out["reward"], (T, NP), self.real)
            io.reward, desc.rew_ts = _ptr(out["reward"]), NP
        if "done" in want:
            self._check_buffer("out['done']", out["done"], (T, NP), torch.uint8)
            io.done, desc.done_ts = _ptr(out["done"]), NP
        io.last_ep_ret, io.last_ep_len = _ptr(self.last_ep_ret), _ptr(self.last_ep_len)
        L.check(self.lib.cl_rollout(self.ctx, self._stream(), C.byref(self._bufs), C.byref(io),
                                    C.byref(desc)), self.ctx, "cl_rollout")
        return out

    def derivatives(self, state: torch.Tensor, action: Optional[torch.Tensor] = None) -> torch.Tensor:
        """RHS of one system for SoA `state` [dim, n] (reference `_get_derivatives`)."""
        state = state.to(self.device, self.real).contiguous()
        n = state.shape[1]
        if action is not None:
            action = action.to(self.device, torch.float32).contiguous()
        out = torch.empty_like(state)
        L.check(self.lib.cl_derivatives(self.ctx, self._stream(), _ptr(state), _ptr(action), _ptr(out), n),
                self.ctx, "cl_derivatives")
        return out

    def stats(self, clear: bool = False) -> Dict[str, float]:
        out = torch.empty(L.NSTATS, dtype=torch.float64, device=self.device)
        L.check(self.lib.cl_stats(self.ctx, self._stream(), C.byref(self._bufs), _ptr(out), int(clear)),
                self.ctx, "cl_stats")
        v = out.cpu().tolist()
        return dict(zip(L.STAT_NAMES, v))

    def stats_tensor(self, clear: bool = False) -> torch.Tensor:
        """Device copy of the 8-double statistics vector (input of the NCCL all-reduce)."""
        out = torch.empty(L.NSTATS, dtype=torch.float64, device=self.device)
        L.check(self.lib.cl_stats(self.ctx, self._stream(), C.byref(self._bufs), _ptr(out), int(clear)),
                self.ctx, "cl_stats")
        return out

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self) -> Dict[str, object]:
        return {
            "kind": self.kind_name, "num_envs": self.num_envs, "step_index": self.step_index,
            "state": self.state.clone(), "aux_int": self.aux_int.clone(),
            "ep_len": self.ep_len.clone(), "ep_return": self.ep_return.clone(),
            "stats": self.stats_buf.clone(),
        }

    def load_state_dict(self, sd: Dict[str, object]) -> None:
        if sd["kind"] != self.kind_name or sd["num_envs"] != self.num_envs:
            raise ValueError("state_dict does not match this batch")
        self.state.copy_(sd["state"]); self.aux_int.copy_(sd["aux_int"])
        self.ep_len.copy_(sd["ep_len"]); self.ep_return.copy_(sd["ep_return"])
        self.stats_buf.copy_(sd["stats"])
        self.step_index = int(sd["step_index"])

    # ------------------------------------------------------------------ host (numpy) path
    def _host_stream(self):
        return self._stream()

    def host_action_buffer(self) -> np.ndarray:
        """Pinned f32 [N, act_dim] staging area; write actions here to skip one host copy."""
        if self._pin_view is None:
            p = C.c_void_p()
            L.check(self.lib.cl_host_action_staging(self.ctx, C.byref(p)), self.ctx, "cl_host_action_staging")
            arr = (C.c_float * (self.num_envs * self.act_dim)).from_address(p.value)
            self._pin_view = np.ctypeslib.as_array(arr).reshape(self.num_envs, self.act_dim)
        return self._pin_view

    def step_host_async(self, actions: Optional[np.ndarray]) -> None:
        """SB3 step_async: f32 [N, act_dim] host actions -> (H2D +) kernel (+ D2H), enqueued.
        Passing the array returned by `host_action_buffer()` (or None after writing into it)
        skips the user->pinned staging copy."""
        if actions is None or actions is self._pin_view:
            ap = None
        else:
            a = actions
            if not (type(a) is np.ndarray and a.dtype == np.float32 and a.flags.c_contiguous):
                a = np.ascontiguousarray(actions, dtype=np.float32)
            if a.shape != (self.num_envs, self.act_dim):
                raise ValueError(f"actions must be [{self.num_envs}, {self.act_dim}], got {a.shape}")
            ap = a.__array_interface__["data"][0]
            self._keep_act = a
        rc = self.lib.cl_step_host_async(self.ctx, self._host_stream(), self._bufs_ref, ap)
        if rc != 0:
            L.check(rc, self.ctx, "cl_step_host_async")

    def step_host_wait(self):
        """SB3 step_wait: blocks, returns zero-copy numpy views of the pinned result slot
        (valid for the next 2 steps): obs f32 [N, obs_dim], reward f32 [N], done u8 [N],
        term_obs, last_ep_ret, last_ep_len, n_done."""
        v = self._host_view
        rc = self.lib.cl_step_host_wait_view(self.ctx, self._host_stream(), self._host_view_ref)
        if rc != 0:
            L.check(rc, self.ctx, "cl_step_host_wait_view")
        key = v.obs
        views = self._host_views.get(key)
        if views is None:
            N, O = self.num_envs, self.obs_dim

            def arr(ptr, ctype, shape):
                n = int(np.prod(shape))
                return np.ctypeslib.as_array((ctype * n).from_address(ptr)).reshape(shape)

            views = (arr(v.obs, C.c_float, (N, O)), arr(v.reward, C.c_float, (N,)),
                     arr(v.done, C.c_uint8, (N,)), arr(v.term_obs, C.c_float, (N, O)),
                     arr(v.last_ep_ret, C.c_double, (N,)), arr(v.last_ep_len, C.c_int32, (N,)))
            self._host_views[key] = views
        return (*views, v.n_done)

    def set_host_mode(self, mode: str, slices: int = 1) -> None:
        """How `step_host_async` moves the data: "dma", "zerocopy", "pipelined" (`slices` env slices
        alternating over two streams) or "streamed" (one launch issued before the caller's array is
        staged slice by slice into the pinned buffer the kernel reads).  Results do not depend on it."""
        m = {"dma": L.HOST_DMA, "zerocopy": L.HOST_ZEROCOPY, "pipelined": L.HOST_PIPELINED,
             "streamed": L.HOST_STREAMED}[mode]
        L.check(self.lib.cl_host_set_mode(self.ctx, m, int(slices)), self.ctx, "cl_host_set_mode")

    @property
    def streamed_fallbacks(self) -> int:
        """Streamed host steps that were called off and redone as zero-copy steps (synchronous launches)."""
        return self.lib.cl_host_streamed_fallbacks(self.ctx)

    def reset_host(self) -> np.ndarray:
        obs = np.empty((self.num_envs, self.obs_dim), np.float32)
        L.check(self.lib.cl_reset_host(self.ctx, self._host_stream(), C.byref(self._bufs),
                                       C.c_void_p(obs.ctypes.data)), self.ctx, "cl_reset_host")
        return obs

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.lib.cl_host_h2d_bytes(self.ctx)

    @property
    def d2h_bytes_per_step(self) -> int:
        return self.lib.cl_host_d2h_bytes(self.ctx)


def _install_nvtx_ranges() -> None:
    """CHAOS_B200_NVTX=1: every ChaosBatch entry point that launches work runs inside an NVTX range
    ("chaos.step", "chaos.rollout", "chaos.step_host_async", ...), so a profile can be cut per call
    (`ncu --nvtx --nvtx-include "chaos.rollout/" ...`, Nsight Systems timelines).  Off by default: the
    wrappers cost ~1 us per call."""
    def ranged(fn, name):
        def wrapper(*a, **k):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
        return wrapper
    for name in ("reset", "step", "rollout", "derivatives", "step_host_async", "step_host_wait", "reset_host"):
        setattr(ChaosBatch, name, ranged(getattr(ChaosBatch, name), "chaos." + name))


if os.environ.get("CHAOS_B200_NVTX") == "1":
    _install_nvtx_ranges()


def measure_fma_peak(device: int = 0, dtype_bytes: int = 8, seconds: float = 0.5) -> float:
    """TFLOP/s of a register-resident FMA chain (2 flop per FMA) on `device`."""
    out = C.c_double()
    L.check(L.load().cl_measure_fma_peak(device, dtype_bytes, seconds, C.byref(out)), None,
            "cl_measure_fma_peak")
    return out.value
