"""SB3 `VecEnv` surface over a ChaosBatch (the reference-facing plugin boundary).

Mirrors stable_baselines3==2.7.1 `common/vec_env/base_vec_env.py::VecEnv` (un-vendored
third-party; call sites code/train.py:100, code/lorenz_pmsm/train.py:115-118,166-170,
code/lorenz_filter/train.py:109-115): `reset`, `step_async`, `step_wait`, `step`, `close`,
`get_attr`, `set_attr`, `env_method`, `env_is_wrapped`, `seed`, `set_options`, `reset_infos`
and the auto-reset contract of `DummyVecEnv.step_wait` (terminal observation in
`infos[i]["terminal_observation"]`, `"TimeLimit.truncated"`, Monitor-style
`infos[i]["episode"] = {"r", "l", "t"}`).  When stable_baselines3 is importable the class
subclasses its real `VecEnv`, so `isinstance` checks in SB3 algorithms and wrappers
(`VecNormalize`, `VecFrameStack`, `VecMonitor`) pass; otherwise an API-identical ABC is used.
"""
from __future__ import annotations

import time
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .core import ChaosBatch
from .spaces import box_for

try:  # pragma: no cover - SB3 is absent in the build image
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv as _SB3VecEnv  # type: ignore
    HAVE_SB3 = True
except Exception:  # noqa: BLE001
    _SB3VecEnv = None
    HAVE_SB3 = False


class _VecEnvShim(ABC):
    """API-compatible stand-in for SB3's VecEnv base class."""

    def __init__(self, num_envs: int, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space
        self.reset_infos: List[Dict[str, Any]] = [{} for _ in range(num_envs)]
        self._seeds: List[Optional[int]] = [None for _ in range(num_envs)]
        self._options: List[Dict[str, Any]] = [{} for _ in range(num_envs)]
        try:
            render_modes = self.get_attr("render_mode")
        except AttributeError:
            render_modes = [None for _ in range(num_envs)]
        self.render_mode = render_modes[0]

    def _reset_seeds(self) -> None:
        self._seeds = [None for _ in range(self.num_envs)]

    def _reset_options(self) -> None:
        self._options = [{} for _ in range(self.num_envs)]

    @abstractmethod
    def reset(self): ...

    @abstractmethod
    def step_async(self, actions: np.ndarray) -> None: ...

    @abstractmethod
    def step_wait(self): ...

    @abstractmethod
    def close(self) -> None: ...

    @abstractmethod
    def get_attr(self, attr_name: str, indices=None) -> List[Any]: ...

    @abstractmethod
    def set_attr(self, attr_name: str, value: Any, indices=None) -> None: ...

    @abstractmethod
    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs) -> List[Any]: ...

    @abstractmethod
    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]: ...

    def step(self, actions: np.ndarray):
        self.step_async(actions)
        return self.step_wait()

    def get_images(self):
        raise NotImplementedError

    def render(self, mode: Optional[str] = None):
        return None

    def seed(self, seed: Optional[int] = None) -> Sequence[Optional[int]]:
        if seed is None:
            seed = int(np.random.randint(0, np.iinfo(np.uint32).max, dtype=np.uint32))
        self._seeds = [seed + idx for idx in range(self.num_envs)]
        return self._seeds

    def set_options(self, options=None) -> None:
        if options is None:
            options = {}
        if isinstance(options, dict):
            self._options = [dict(options) for _ in range(self.num_envs)]
        else:
            self._options = [dict(o) for o in options]

    @property
    def unwrapped(self):
        return self

    def getattr_depth_check(self, name: str, already_found: bool):
        if hasattr(self, name) and already_found:
            return f"{type(self).__module__}.{type(self).__name__}"
        return None

    def _get_indices(self, indices) -> Sequence[int]:
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices


VecEnv = _SB3VecEnv if HAVE_SB3 else _VecEnvShim

# attribute name -> (plane slice in the state block) per kind; names are the reference's.
ATTR_PLANES = {
    "lorenz3": {"state1": (0, 3), "t": (3, 4)},
    "lorenz3_pair": {"state1": (0, 3), "t": (3, 4), "state2": (4, 10), "state12": (4, 7)},
    "lorenz4_pair": {"state1": (0, 4), "state2": (4, 8), "t": (8, 9)},
    "hr_sync": {"state_master": (0, 3), "state_slave": (3, 6), "sigma": (6, 7), "filtered_action": (7, 9)},
    "pmsm_sync": {"state1": (0, 3), "state2": (3, 6), "lambda_coef": (6, 7), "m_t": (7, 8), "v_t": (8, 9)},
    "pmsm_classic": {"state1": (0, 3), "state2": (3, 6), "t": (6, 7)},
    "pmsm_single": {"state1": (0, 3), "t": (3, 4)},
    "lorenz_rk4": {"state1": (0, 3), "sigma": (3, 4), "rho": (4, 5), "beta": (5, 6)},
    "lorenz_rk4_f32": {"state1": (0, 3), "sigma": (3, 4), "rho": (4, 5), "beta": (5, 6)},
    "pmsm_rk4": {"state1": (0, 3), "state2": (3, 6), "sigma": (6, 7), "gamma": (7, 8)},
    "memristive4_pair": {"state1": (0, 4), "state2": (4, 8), "t": (8, 9)},
    "pmsm_free": {"state1": (0, 3), "t": (3, 4)},
}
_SCALAR_ATTRS = {"t", "sigma", "lambda_coef", "m_t", "v_t", "rho", "beta", "gamma"}
# constructor constants of the reference classes that callers read through get_attr
CONST_ATTRS = {
    "hr_sync": {"dt": 0.001, "scale_factor": 50.0, "action_alpha": 0.95},
    "pmsm_sync": {"sigma": 5.46, "gamma": 20.0, "dt": 0.001, "f_max": 50, "max_steps": 2000},
}


class _LazyInfos(list):
    """The `infos` list of the VecEnv contract (one dict per env).  Dicts of envs that finished an
    episode in the last step (terminal_observation / TimeLimit.truncated / episode) are built on
    first access instead of inside step_wait(): with tens of thousands of envs ending together,
    building them costs more than the step itself, and loops that never look at `infos`
    (random rollouts, custom collectors) should not pay for it.  All read paths materialise
    first, so the object behaves exactly like the eager list."""

    def __init__(self, n):
        super().__init__({} for _ in range(n))
        self._pending = None
        self._dirty = []

    def _begin_step(self, pending):
        setitem = list.__setitem__
        for i in self._dirty:
            setitem(self, i, {})
        self._dirty = []
        self._pending = pending

    def _materialize(self):
        p = self._pending
        if p is None:
            return
        self._pending = None
        ids, tobs, trunc, rets, lens, now = p
        setitem = list.__setitem__
        if rets is not None:
            for i, to, tr, r, ln in zip(ids, tobs, trunc, rets, lens):
                setitem(self, i, {"terminal_observation": to, "TimeLimit.truncated": tr,
                                  "episode": {"r": round(r, 6), "l": ln, "t": now}})  # SB3 Monitor rounds r
        else:
            for i, to, tr in zip(ids, tobs, trunc):
                setitem(self, i, {"terminal_observation": to, "TimeLimit.truncated": tr})
        self._dirty = ids

    def __getitem__(self, k):
        self._materialize()
        return list.__getitem__(self, k)

    def __iter__(self):
        self._materialize()
        return list.__iter__(self)

    def __reversed__(self):
        self._materialize()
        return list.__reversed__(self)

    def __contains__(self, x):
        self._materialize()
        return list.__contains__(self, x)

    def __eq__(self, other):
        self._materialize()
        return list.__eq__(self, other)

    __hash__ = None

    def __repr__(self):
        self._materialize()
        return list.__repr__(self)

    def copy(self):
        self._materialize()
        return list(list.__iter__(self))

    def __reduce__(self):
        self._materialize()
        return (list, (list(list.__iter__(self)),))


class BatchedChaosVecEnv(VecEnv):
    """`num_envs` chaos-control envs on one B200 behind the SB3 VecEnv contract.

    >>> env = BatchedChaosVecEnv("hr_sync", 4096)        # replaces DummyVecEnv([make("lorenz_try-v0")])
    >>> model = PPO("MlpPolicy", env, n_steps=2048, ...)  # code/train.py:112-118 unchanged

    numpy path (SB3): `reset()`, `step_async(a)`, `step_wait()`; host<->device copies go
    through pinned staging inside the C library.  Tensor path: `reset_tensor()`,
    `step_tensor(a)` return device tensors (zero-copy views; DLPack-exportable) with no host
    round trip.
    """

    # obs + reward bytes up to which step_wait() returns private copies by default (a copy of this size
    # costs a few microseconds; at 65,536 envs it would cost more than the step itself)
    COPY_OUTPUTS_AUTO_BYTES = 256 * 1024

    def __init__(self, kind: str = "hr_sync", num_envs: int = 1, *, device="cuda:0", seed: int = 0,
                 monitor: bool = True, copy_outputs: Optional[bool] = None, **kwargs):
        """`copy_outputs`: SB3's DummyVecEnv returns fresh copies of obs / rewards / dones every step.
        True does the same; False returns zero-copy views of a ring of 3 pinned result slots (obs, rewards,
        and `dones` on steps in which no episode ended), valid until the third following step (enough for SB3's own algorithms, which copy what they keep
        at once); None (default) copies while obs + reward are at most 256 KiB per step and aliases
        above.  `infos` is one list object reused across steps either way: entries of envs that
        finished an episode are rebuilt each step, so keep `infos[i]` dicts, not the list."""
        self._kw = dict(kwargs)
        self._kind = kind
        self._device = device
        self._seed0 = int(seed)
        self._monitor = bool(monitor)
        self.batch = ChaosBatch(kind, num_envs, device=device, seed=seed, autoreset=True, **kwargs)
        self._t_start = time.time()
        self._infos = _LazyInfos(num_envs)
        self._waiting = False
        if copy_outputs is None:
            copy_outputs = num_envs * (self.batch.obs_dim + 1) * 4 <= self.COPY_OUTPUTS_AUTO_BYTES
        self._copy_outputs = bool(copy_outputs)
        super().__init__(num_envs, box_for(self.batch.layout, "obs"), box_for(self.batch.layout, "act"))

    # ---- numpy / SB3 path --------------------------------------------------------------
    def reset(self) -> np.ndarray:
        if any(s is not None for s in self._seeds):
            self._reseed(int(self._seeds[0]))
        obs = self.batch.reset_host()
        self.reset_infos = [{} for _ in range(self.num_envs)]
        self._reset_seeds()
        self._reset_options()
        return obs

    def step_async(self, actions: np.ndarray) -> None:
        self.batch.step_host_async(actions)
        self._waiting = True

    def step_wait(self):
        obs, rew, done, term_obs, ler, lel, n_done = self.batch.step_host_wait()
        self._waiting = False
        infos = self._infos
        pending = None
        if not n_done:
            # nobody finished: every flag byte is 0, i.e. already a valid all-False bool array.  Without
            # copies that view IS the result (like obs / rewards it lives in the 3-slot ring); comparing
            # 65,536 bytes would cost 5 us of an 80 us step
            dones = done.view(np.bool_) if not self._copy_outputs else np.zeros(self.num_envs, np.bool_)
        else:
            dones = done != 0
            idx = np.flatnonzero(dones)
            flags = done[idx]
            trunc = (((flags & L.DONE_TRUNCATED) != 0) & ((flags & L.DONE_TERMINATED) == 0)).tolist()
            pending = (idx.tolist(), term_obs[idx], trunc,
                       ler[idx].tolist() if self._monitor else None,
                       lel[idx].tolist() if self._monitor else None,
                       round(time.time() - self._t_start, 6))
        infos._begin_step(pending)
        if self._copy_outputs:
            return obs.copy(), rew.copy(), dones, infos     # `dones` is already a fresh array
        return obs, rew, dones, infos

    def close(self) -> None:
        self.batch.close()

    # ---- tensor path (no host round trip) ----------------------------------------------
    def reset_tensor(self) -> torch.Tensor:
        return self.batch.reset()

    def step_tensor(self, actions: torch.Tensor):
        """Returns (obs [N,obs_dim], reward [N], done_flags u8 [N]) device views."""
        return self.batch.step(actions)

    def obs_dlpack(self):
        return torch.utils.dlpack.to_dlpack(self.batch._view(self.batch.obs_planes))

    # ---- attribute access (code/lorenz_pmsm/test_evaluate.py:100-108,123-125) -----------
    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        idx = list(self._get_indices(indices))
        if attr_name == "render_mode":
            return [None for _ in idx]
        planes = ATTR_PLANES[self.batch.kind_name]
        if attr_name in planes:
            lo, hi = planes[attr_name]
            block = self.batch.state[lo:hi][:, idx].t().cpu().numpy()
            if attr_name in _SCALAR_ATTRS:
                return [block[k, 0].item() for k in range(len(idx))]
            return [block[k].copy() for k in range(len(idx))]
        if attr_name in ("current_step", "ep_len"):
            return self.batch.ep_len[idx].cpu().tolist()
        if attr_name == "adam_step" and self.batch.layout.n_int:
            return self.batch.aux_int[0][idx].cpu().tolist()
        consts = CONST_ATTRS.get(self.batch.kind_name, {})
        if attr_name in consts:
            return [consts[attr_name] for _ in idx]
        if attr_name in self._kw:
            return [self._kw[attr_name] for _ in idx]
        if attr_name in ("observation_space", "action_space"):
            return [getattr(self, attr_name) for _ in idx]
        raise AttributeError(attr_name)

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        idx = list(self._get_indices(indices))
        planes = ATTR_PLANES[self.batch.kind_name]
        dev = self.batch.device
        if attr_name in planes:
            lo, hi = planes[attr_name]
            v = torch.as_tensor(np.asarray(value, np.float64).reshape(-1), dtype=self.batch.real, device=dev)
            if v.numel() != hi - lo:
                raise ValueError(f"{attr_name} expects {hi - lo} values")
            self.batch.state[lo:hi, idx] = v[:, None]
            return
        if attr_name in ("current_step", "ep_len"):
            self.batch.ep_len[idx] = int(value)
            return
        if attr_name == "adam_step" and self.batch.layout.n_int:
            self.batch.aux_int[0, idx] = int(value)
            return
        raise AttributeError(attr_name)

    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs) -> List[Any]:
        idx = list(self._get_indices(indices))
        if method_name in ("_get_derivatives", "get_derivatives"):
            state = np.asarray(method_args[0], np.float64).reshape(-1)
            action = method_args[1] if len(method_args) > 1 else method_kwargs.get("action")
            st = torch.as_tensor(state, dtype=self.batch.real, device=self.batch.device)[:, None]
            ac = None
            if action is not None:
                ac = torch.as_tensor(np.asarray(action, np.float32).reshape(-1), device=self.batch.device)[:, None]
            d = self.batch.derivatives(st.contiguous(), None if ac is None else ac.contiguous())
            out = d[:, 0].cpu().numpy()
            return [out.copy() for _ in idx]
        if method_name == "reset":
            # gymnasium reset of the selected envs only (SB3: env_method("reset", indices=...) calls
            # envs[i].reset() for i in indices): masked device reset, others keep their episodes
            if len(idx) == self.num_envs:
                obs = self.reset()
                return [(obs[i], {}) for i in idx]
            mask = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.batch.device)
            mask[torch.as_tensor(idx, dtype=torch.long, device=self.batch.device)] = 1
            obs = self.batch.reset(mask)[idx].float().cpu().numpy()
            return [(obs[k], {}) for k in range(len(idx))]
        raise AttributeError(method_name)

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False for _ in self._get_indices(indices)]

    # ---- seeding ---------------------------------------------------------------------------
    def _reseed(self, seed: int) -> None:
        old = self.batch
        sd_step = 0
        self.batch = ChaosBatch(self._kind, self.num_envs, device=self._device, seed=seed,
                                autoreset=True, **self._kw)
        self.batch.step_index = sd_step
        old.close()
        self._seed0 = seed

    def stats(self, clear: bool = False):
        return self.batch.stats(clear)
