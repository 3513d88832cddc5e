"""Single-env facades with the reference classes' constructor kwargs, attributes and
return conventions, each a `num_envs == 1` ChaosBatch on the GPU (no CPU arithmetic).

| class here            | reference class (gym_lorenz/envs/...)                         | API        |
|-----------------------|---------------------------------------------------------------|------------|
| lorenzEnv_transient   | dynamic.py:5-93 `lorenzEnv_transient`                         | old gym    |
|   .lorenzEnv_transient| dynamic.py:109-233 nested class (frozen target)               | old gym    |
| Lorenz4PairEnv        | lorenz_env_transient.py:247-376 `lorenzEnv_transient`         | old gym    |
| HRSyncEnv             | lorenz_env_try.py:13-179                                      | gymnasium  |
| PMSM_Sync_Env         | lorenz_env_try_pmsm.py:7-184                                  | gymnasium  |
| PMSMClassicEnv        | lorenz_env_transient_pmsm.py:17-137 `lorenzEnv_transient`     | old gym    |
| PMSMSingleEnv         | lorenz_env_transient1.py `lorenzEnv_transient`                | old gym    |
| LorenzRK4Env          | (north-star) RK4 x S Lorenz targeting env                      | gymnasium  |

For many envs use `BatchedChaosVecEnv`; these facades exist so that scripts written against
one env object (`env.reset()`, `env.step(a)`, `env.state1 = ...`,
code/lorenz_pmsm/test_evaluate.py:100-108) keep working.
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np
import torch

from . import _lib as L
from .core import ChaosBatch
from .spaces import box_for
from .vec_env import ATTR_PLANES, CONST_ATTRS, _SCALAR_ATTRS

try:  # pragma: no cover
    import gymnasium as _gym  # type: ignore
    _EnvBase = _gym.Env
except Exception:  # noqa: BLE001
    _EnvBase = object


class _SingleEnv(_EnvBase):
    KIND = ""
    OLD_GYM = False
    metadata = {"render.modes": ["human", "rgb_array"], "render_modes": []}
    render_mode = None

    def __init__(self, device: Optional[str] = None, seed: int = 0, **kw):
        object.__setattr__(self, "_ready", False)
        self._kw = kw
        self._device = device or "cuda:0"
        self._seed = int(seed)
        self._make(self._seed)
        self.observation_space = box_for(self._batch.layout, "obs")
        self.action_space = box_for(self._batch.layout, "act")
        for k, v in CONST_ATTRS.get(self.KIND, {}).items():
            object.__setattr__(self, k, v)
        for k, v in kw.items():
            object.__setattr__(self, k, v)
        object.__setattr__(self, "_ready", True)

    def _make(self, seed: int) -> None:
        # gymnasium applies TimeLimit outside the env (gym_lorenz/__init__.py:12,20): none here
        self._batch = ChaosBatch(self.KIND, 1, device=self._device, seed=seed, autoreset=False,
                                 max_episode_steps=0, obs_f64=self.OLD_GYM, **self._kw)

    # ---- reference attribute surface (state1/state2/t/..., settable) --------------------
    def __getattr__(self, name: str) -> Any:
        planes = ATTR_PLANES.get(type(self).KIND, {})
        if name in planes and self.__dict__.get("_ready"):
            lo, hi = planes[name]
            v = self._batch.state[lo:hi, 0].cpu().numpy()
            return v[0].item() if name in _SCALAR_ATTRS else v
        if name == "current_step" and self.__dict__.get("_ready"):
            return int(self._batch.ep_len[0].item())
        if name == "adam_step" and self.__dict__.get("_ready"):
            return int(self._batch.aux_int[0, 0].item())
        raise AttributeError(name)

    def __setattr__(self, name: str, value: Any) -> None:
        if self.__dict__.get("_ready"):
            planes = ATTR_PLANES.get(type(self).KIND, {})
            if name in planes:
                lo, hi = planes[name]
                v = torch.as_tensor(np.asarray(value, np.float64).reshape(-1), dtype=self._batch.real,
                                    device=self._batch.device)
                self._batch.state[lo:hi, 0] = v
                return
            if name == "current_step":
                self._batch.ep_len[0] = int(value)
                return
            if name == "adam_step":
                self._batch.aux_int[0, 0] = int(value)
                return
        object.__setattr__(self, name, value)

    # ---- stepping ------------------------------------------------------------------------
    def _obs_np(self, obs_t: torch.Tensor) -> np.ndarray:
        return obs_t[0].cpu().numpy().copy()

    def _step_raw(self, action):
        a = torch.as_tensor(np.asarray(action, np.float32).reshape(1, -1), device=self._batch.device)
        obs, rew, done = self._batch.step(a)
        flag = int(done[0].item())
        return self._obs_np(obs), rew[0].item(), flag

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        if seed is not None and int(seed) != self._seed:
            # re-key Philox; persistent fields (Adam-dual state, per-env params) survive reset
            keep, keep_aux = self._batch.state.clone(), self._batch.aux_int.clone()
            self._batch.close()
            self._seed = int(seed)
            self._make(self._seed)
            self._batch.state.copy_(keep)
            self._batch.aux_int.copy_(keep_aux)
        obs = self._obs_np(self._batch.reset())
        return obs if self.OLD_GYM else (obs, {})

    def step(self, action):
        obs, rew, flag = self._step_raw(action)
        if self.OLD_GYM:
            return obs, np.float64(rew), bool(flag & L.DONE_TERMINATED), {}
        return obs, float(rew), bool(flag & L.DONE_TERMINATED), bool(flag & L.DONE_TRUNCATED), {}

    def render(self, mode: str = "human"):
        return None

    def close(self):
        self._batch.close()

    def _get_observation(self):
        return self._obs_np(self._batch._view(self._batch.obs_planes))


class _OldGymGetters:
    """_get_current* helpers of the classic envs (dynamic.py:52-59)."""

    def _pair(self, k):
        s2 = getattr(self, "state2", None)
        second = float(s2[k]) if s2 is not None else 0.0
        return [float(self.state1[k]), second]

    def _get_current(self):
        return self._pair(0)

    def _get_current1(self):
        return self._pair(1)

    def _get_current2(self):
        return self._pair(2)

    get_current, get_current1, get_current2 = _get_current, _get_current1, _get_current2


class LorenzPairEnv(_OldGymGetters, _SingleEnv):
    KIND, OLD_GYM = "lorenz3_pair", True


class lorenzEnv_transient(_OldGymGetters, _SingleEnv):  # noqa: N801 - reference class name
    KIND, OLD_GYM = "lorenz3", True
    lorenzEnv_transient = LorenzPairEnv  # the reference nests the pair variant as a class attribute

    @property
    def state2(self):
        return np.zeros(6)  # dynamic.py:44


class Lorenz4PairEnv(_OldGymGetters, _SingleEnv):
    KIND, OLD_GYM = "lorenz4_pair", True

    def get_current3(self):
        return self._pair(3)


class PMSMClassicEnv(_OldGymGetters, _SingleEnv):
    KIND, OLD_GYM = "pmsm_classic", True


class PMSMSingleEnv(_OldGymGetters, _SingleEnv):
    KIND, OLD_GYM = "pmsm_single", True

    @property
    def state2(self):
        return np.zeros(6)


class Memristive4PairEnv(_OldGymGetters, _SingleEnv):
    """lorenz_env_transient2.py::lorenzEnv_transient."""
    KIND, OLD_GYM = "memristive4_pair", True

    def get_current3(self):
        return self._pair(3)


class PMSMFreeEnv(_OldGymGetters, _SingleEnv):
    """lorenz_singlecontrol.py::lorenzEnv_transient -- `step()` takes no action."""
    KIND, OLD_GYM = "pmsm_free", True

    @property
    def state2(self):
        return np.zeros(6)

    def step(self, action=None):  # noqa: D401 - reference signature is step(self)
        return super().step(np.zeros(2, np.float32))


class HRSyncEnv(_SingleEnv):
    """lorenz_env_try.py::HRSyncEnv(add_noise=False, eval_mode=False, add_filter=False)."""
    KIND = "hr_sync"

    def __init__(self, add_noise: bool = False, eval_mode: bool = False, add_filter: bool = False, **kw):
        super().__init__(add_noise=add_noise, eval_mode=eval_mode, add_filter=add_filter, **kw)


class PMSM_Sync_Env(_SingleEnv):  # noqa: N801 - reference class name
    """lorenz_env_try_pmsm.py::PMSM_Sync_Env(alpha=0.5, add_noise=False)."""
    KIND = "pmsm_sync"

    def __init__(self, alpha: float = 0.5, add_noise: bool = False, **kw):
        super().__init__(alpha=alpha, add_noise=add_noise, **kw)

    def _get_derivatives(self, state, action, noise=(0, 0, 0)):
        st = torch.as_tensor(np.asarray(state, np.float32).reshape(3, 1), device=self._batch.device)
        ac = torch.as_tensor(np.asarray(action, np.float32).reshape(2, 1), device=self._batch.device)
        if np.any(np.asarray(noise) != 0):
            raise NotImplementedError("noise is injected inside the step kernel; pass noise=0 here")
        return self._batch.derivatives(st, ac)[:, 0].cpu().numpy()


class LorenzRK4Env(_SingleEnv):
    """North-star Lorenz targeting env: RK4 x `substeps` per control interval `dt`."""
    KIND = "lorenz_rk4"

    def __init__(self, substeps: int = 16, dt: float = 0.01, **kw):
        super().__init__(substeps=substeps, dt=dt, **kw)


def register_gymnasium() -> bool:
    """Register the reference's ids (gym_lorenz/__init__.py:4-23) when gymnasium exists."""
    try:  # pragma: no cover
        from gymnasium.envs.registration import register, registry
    except Exception:  # noqa: BLE001
        return False
    specs = {  # pragma: no cover
        "lorenz_try-v0": ("gym_lorenz_b200.envs:HRSyncEnv", 5000),
        "lorenz_pmsm-v0": ("gym_lorenz_b200.envs:PMSM_Sync_Env", 2000),
    }
    for env_id, (entry, steps) in specs.items():  # pragma: no cover
        if env_id not in registry:
            register(id=env_id, entry_point=entry, max_episode_steps=steps, reward_threshold=1e50)
    return True  # pragma: no cover
