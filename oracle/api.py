"""ctypes front-end of oracle/chaos_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It mirrors the buffer layout of the C-ABI (SoA planes [c][n_pad]) so the
same seeded inputs can be fed to both sides.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libchaos_oracle.so")

KINDS = {
    "lorenz3": 0, "lorenz3_pair": 1, "lorenz4_pair": 2, "hr_sync": 3, "pmsm_sync": 4,
    "pmsm_classic": 5, "pmsm_single": 6, "lorenz_rk4": 7, "lorenz_rk4_f32": 8, "pmsm_rk4": 9,
    "memristive4_pair": 10, "pmsm_free": 11,
}
F_ADD_NOISE, F_EVAL_MODE, F_ADD_FILTER, F_AUTORESET = 1, 2, 4, 8


class OrcCfg(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("flags", C.c_int32), ("max_episode_steps", C.c_int32),
        ("substeps", C.c_int32), ("n", C.c_int64), ("n_pad", C.c_int64),
        ("env_id_base", C.c_int64), ("seed", C.c_uint64), ("step_index", C.c_uint64),
        ("dt", C.c_double), ("alpha", C.c_double), ("act_limit", C.c_double),
        ("act_gain", C.c_double), ("param_jitter", C.c_double),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "chaos_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libchaos_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_n_state.restype = C.c_int
        _lib.orc_obs_dim.restype = C.c_int
        _lib.orc_act_dim.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """A batch of `n` oracle envs of one kind, laid out like the CUDA side."""

    def __init__(self, kind, n, *, flags=0, max_episode_steps=0, substeps=1, dt=0.01, alpha=0.5,
                 act_limit=1.0, act_gain=1.0, param_jitter=0.0, seed=0, env_id_base=0, n_pad=None):
        self.kind = KINDS[kind] if isinstance(kind, str) else int(kind)
        self.n = int(n)
        self.n_pad = int(n_pad) if n_pad is not None else ((self.n + 127) // 128) * 128
        L = lib()
        self.n_state = L.orc_n_state(self.kind)
        self.obs_dim = L.orc_obs_dim(self.kind)
        self.act_dim = L.orc_act_dim(self.kind)
        self.real = np.float32 if self.kind in (4, 8) else np.float64
        self.cfg = OrcCfg(self.kind, flags, max_episode_steps, substeps, self.n, self.n_pad,
                          env_id_base, seed, 0, dt, alpha, act_limit, act_gain, param_jitter)
        self.state = np.zeros((self.n_state, self.n_pad), self.real)
        self.aux_int = np.zeros((1, self.n_pad), np.int32)
        self.ep_len = np.zeros(self.n_pad, np.int32)
        self.ep_return = np.zeros(self.n_pad, np.float64)
        self.stats = np.zeros(8, np.float64)
        L.orc_init_persistent(C.byref(self.cfg), _p(self.state), _p(self.aux_int), _p(self.ep_len),
                              _p(self.ep_return))

    @property
    def step_index(self):
        return self.cfg.step_index

    @step_index.setter
    def step_index(self, v):
        self.cfg.step_index = int(v)

    def reset(self, mask=None):
        obs = np.zeros((self.obs_dim, self.n_pad), np.float64)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().orc_reset(C.byref(self.cfg), _p(self.state), _p(self.aux_int), _p(self.ep_len),
                        _p(self.ep_return), _p(m), _p(obs))
        self.cfg.step_index += 1
        return obs

    def rollout(self, T, action=None, noise=None, synth_amp=1.0, want_term=True):
        """action: f32 [T][act_dim][n_pad] or None (Philox synthetic).  Returns a dict."""
        T = int(T)
        if action is not None:
            action = np.ascontiguousarray(action, np.float32)
            assert action.shape == (T, self.act_dim, self.n_pad), action.shape
        if noise is not None:
            noise = np.ascontiguousarray(noise, np.float64)
            assert noise.shape == (3, self.n_pad)
        out = {
            "obs": np.zeros((T, self.obs_dim, self.n_pad), np.float64),
            "reward": np.zeros((T, self.n_pad), np.float64),
            "done": np.zeros((T, self.n_pad), np.uint8),
            "term_obs": np.zeros((T, self.obs_dim, self.n_pad), np.float64) if want_term else None,
            "last_ep_ret": np.zeros(self.n_pad, np.float64),
            "last_ep_len": np.zeros(self.n_pad, np.int32),
        }
        lib().orc_rollout(C.byref(self.cfg), C.c_int(T), C.c_double(synth_amp), _p(self.state),
                          _p(self.aux_int), _p(self.ep_len), _p(self.ep_return), _p(self.stats),
                          _p(action), _p(noise), _p(out["obs"]), _p(out["reward"]), _p(out["done"]),
                          _p(out["term_obs"]), _p(out["last_ep_ret"]), _p(out["last_ep_len"]))
        self.cfg.step_index += T
        return out

    def step(self, action, noise=None):
        """action: f32 [act_dim][n_pad] (SoA).  Returns obs[obs_dim][n_pad], reward, done."""
        a = np.ascontiguousarray(action, np.float32).reshape(1, self.act_dim, self.n_pad)
        o = self.rollout(1, a, noise)
        return o["obs"][0], o["reward"][0], o["done"][0], o

    def rollout_timed(self, T, synth_amp=1.0):
        """Synthetic-action rollout without output buffers (cpu_baseline timing)."""
        lib().orc_rollout(C.byref(self.cfg), C.c_int(int(T)), C.c_double(synth_amp), _p(self.state),
                          _p(self.aux_int), _p(self.ep_len), _p(self.ep_return), _p(self.stats),
                          None, None, None, None, None, None, None, None)
        self.cfg.step_index += int(T)


def max_threads() -> int:
    return lib().orc_max_threads()


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int(int(n)))
