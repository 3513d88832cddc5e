"""Helpers shared by the oracle-vs-reference tests and the golden-vector generator.

`drive_reference(kind, ...)` steps the UNMODIFIED reference class (oracle/ref_loader.py)
from an injected state with given actions (and injected noise where the env draws any) and
returns per-step dumps; `drive_oracle(...)` does the same with oracle/chaos_oracle.c.
"""
from __future__ import annotations

import numpy as np

from oracle import api as O
from oracle import ref_loader as R


class _FixedNormal:
    """Replaces the env's RNG so that the reference consumes OUR standard normals."""

    def __init__(self, z):
        self.z = np.asarray(z, np.float64)
        self.k = 0

    def normal(self, loc=0.0, scale=1.0, size=None):
        z = self.z[self.k]
        self.k += 1
        return loc + scale * z

    def uniform(self, *a, **k):  # not used while stepping
        raise RuntimeError("unexpected uniform draw")


def make_reference(kind, **kw):
    return {
        "lorenz3": R.lorenz3, "lorenz3_pair": R.lorenz3_pair, "lorenz4_pair": R.lorenz4_pair,
        "hr_sync": R.hr_sync, "pmsm_sync": R.pmsm_sync, "pmsm_classic": R.pmsm_classic,
        "pmsm_single": R.pmsm_single, "memristive4_pair": R.memristive4_pair, "pmsm_free": R.pmsm_free,
    }[kind](**kw)


def inject(kind, env, st):
    """Put oracle-layout state vector `st` (one env, plane order of the C-ABI) into `env`."""
    st = np.asarray(st)
    if kind == "pmsm_free":
        env.state1 = [float(v) for v in st[:3]]   # a Python list in the reference
        env.state2 = np.array([0, 0, 0, 0, 0, 0])
        env.t = float(st[3])
        env.state0 = [0.0] * 6
    elif kind in ("lorenz3", "pmsm_single"):
        env.state1 = np.array(st[:3], np.float64)
        env.state2 = np.array([0, 0, 0, 0, 0, 0])
        env.t = float(st[3])
        env.state0 = [0.0] * 6
    elif kind == "lorenz3_pair":
        env.state1 = np.array(st[:3], np.float64)
        env.t = float(st[3])
        env.state2 = [float(v) for v in st[4:10]]
        env.state12 = np.array(st[4:7], np.float64)
        env.state0 = [0.0] * 6
    elif kind in ("lorenz4_pair", "memristive4_pair"):
        env.state1 = np.array(st[:4], np.float64)
        # 8-list of np.float64 scalars after reset()/step() (Python floats would turn
        # `float + np.float32` into float32 under NEP 50 -- not what the reference holds)
        env.state2 = [np.float64(v) for v in st[4:8]] + [np.float64(0.0)] * 4
        env.t = float(st[8])
        env.state0 = [0.0] * 8
    elif kind == "hr_sync":
        env.state_master = np.array(st[:3], np.float64)
        env.state_slave = np.array(st[3:6], np.float64)
        env.sigma = float(st[6])
        env.filtered_action = np.array(st[7:9], np.float32)
    elif kind == "pmsm_sync":
        env.state1 = np.array(st[:3], np.float32)
        env.state2 = np.array(st[3:6], np.float32)
        env.lambda_coef = np.float32(st[6])
        env.m_t = np.float32(st[7])
        env.v_t = np.float32(st[8])
    elif kind == "pmsm_classic":
        env.state1 = np.array(st[:3], np.float64)
        env.state2 = np.array([*st[3:6], 0.0, 0.0, 0.0], np.float64)  # 6-array after reset()
        env.t = float(st[6])
        env.state0 = [0.0] * 6
    else:
        raise KeyError(kind)


def extract(kind, env):
    if kind in ("lorenz3", "pmsm_single", "pmsm_free"):
        return np.array([*env.state1[:3], env.t], np.float64)
    if kind == "lorenz3_pair":
        return np.array([*env.state1[:3], env.t, *env.state2[:6]], np.float64)
    if kind in ("lorenz4_pair", "memristive4_pair"):
        return np.array([*env.state1[:4], *env.state2[:4], env.t], np.float64)
    if kind == "hr_sync":
        return np.array([*env.state_master, *env.state_slave, env.sigma, *env.filtered_action], np.float64)
    if kind == "pmsm_sync":
        return np.array([*env.state1, *env.state2, env.lambda_coef, env.m_t, env.v_t], np.float64)
    if kind == "pmsm_classic":
        return np.array([*env.state1[:3], *env.state2[:3], env.t], np.float64)
    raise KeyError(kind)


NOISY = {"pmsm_classic": True}


def uses_noise(kind, kw):
    return kind in ("pmsm_classic", "pmsm_free") or (kind in ("hr_sync", "pmsm_sync") and kw.get("add_noise", False))


def drive_reference(kind, st0, actions, noise=None, adam_step=0, cur_step=0, **kw):
    """Free-running: inject st0 once, then step len(actions) times.  Returns dict of arrays."""
    env = make_reference(kind, **kw)
    inject(kind, env, st0)
    if kind == "pmsm_sync":
        env.adam_step = int(adam_step)
        env.current_step = int(cur_step)
    T = len(actions)
    if noise is not None:
        fz = _FixedNormal(noise)
        if kind == "pmsm_sync":
            env.np_random = fz
        else:
            np_random_state = np.random.normal
            np.random.normal = fz.normal
    elif kind in ("pmsm_sync",):
        env.np_random = np.random.default_rng(0)
    states, obs, rew, done = [], [], [], []
    try:
        for t in range(T):
            if kind == "pmsm_free":
                import contextlib, io
                with contextlib.redirect_stdout(io.StringIO()):  # the reference prints while t <= 1
                    out = env.step()
            else:
                out = env.step(np.asarray(actions[t], np.float32))
            if len(out) == 4:
                o, r, d, _ = out
                flag = 1 if d else 0
            else:
                o, r, term, trunc, _ = out
                flag = (1 if term else 0) | (2 if trunc else 0)
            states.append(extract(kind, env))
            obs.append(np.asarray(o, np.float64))
            rew.append(float(r))
            done.append(flag)
    finally:
        if noise is not None and kind != "pmsm_sync":
            np.random.normal = np_random_state
    res = {"state": np.array(states), "obs": np.array(obs), "reward": np.array(rew),
           "done": np.array(done, np.uint8)}
    if kind == "pmsm_sync":
        res["adam_step"] = env.adam_step
    return res


def oracle_flags(kind, kw):
    f = 0
    if kw.get("add_noise"):
        f |= O.F_ADD_NOISE
    if kw.get("eval_mode"):
        f |= O.F_EVAL_MODE
    if kw.get("add_filter"):
        f |= O.F_ADD_FILTER
    return f


def drive_oracle(kind, st0, actions, noise=None, adam_step=0, cur_step=0, **kw):
    T = len(actions)
    orc = O.Oracle(kind, 1, flags=oracle_flags(kind, kw), alpha=kw.get("alpha", 0.5))
    orc.state[:, 0] = np.asarray(st0, orc.real)
    if kind == "pmsm_sync":
        orc.aux_int[0, 0] = adam_step
        orc.ep_len[0] = cur_step
    states, obs, rew, done = [], [], [], []
    for t in range(T):
        a = np.zeros((orc.act_dim, orc.n_pad), np.float32)
        a[:, 0] = actions[t]
        nz = None
        if noise is not None:
            nz = np.zeros((3, orc.n_pad), np.float64)
            nz[:, 0] = noise[t]
        o, r, d, _ = orc.step(a, nz)
        states.append(orc.state[:, 0].astype(np.float64).copy())
        obs.append(o[:, 0].copy())
        rew.append(r[0])
        done.append(d[0])
    res = {"state": np.array(states), "obs": np.array(obs), "reward": np.array(rew),
           "done": np.array(done, np.uint8)}
    if kind == "pmsm_sync":
        res["adam_step"] = int(orc.aux_int[0, 0])
    return res
