"""Shared test helpers: golden loading, exact/relative comparison, GPU drivers."""
from __future__ import annotations

import ast
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

PARITY_CASES = [
    "lorenz3", "lorenz3_pair", "lorenz4_pair", "hr_sync", "hr_sync_filter", "hr_sync_noise",
    "hr_sync_diverge", "pmsm_sync_a050", "pmsm_sync_a033_noise", "pmsm_sync_diverge",
    "pmsm_classic", "pmsm_single", "memristive4_pair", "pmsm_free",
]
# Per-interval tolerances.  f64 kinds: the north_star bar (1e-12 relative).  PMSM_SYNC is a
# float32 env: states are pure f32 mul/add (bit-exact); reward / lambda / v_t go through
# powf, where libm and the device differ by <= 1 ulp(f32) in rare cases -> 2.5e-7 relative.
RTOL = {"pmsm_sync": 2.5e-7}


def rtol_for(kind):
    return RTOL.get(kind, 1e-12)


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"parity_{name}.npz"))
    d = {k: z[k] for k in z.files}
    d["kind"] = str(d["kind"])
    d["kwargs"] = ast.literal_eval(str(d["kwargs"]))
    return d


def same_nonfinite(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return (np.array_equal(np.isnan(a), np.isnan(b)) and
            np.array_equal(np.isposinf(a), np.isposinf(b)) and
            np.array_equal(np.isneginf(a), np.isneginf(b)))


def max_rel(a, b, floor=1e-300):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    m = np.isfinite(a) & np.isfinite(b)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), floor)))


def assert_close(a, b, rtol, what="", atol=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert same_nonfinite(a, b), f"{what}: inf/nan pattern differs"
    m = np.isfinite(a) & np.isfinite(b)
    err = np.abs(a[m] - b[m])
    tol = atol + rtol * np.abs(b[m])
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.size} beyond rtol={rtol:g}; "
                           f"max rel {max_rel(a, b):.3e}")


def flags_from_kwargs(L, kw):
    return dict(add_noise=bool(kw.get("add_noise", False)), eval_mode=bool(kw.get("eval_mode", False)),
                add_filter=bool(kw.get("add_filter", False)), alpha=float(kw.get("alpha", 0.5)))


# ---- GPU drivers (import torch lazily) ---------------------------------------------------

def gpu_batch(kind, n, **kw):
    from gym_lorenz_b200.core import ChaosBatch
    return ChaosBatch(kind, n, device="cuda:0", **kw)


def gpu_set_state(batch, st, adam=None, ep_len=None):
    """st: [n, n_state] array in the reference's plane order."""
    import torch
    n = st.shape[0]
    batch.state[:, :n] = torch.as_tensor(np.ascontiguousarray(st.T), dtype=batch.real, device=batch.device)
    if adam is not None:
        batch.aux_int[0, :n] = torch.as_tensor(np.asarray(adam, np.int32), device=batch.device)
    if ep_len is not None:
        batch.ep_len[:n] = torch.as_tensor(np.asarray(ep_len, np.int32), device=batch.device)


def gpu_get_state(batch, n=None):
    n = batch.num_envs if n is None else n
    return batch.state[:, :n].t().double().cpu().numpy()


def gpu_noise(batch, nz):
    """nz: [n, 3] standard normals -> f64 [3, n_pad] device tensor."""
    import torch
    t = torch.zeros((3, batch.n_pad), dtype=torch.float64, device=batch.device)
    t[:, : nz.shape[0]] = torch.as_tensor(np.ascontiguousarray(nz.T), device=batch.device)
    return t


F32_OBS_KINDS = ("hr_sync", "pmsm_sync")  # gymnasium envs return float32 observations


def gpu_free_run(case, obs_f64=None):
    """Replay a golden case on the GPU: K envs, T free-running steps (no auto-reset)."""
    import torch
    kind, kw = case["kind"], case["kwargs"]
    obs_f64 = (kind not in F32_OBS_KINDS) if obs_f64 is None else obs_f64
    K, T = case["actions"].shape[:2]
    b = gpu_batch(kind, K, autoreset=False, max_episode_steps=0, obs_f64=obs_f64,
                  **flags_from_kwargs(None, kw))
    gpu_set_state(b, case["st0"], adam=case["adam0"] if kind == "pmsm_sync" else None)
    noisy = bool(np.any(case["noise"] != 0))
    S, OB, RW, DN = [], [], [], []
    for t in range(T):
        a = torch.as_tensor(case["actions"][:, t], device=b.device)
        nz = gpu_noise(b, case["noise"][:, t]) if noisy else None
        obs, rew, done = b.step(a, noise=nz)
        S.append(gpu_get_state(b)); OB.append(obs.double().cpu().numpy())
        RW.append(rew.double().cpu().numpy()); DN.append(done.cpu().numpy())
    b.close()
    return {"state": np.stack(S, 1), "obs": np.stack(OB, 1), "reward": np.stack(RW, 1),
            "done": np.stack(DN, 1)}


def gpu_teacher_forced(case, obs_f64=None):
    """Every (env k, step t) as its own env started from the reference's state at t-1."""
    import torch
    kind, kw = case["kind"], case["kwargs"]
    obs_f64 = (kind not in F32_OBS_KINDS) if obs_f64 is None else obs_f64
    K, T = case["actions"].shape[:2]
    prev = np.concatenate([case["st0"][:, None, :], case["state"][:, :-1, :]], axis=1)  # [K,T,ns]
    n = K * T
    b = gpu_batch(kind, n, autoreset=False, max_episode_steps=0, obs_f64=obs_f64,
                  **flags_from_kwargs(None, kw))
    adam = None
    if kind == "pmsm_sync":
        adam = (case["adam0"][:, None] + np.arange(T)[None, :]).reshape(-1)
    gpu_set_state(b, prev.reshape(n, -1), adam=adam)
    a = torch.as_tensor(case["actions"].reshape(n, -1), device=b.device)
    noisy = bool(np.any(case["noise"] != 0))
    nz = gpu_noise(b, case["noise"].reshape(n, 3)) if noisy else None
    obs, rew, done = b.step(a, noise=nz)
    out = {"state": gpu_get_state(b).reshape(K, T, -1), "obs": obs.double().cpu().numpy().reshape(K, T, -1),
           "reward": rew.double().cpu().numpy().reshape(K, T), "done": done.cpu().numpy().reshape(K, T)}
    b.close()
    return out
