"""Chaotic key-stream trajectories, batched (last item).

The reference's image-encryption demo draws its key material from ONE long trajectory of the 4-D pair env
(code/chaos_apl/main.py:1270-1309 `generate(num)`): step the env `num` times, ignore `done` (no reset), and from
step index 1500 on record `get_current()..get_current3()` = components 0..3 of `state1` and of `state2`, eight
float64 sequences that are then quantised (`np.mod(np.round(x * 10**k), 8) + 1`, main.py:479-480,930-937,1110-1113;
`np.mod(np.round(K * 10**4), 256)`, main.py:230).  Here `n_streams` independent trajectories are generated at
once by the env step kernel (cl_step through ChaosBatch, auto-reset off), `chunk` steps per CUDA-graph replay with
the state planes copied into the trace after every step -- no host round trip until the caller asks for one.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from .core import ChaosBatch

# kinds whose state planes 0..3 / 4..7 are state1 / state2 of a 4-D pair (get_current3 exists)
PAIR_KINDS = ("lorenz4_pair", "memristive4_pair")


def generate(num: int, *, n_streams: int = 1, kind: str = "lorenz4_pair", burn_in: int = 1500, seed: int = 0,
             device="cuda:0", state0: Optional[np.ndarray] = None,
             actions: Optional[Callable[[int], torch.Tensor]] = None, chunk: int = 64,
             use_cuda_graph: bool = True) -> torch.Tensor:
    """Batched `generate(num)`: float64 device tensor [8, num - burn_in, n_streams]; row k is the reference's
    `list_obs{k+1}` (rows 0..3 = state1[0..3], rows 4..7 = state2[0..3]) for each stream.

    `state0` [n_streams, 8] overrides the Philox reset draw (`env.reset()` in the reference); `actions(i)` returns
    the f32 [n_streams, act_dim] device tensor applied at step i (default zeros: `lorenz4_pair` ignores its action,
    lorenz_env_transient.py:316-318; with a callable the steps run eagerly).  Episodes never end (the reference
    loop ignores `dones`)."""
    if kind not in PAIR_KINDS:
        raise ValueError(f"kind must be one of {PAIR_KINDS}")
    num, burn_in, n = int(num), int(burn_in), int(n_streams)
    if num <= burn_in:
        raise ValueError("num must exceed burn_in")
    b = ChaosBatch(kind, n, device=device, seed=seed, autoreset=False, max_episode_steps=0)
    try:
        b.reset()
        if state0 is not None:
            st = np.asarray(state0, np.float64)
            if st.shape != (n, 8):
                raise ValueError(f"state0 must be [{n}, 8]")
            b.state[:8, :n] = torch.as_tensor(np.ascontiguousarray(st.T), device=b.device)
            b.state[8, :n] = 0.0
        dev = b.device
        out = torch.empty((8, num - burn_in, n), dtype=torch.float64, device=dev)
        zero = torch.zeros((n, b.act_dim), dtype=torch.float32, device=dev)
        if actions is not None or not use_cuda_graph:
            for i in range(num):
                b.step(zero if actions is None else actions(i))
                if i >= burn_in:
                    out[:, i - burn_in, :] = b.state[:8, :n]
            return out
        chunk = max(1, min(int(chunk), num))
        trace = torch.empty((chunk, 8, b.n_pad), dtype=torch.float64, device=dev)
        b.set_graph_mode(True)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):               # warm-up outside capture, undone afterwards
            sd = b.state_dict()
            b.step(zero)
            trace[0].copy_(b.state[:8])
            b.load_state_dict(sd)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for j in range(chunk):
                b.step(zero)
                trace[j].copy_(b.state[:8])
        done = 0
        while done + chunk <= num:
            graph.replay()
            lo, hi = max(done, burn_in), done + chunk
            if hi > lo:
                out[:, lo - burn_in:hi - burn_in, :] = trace[lo - done:hi - done, :, :n].permute(1, 0, 2)
            done += chunk
        b.set_graph_mode(False)
        for i in range(done, num):                  # tail shorter than a chunk
            b.step(zero)
            if i >= burn_in:
                out[:, i - burn_in, :] = b.state[:8, :n]
        return out
    finally:
        b.close()


def quantize(x: torch.Tensor, scale: float = 1e4, modulus: int = 8, offset: int = 1) -> torch.Tensor:
    """`(np.mod(np.round(x * scale), modulus) + offset).astype(np.uint8)` (code/chaos_apl/main.py:479-480; with
    modulus=256, offset=0: main.py:230) on the tensor's device.  Round-half-to-even and a non-negative remainder,
    like NumPy."""
    r = torch.remainder(torch.round(x.double() * float(scale)), float(modulus)) + float(offset)
    return r.to(torch.uint8)
