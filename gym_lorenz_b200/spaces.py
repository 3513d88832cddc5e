"""`Box` space: gymnasium's when importable (so SB3 accepts it), else a minimal stand-in
with the attributes the reference scripts and SB3-style code read (low/high/shape/dtype,
sample, contains)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is absent in the build image
    from gymnasium.spaces import Box  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Box:  # type: ignore[no-redef]
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(int(s) for s in shape)
            self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else np.asarray(low, self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else np.asarray(high, self.dtype)
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        def __eq__(self, other):
            return isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype \
                and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high)


def box_for(layout, which: str) -> "Box":
    """observation_space / action_space of the reference env for a C-ABI layout."""
    if which == "obs":
        lo = -np.inf if layout.obs_low < -1e300 else layout.obs_low
        hi = np.inf if layout.obs_high > 1e300 else layout.obs_high
        return Box(low=lo, high=hi, shape=(layout.obs_dim,), dtype=np.float32)
    return Box(low=layout.act_low, high=layout.act_high, shape=(layout.act_dim,), dtype=np.float32)
