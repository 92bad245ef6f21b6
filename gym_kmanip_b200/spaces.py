"""Minimal ``gymnasium.spaces`` stand-in (Box, Dict) used when gymnasium is not importable.

The reference declares its observation / action spaces with ``gymnasium.spaces`` (reference
gym_kmanip/env_base.py:116-190).  gymnasium is absent from the build image, so the drop-in classes use
the real package when it imports and this shim otherwise; only what ``check_env`` exercises is provided
(shape / dtype / bounds, ``sample``, ``contains``, ``seed``).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import numpy as np

try:                                    # pragma: no cover - gymnasium is not in the build image
    from gymnasium import Env as GymEnv
    from gymnasium.spaces import Box, Dict
    HAVE_GYMNASIUM = True
except Exception:                       # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32, seed: Optional[int] = None):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return seed

        def sample(self):
            if np.issubdtype(self.dtype, np.integer):
                return self._rng.integers(self.low, self.high, size=self.shape, endpoint=True).astype(self.dtype)
            return self._rng.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return (x.shape == self.shape and np.can_cast(x.dtype, self.dtype)
                    and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high)))

        __contains__ = contains

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Dict:
        def __init__(self, spaces=None, seed: Optional[int] = None):
            self.spaces = OrderedDict(spaces or {})
            if seed is not None:
                self.seed(seed)

        def seed(self, seed=None):
            for i, s in enumerate(self.spaces.values()):
                s.seed(None if seed is None else seed + i)
            return seed

        def sample(self):
            return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

        def contains(self, x) -> bool:
            return (isinstance(x, dict) and list(x.keys()) == list(self.spaces.keys())
                    and all(s.contains(x[k]) for k, s in self.spaces.items()))

        __contains__ = contains

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def __len__(self):
            return len(self.spaces)

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k}: {s}" for k, s in self.spaces.items()) + ")"

    class GymEnv:
        metadata: dict = {}
        render_mode = None
        spec = None

        @property
        def unwrapped(self):
            return self

        def reset(self, seed=None, options=None):
            if seed is not None:
                self.np_random = np.random.default_rng(seed)

        def close(self):
            pass
