"""gym.spaces stand-in (oracle/test infrastructure only; see gym/__init__.py)."""
import numpy as np


class Space:
    pass


class Discrete(Space):
    def __init__(self, n):
        self.n = n
        self.shape = ()
        self._rng = np.random.RandomState(0)

    def sample(self):
        return int(self._rng.randint(self.n))


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is not None:
            low = np.full(shape, low, dtype=np.float64)
            high = np.full(shape, high, dtype=np.float64)
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = dtype
        self._rng = np.random.RandomState(0)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)
