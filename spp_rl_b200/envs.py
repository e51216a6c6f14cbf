"""Environments for running the drop-in classes offline (gym and MuJoCo are not available in this image).

`make(name)` follows the slice of the gym-0.15 API the reference touches (rltoolkit/rl.py:45-53,185;
rltoolkit/algorithms/ddpg/ddpg.py:194,209): observation_space.shape/high, action_space.shape/high/sample(),
_max_episode_steps, reset() -> obs, step(a) -> (obs, reward, done, info).  Pendulum-v0 is the classic-control
dynamics; the MuJoCo names are SHAPE-ONLY synthetic stand-ins (unbounded observations, |a| <= 1, 1000 steps).
Pass a real gym environment through the `env=` constructor argument of the algorithm classes to use one.
"""
import math
import os
import warnings

import numpy as np


class _Box:
    def __init__(self, low, high):
        self.low = np.asarray(low, np.float32)
        self.high = np.asarray(high, np.float32)
        self.shape = self.low.shape
        self._rng = np.random.RandomState(0)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(np.float32)


class Pendulum:
    _max_episode_steps = 200

    def __init__(self, seed=0):
        self.action_space = _Box([-2.0], [2.0])
        self.observation_space = _Box([-1.0, -1.0, -8.0], [1.0, 1.0, 8.0])
        self.rng = np.random.RandomState(seed)
        self.t = 0

    def _obs(self):
        th, thd = self.state
        return np.array([math.cos(th), math.sin(th), thd])

    def reset(self):
        self.state = self.rng.uniform([-math.pi, -1.0], [math.pi, 1.0])
        self.t = 0
        return self._obs()

    def step(self, u):
        th, thd = self.state
        u = float(np.clip(np.asarray(u, np.float64).reshape(-1)[0], -2.0, 2.0))
        ang = ((th + math.pi) % (2 * math.pi)) - math.pi
        cost = ang ** 2 + 0.1 * thd ** 2 + 0.001 * u ** 2
        thd = thd + (-3 * 10.0 / 2 * math.sin(th + math.pi) + 3.0 * u) * 0.05
        th = th + thd * 0.05
        thd = float(np.clip(thd, -8.0, 8.0))
        self.state = np.array([th, thd])
        self.t += 1
        return self._obs(), -cost, self.t >= self._max_episode_steps, {}

    def close(self):
        pass


class SyntheticControl:
    _max_episode_steps = 1000

    def __init__(self, ob_dim, ac_dim, seed=0):
        self.ob_dim, self.ac_dim = ob_dim, ac_dim
        self.action_space = _Box(-np.ones(ac_dim), np.ones(ac_dim))
        self.observation_space = _Box(np.full(ob_dim, -np.inf), np.full(ob_dim, np.inf))
        self.rng = np.random.RandomState(seed)
        self.mix = np.random.RandomState(4321 + ob_dim).randn(ob_dim, ac_dim) / math.sqrt(ac_dim)
        self.t = 0

    def reset(self):
        self.state = 0.1 * self.rng.randn(self.ob_dim)
        self.t = 0
        return self.state.copy()

    def step(self, a):
        a = np.clip(np.asarray(a, np.float64).reshape(-1), -1.0, 1.0)
        self.state = 0.98 * self.state + 0.1 * np.tanh(self.mix @ a) + 0.02 * self.rng.randn(self.ob_dim)
        self.t += 1
        done = bool(self.t >= self._max_episode_steps or self.rng.rand() < 0.004)
        return self.state.copy(), float(self.state[0] - 0.01 * np.square(a).sum()), done, {}

    def close(self):
        pass


_SHAPES = {"Hopper-v2": (11, 3), "HalfCheetah-v2": (17, 6), "Walker2d-v2": (17, 6), "Ant-v2": (111, 8)}


def _real_gym():
    """The installed gym / gymnasium package, or None (the oracle's import shim under oracle/shims is not a simulator)."""
    for mod in ("gym", "gymnasium"):
        try:
            g = __import__(mod)
        except Exception:
            continue
        if getattr(g, "__spp_shim__", False) or "oracle/shims" in (getattr(g, "__file__", "") or "").replace(os.sep, "/"):
            continue
        return g
    return None


def make(name, seed=0, synthetic=None):
    """gym.make(name) when a real gym is installed (rltoolkit/rl.py:45).  Without one: Pendulum-v0 is the classic-control dynamics
    restated above; the MuJoCo names fall back to the SHAPE-ONLY SyntheticControl stand-in -- silently only on explicit opt-in
    (synthetic=True or SPP_RL_SYNTHETIC_ENVS=1), otherwise with a RuntimeWarning, because returns measured on it mean nothing."""
    if synthetic is None:
        synthetic = os.environ.get("SPP_RL_SYNTHETIC_ENVS", "") == "1"
    g = None if synthetic else _real_gym()
    if g is not None:
        try:
            return g.make(name)
        except Exception as e:      # e.g. mujoco-py missing: say so instead of silently training on fake dynamics
            warnings.warn("gym.make(%r) failed (%s); falling back to the offline stand-in" % (name, e), RuntimeWarning, stacklevel=2)
    if name == "Pendulum-v0":
        return Pendulum(seed)
    if name in _SHAPES:
        if not synthetic:
            warnings.warn("no gym/MuJoCo in this environment: %r is a SHAPE-ONLY synthetic stand-in (SyntheticControl); returns are "
                          "not comparable with the reference's.  Pass env=<gym env>, or opt in with synthetic=True / "
                          "SPP_RL_SYNTHETIC_ENVS=1 to silence this warning." % (name,), RuntimeWarning, stacklevel=2)
        return SyntheticControl(*_SHAPES[name], seed=seed)
    raise KeyError("unknown environment %r (pass env=<gym env> to use a real one)" % (name,))
