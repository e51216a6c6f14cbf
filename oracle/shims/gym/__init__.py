"""Minimal stand-in for gym 0.15 -- ORACLE/TEST INFRASTRUCTURE ONLY.

The reference (rltoolkit) imports `gym` at module scope (rltoolkit/rl.py:6,45-53,185;
rltoolkit/tensorboard_logger.py:8; rltoolkit/acm/acm.py:102) and gym is not installed in
this image.  This stub gives exactly the surface those call sites touch so the UNMODIFIED
reference can be imported in the build container to generate golden fixtures
(tests/golden/make_golden.py).  It is never imported by the product package.

Environments:
  * Pendulum-v0   -- faithful restatement of the classic-control dynamics (obs 3, act 1,
                     |a| <= 2, obs high [1, 1, 8], 200-step limit).
  * CartPole-v0   -- faithful classic-control dynamics (obs 4, Discrete(2), 200 steps).
  * Hopper-v2, HalfCheetah-v2, Walker2d-v2, Ant-v2 -- SHAPE-ONLY synthetic stand-ins
                     (MuJoCo is unavailable offline): unbounded observation space,
                     |a| <= 1, 1000-step limit, smooth deterministic pseudo-dynamics.
"""
import math

import numpy as np

from . import spaces  # noqa: F401
from .spaces import Box, Discrete

__version__ = "0.15.4-stub"


class Env:
    _max_episode_steps = 1000
    metadata = {"render.modes": []}

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def close(self):
        pass

    def render(self, mode="human"):
        return None


class PendulumEnv(Env):
    _max_episode_steps = 200

    def __init__(self):
        self.max_speed = 8.0
        self.max_torque = 2.0
        self.dt = 0.05
        self.g = 10.0
        self.m = 1.0
        self.l = 1.0
        high = np.array([1.0, 1.0, self.max_speed], dtype=np.float32)
        self.action_space = Box(-self.max_torque, self.max_torque, shape=(1,))
        self.observation_space = Box(-high, high)
        self.np_random = np.random.RandomState(0)
        self._t = 0

    def _obs(self):
        th, thdot = self.state
        return np.array([math.cos(th), math.sin(th), thdot])

    def reset(self):
        self.state = self.np_random.uniform(low=[-math.pi, -1.0], high=[math.pi, 1.0])
        self._t = 0
        return self._obs()

    def step(self, u):
        th, thdot = self.state
        u = float(np.clip(np.asarray(u, dtype=np.float64).reshape(-1)[0],
                          -self.max_torque, self.max_torque))
        ang = ((th + math.pi) % (2 * math.pi)) - math.pi
        cost = ang ** 2 + 0.1 * thdot ** 2 + 0.001 * (u ** 2)
        newthdot = thdot + (-3 * self.g / (2 * self.l) * math.sin(th + math.pi)
                            + 3.0 / (self.m * self.l ** 2) * u) * self.dt
        newth = th + newthdot * self.dt
        newthdot = float(np.clip(newthdot, -self.max_speed, self.max_speed))
        self.state = np.array([newth, newthdot])
        self._t += 1
        done = self._t >= self._max_episode_steps  # TimeLimit wrapper behaviour
        return self._obs(), -cost, done, {}


class CartPoleEnv(Env):
    _max_episode_steps = 200

    def __init__(self):
        self.gravity, self.masscart, self.masspole = 9.8, 1.0, 0.1
        self.total_mass = self.masspole + self.masscart
        self.length = 0.5
        self.polemass_length = self.masspole * self.length
        self.force_mag, self.tau = 10.0, 0.02
        self.theta_threshold_radians = 12 * 2 * math.pi / 360
        self.x_threshold = 2.4
        high = np.array([self.x_threshold * 2, np.finfo(np.float32).max,
                         self.theta_threshold_radians * 2, np.finfo(np.float32).max])
        self.action_space = Discrete(2)
        self.observation_space = Box(-high, high)
        self.np_random = np.random.RandomState(0)
        self._t = 0

    def reset(self):
        self.state = self.np_random.uniform(low=-0.05, high=0.05, size=(4,))
        self._t = 0
        return np.array(self.state)

    def step(self, action):
        x, x_dot, theta, theta_dot = self.state
        force = self.force_mag if int(action) == 1 else -self.force_mag
        costheta, sintheta = math.cos(theta), math.sin(theta)
        temp = (force + self.polemass_length * theta_dot ** 2 * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * costheta ** 2 / self.total_mass))
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass
        x, x_dot = x + self.tau * x_dot, x_dot + self.tau * xacc
        theta, theta_dot = theta + self.tau * theta_dot, theta_dot + self.tau * thetaacc
        self.state = (x, x_dot, theta, theta_dot)
        self._t += 1
        done = bool(x < -self.x_threshold or x > self.x_threshold
                    or theta < -self.theta_threshold_radians
                    or theta > self.theta_threshold_radians
                    or self._t >= self._max_episode_steps)
        return np.array(self.state), 1.0, done, {}


class SyntheticMujocoEnv(Env):
    """Shape-only stand-in: obs' = 0.98*obs + 0.1*tanh(M a) + 0.02*noise; random early ends."""

    _max_episode_steps = 1000

    def __init__(self, ob_dim, ac_dim, seed=0):
        self.ob_dim, self.ac_dim = ob_dim, ac_dim
        self.action_space = Box(-1.0, 1.0, shape=(ac_dim,))
        high = np.full(ob_dim, np.inf)
        self.observation_space = Box(-high, high)
        self.np_random = np.random.RandomState(seed)
        self._mix = np.random.RandomState(1234 + ob_dim).randn(ob_dim, ac_dim) / math.sqrt(ac_dim)
        self._t = 0

    def reset(self):
        self.state = 0.1 * self.np_random.randn(self.ob_dim)
        self._t = 0
        return self.state.copy()

    def step(self, a):
        a = np.clip(np.asarray(a, dtype=np.float64).reshape(-1), -1.0, 1.0)
        self.state = (0.98 * self.state + 0.1 * np.tanh(self._mix @ a)
                      + 0.02 * self.np_random.randn(self.ob_dim))
        self._t += 1
        rew = float(self.state[0] - 0.01 * np.square(a).sum())
        done = bool(self._t >= self._max_episode_steps or self.np_random.rand() < 0.004)
        return self.state.copy(), rew, done, {}


_REGISTRY = {
    "Pendulum-v0": lambda: PendulumEnv(),
    "CartPole-v0": lambda: CartPoleEnv(),
    "Hopper-v2": lambda: SyntheticMujocoEnv(11, 3),
    "HalfCheetah-v2": lambda: SyntheticMujocoEnv(17, 6),
    "Walker2d-v2": lambda: SyntheticMujocoEnv(17, 6),
    "Ant-v2": lambda: SyntheticMujocoEnv(111, 8),
}


def make(name):
    if name not in _REGISTRY:
        raise KeyError("gym stub: unknown environment %r" % (name,))
    return _REGISTRY[name]()
