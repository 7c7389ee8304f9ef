"""Host-side CartPole-v1 for the greedy evaluation rollouts of the training driver
(reference: gymnasium's env used at /root/reference/src/train.py:18-19,173-187; gymnasium is not available).
Published CartPole-v1 dynamics: Euler integration, tau 0.02 s, +-10 N, termination at 12 degrees / 2.4 m,
500-step time limit, reset state uniform(-0.05, 0.05) from a PCG64 generator seeded with the reset seed."""
from __future__ import annotations

import math

import numpy as np


class CartPole:
    GRAVITY, M_CART, M_POLE, HALF_LEN, FORCE, TAU = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    THETA_MAX = 12 * 2 * math.pi / 360
    X_MAX = 2.4
    STEP_LIMIT = 500

    def __init__(self):
        self._rng = np.random.Generator(np.random.PCG64())
        self._s = np.zeros(4)
        self._t = 0

    def reset(self, seed=None):
        if seed is not None:
            self._rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        self._s = self._rng.uniform(low=-0.05, high=0.05, size=(4,))
        self._t = 0
        return self._s.astype(np.float32), {}

    def step(self, action):
        x, xd, th, thd = self._s
        f = self.FORCE if action == 1 else -self.FORCE
        c, s = math.cos(th), math.sin(th)
        m = self.M_POLE + self.M_CART
        ml = self.M_POLE * self.HALF_LEN
        tmp = (f + ml * thd ** 2 * s) / m
        tha = (self.GRAVITY * s - c * tmp) / (self.HALF_LEN * (4.0 / 3.0 - self.M_POLE * c ** 2 / m))
        xa = tmp - ml * tha * c / m
        self._s = np.array([x + self.TAU * xd, xd + self.TAU * xa, th + self.TAU * thd, thd + self.TAU * tha])
        self._t += 1
        done = bool(abs(self._s[0]) > self.X_MAX or abs(self._s[2]) > self.THETA_MAX)
        return self._s.astype(np.float32), 1.0, done, self._t >= self.STEP_LIMIT, {}
