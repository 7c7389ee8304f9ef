"""CartPole-v1 restatement -- TEST INFRASTRUCTURE (oracle side of the evaluation rollouts).

gymnasium is not installable here; this restates the published CartPole-v1 dynamics (gymnasium
classic_control/cartpole.py: Euler integration, tau = 0.02, force 10 N, 12 degree / 2.4 m termination, 500-step
time limit, reset state uniform(-0.05, 0.05) from a PCG64 generator seeded with the reset seed).  It backs the
`gymnasium` shim under tests/golden/shims/ that lets the reference's own train.py run unmodified (train.py:18-19,
173-187).  PARITY UNPINNED for the environment: no reference fixture pins gymnasium's numbers."""
import math

import numpy as np


class CartPoleV1:
    gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    theta_threshold = 12 * 2 * math.pi / 360
    x_threshold = 2.4
    max_steps = 500

    def __init__(self):
        self.np_random = np.random.Generator(np.random.PCG64())
        self.state = None
        self.steps = 0

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        self.state = self.np_random.uniform(low=-0.05, high=0.05, size=(4,))
        self.steps = 0
        return np.array(self.state, dtype=np.float32), {}

    def step(self, action):
        x, x_dot, theta, theta_dot = self.state
        force = self.force_mag if action == 1 else -self.force_mag
        costheta, sintheta = math.cos(theta), math.sin(theta)
        total_mass = self.masspole + self.masscart
        polemass_length = self.masspole * self.length
        temp = (force + polemass_length * theta_dot ** 2 * sintheta) / total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * costheta ** 2 / total_mass))
        xacc = temp - polemass_length * thetaacc * costheta / total_mass
        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc
        self.state = np.array([x, x_dot, theta, theta_dot], dtype=np.float64)
        self.steps += 1
        terminated = bool(x < -self.x_threshold or x > self.x_threshold
                          or theta < -self.theta_threshold or theta > self.theta_threshold)
        truncated = self.steps >= self.max_steps
        return np.array(self.state, dtype=np.float32), 1.0, terminated, truncated, {}
