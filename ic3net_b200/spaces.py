"""Space descriptors (the subset of ``gym.spaces`` the reference touches).

The reference uses gym only as a registry and for these descriptors
(predator_prey_env.py:95,107; traffic_junction_env.py:109,135-148;
env_wrappers.py:21-50); no arithmetic lives there, so the batched envs carry
their own minimal copies and need no gym install.
"""
import numpy as np


class Box(object):
    def __init__(self, low=0, high=1, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class Discrete(object):
    def __init__(self, n):
        self.n = n
        self.shape = ()


class MultiDiscrete(object):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec)
        self.shape = self.nvec.shape


class MultiBinary(object):
    def __init__(self, n):
        self.n = n
        self.shape = tuple(n) if isinstance(n, (tuple, list)) else (n,)


class Tuple(object):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)
