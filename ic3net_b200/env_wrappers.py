"""Batched counterpart of the reference ``env_wrappers.py:GymWrapper`` (:7-107).

Same properties and methods; observations come back as float32 CUDA tensors
``[B, N, obs_dim]`` instead of a float64 CPU tensor ``[1, N, obs_dim]``.
"""
from inspect import getfullargspec

import numpy as np
import torch


class GymWrapper(object):
    '''
    for multi-agent
    '''
    def __init__(self, env):
        self.env = env

    @property
    def observation_dim(self):
        # env_wrappers.py:14-31
        if hasattr(self.env.observation_space, 'spaces'):
            total_obs_dim = 0
            for space in self.env.observation_space.spaces:
                if hasattr(self.env.action_space, 'shape'):
                    total_obs_dim += int(np.prod(space.shape))
                else:  # Discrete
                    total_obs_dim += 1
            return total_obs_dim
        else:
            return int(np.prod(self.env.observation_space.shape))

    @property
    def num_actions(self):
        if hasattr(self.env.action_space, 'nvec'):
            return int(self.env.action_space.nvec[0])
        elif hasattr(self.env.action_space, 'n'):
            return self.env.action_space.n

    @property
    def dim_actions(self):
        if hasattr(self.env.action_space, 'nvec'):
            return self.env.action_space.shape[0]
        elif hasattr(self.env.action_space, 'n'):
            return 1

    @property
    def action_space(self):
        return self.env.action_space

    @property
    def nenvs(self):
        return self.env.nenvs

    def reset(self, epoch):
        reset_args = getfullargspec(self.env.reset).args
        if 'epoch' in reset_args:
            obs = self.env.reset(epoch)
        else:
            obs = self.env.reset()
        return self._flatten_obs(obs)

    def display(self):
        self.env.render()

    def end_display(self):
        self.env.exit_render()

    def step(self, action):
        # env_wrappers.py:73-80: only the first action head reaches the env
        if self.dim_actions == 1:
            action = action[0]
        obs, r, done, info = self.env.step(action)
        obs = self._flatten_obs(obs)
        return (obs, r, done, info)

    def reward_terminal(self):
        if hasattr(self.env, 'reward_terminal'):
            return self.env.reward_terminal()
        else:
            return np.zeros(1)

    def _flatten_obs(self, obs):
        # the kernels already write rows in the flattened order of env_wrappers.py:88-98
        return obs.reshape(self.env.nenvs, -1, self.observation_dim)

    def get_stat(self):
        if hasattr(self.env, 'get_stat'):
            stat = self.env.get_stat()
            stat.pop('steps_taken', None)
            return stat
        else:
            return dict()
