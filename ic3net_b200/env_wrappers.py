"""Batched counterpart of the reference's ``GymWrapper`` (env_wrappers.py:7-107).

The reference wraps ONE gym environment and hands back a float64 CPU tensor ``[1, N, obs_dim]``; here the wrapped
object is a batch of ``nenvs`` environment instances living in HBM and observations are float32 CUDA tensors
``[nenvs, N, obs_dim]``.  The surface is the reference's: ``observation_dim``, ``num_actions``, ``dim_actions``,
``action_space``, ``reset(epoch)``, ``step(action)``, ``reward_terminal()``, ``get_stat()``, ``display()``,
``end_display()`` (+ ``nenvs``).
"""
from inspect import getfullargspec

import numpy as np


def _flat_size(space):
    return int(np.prod(space.shape))


def per_agent_obs_dim(observation_space, action_space):
    """Length of one agent's flattened observation (env_wrappers.py:14-31).  A composite (Tuple) observation is
    the concatenation of its members, each flattened -- a member counts 1 when the ACTION space carries no
    ``shape`` attribute, the reference's own test (it never triggers for the spaces of this repo)."""
    members = getattr(observation_space, 'spaces', None)
    if members is None:
        return _flat_size(observation_space)
    by_shape = hasattr(action_space, 'shape')
    return sum(_flat_size(m) if by_shape else 1 for m in members)


def action_layout(action_space):
    """(actions of the first head, heads the environment consumes) (env_wrappers.py:33-50): MultiDiscrete spaces
    describe one head per entry of ``nvec``, Discrete spaces a single head of ``n`` actions."""
    if hasattr(action_space, 'nvec'):
        return int(action_space.nvec[0]), action_space.shape[0]
    if hasattr(action_space, 'n'):
        return action_space.n, 1
    return None, None


class GymWrapper(object):
    """Multi-agent wrapper around a batched environment (``PredatorPreyEnv`` / ``TrafficJunctionEnv``)."""

    def __init__(self, env):
        self.env = env

    # ---- static facts of the wrapped spaces ---------------------------------------------------------
    @property
    def observation_dim(self):
        return per_agent_obs_dim(self.env.observation_space, self.env.action_space)

    @property
    def num_actions(self):
        return action_layout(self.env.action_space)[0]

    @property
    def dim_actions(self):
        return action_layout(self.env.action_space)[1]

    @property
    def action_space(self):
        return self.env.action_space

    @property
    def nenvs(self):
        return self.env.nenvs

    # ---- episode control ----------------------------------------------------------------------------
    def reset(self, epoch):
        # the traffic junction's curriculum wants the epoch, predator-prey takes no argument (trainer.py:28-32)
        wants_epoch = 'epoch' in getfullargspec(self.env.reset).args
        first = self.env.reset(epoch) if wants_epoch else self.env.reset()
        return self._as_agent_rows(first)

    def step(self, action):
        """``action``: per-head list.  An environment with one action dimension only sees head 0 -- the gating
        head of IC3Net never reaches it (env_wrappers.py:73-80)."""
        env_action = action[0] if self.dim_actions == 1 else action
        obs, reward, done, info = self.env.step(env_action)
        return self._as_agent_rows(obs), reward, done, info

    def reward_terminal(self):
        fn = getattr(self.env, 'reward_terminal', None)
        return fn() if fn is not None else np.zeros(1)

    def get_stat(self):
        fn = getattr(self.env, 'get_stat', None)
        if fn is None:
            return dict()
        stat = fn()
        stat.pop('steps_taken', None)            # env_wrappers.py:102-104: the trainer counts steps itself
        return stat

    # ---- rendering (curses in the reference; out of scope here: the batched envs raise NotImplementedError) ----
    def display(self):
        self.env.render()

    def end_display(self):
        self.env.exit_render()

    # ---- helpers -------------------------------------------------------------------------------------
    def _as_agent_rows(self, obs):
        # the kernels already emit every agent's observation in the flattened member order of
        # env_wrappers.py:88-98, so flattening is a view: [nenvs, N, obs_dim]
        return obs.reshape(self.env.nenvs, -1, self.observation_dim)

    _flatten_obs = _as_agent_rows                # the reference's name for it (env_wrappers.py:83)
