"""Env factory with the reference's signature (``data.py:6-36``): ``init(env_name, args)``.

Only the two environments of the accelerated path exist; the reference's other
names raise exactly like an unknown name does there (data.py:33-34).
"""
from .env_wrappers import GymWrapper
from .predator_prey_env import PredatorPreyEnv
from .traffic_junction_env import TrafficJunctionEnv

# same ids the reference registers (ic3net_envs/__init__.py:3-11)
REGISTRY = {'PredatorPrey-v0': PredatorPreyEnv, 'TrafficJunction-v0': TrafficJunctionEnv}


def make(env_id):
    return REGISTRY[env_id]()


def init(env_name, args, final_init=True):
    if env_name == 'predator_prey':
        env = make('PredatorPrey-v0')
        env.multi_agent_init(args)
        env = GymWrapper(env)
    elif env_name == 'traffic_junction':
        env = make('TrafficJunction-v0')
        env.multi_agent_init(args)
        env = GymWrapper(env)
    else:
        raise RuntimeError("wrong env name")
    return env
