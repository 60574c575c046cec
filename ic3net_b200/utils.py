"""Helpers the hot path's callers rely on (reference ``utils.py``): ``merge_stat``
(:15-29), ``LogField`` (:13) and ``init_args_for_env`` (:107-132)."""
import numbers
import sys
from collections import namedtuple

import numpy as np

LogField = namedtuple('LogField', ('data', 'plot', 'x_axis', 'divide_by'))


def merge_stat(src, dest):
    # utils.py:15-29: numbers and arrays add, everything else is collected in lists
    for k, v in src.items():
        if k not in dest:
            dest[k] = v
        elif isinstance(v, numbers.Number):
            dest[k] = dest.get(k, 0) + v
        elif isinstance(v, np.ndarray):
            dest[k] = dest.get(k, 0) + v
        else:
            if isinstance(dest[k], list) and isinstance(v, list):
                dest[k].extend(v)
            elif isinstance(dest[k], list):
                dest[k].append(v)
            else:
                dest[k] = [dest[k], v]


def init_args_for_env(parser, argv=None):
    """Let the chosen env contribute its flags (utils.py:107-132)."""
    from . import data
    env_dict = {'predator_prey': 'PredatorPrey-v0', 'traffic_junction': 'TrafficJunction-v0'}
    args = sys.argv if argv is None else argv
    env_name = None
    for index, item in enumerate(args):
        if item == '--env_name':
            env_name = args[index + 1]
    if not env_name or env_name not in env_dict:
        return
    env = data.make(env_dict[env_name])
    env.init_args(parser)
