"""Helpers the hot path's callers rely on (reference ``utils.py``): ``merge_stat``
(:15-29), ``LogField`` (:13) and ``init_args_for_env`` (:107-132)."""
import numbers
import sys
from collections import namedtuple

import numpy as np

LogField = namedtuple('LogField', ('data', 'plot', 'x_axis', 'divide_by'))


def merge_stat(src, dest):
    """Fold ``src`` into ``dest`` in place with the reference's rules (utils.py:15-29): a key new to ``dest`` is
    taken over as is; numbers and numpy arrays accumulate by ``+``; anything else is gathered into a list (two lists
    are concatenated, a value is appended to an existing list, two plain values become a two-element list)."""
    for key, val in src.items():
        if key not in dest:
            dest[key] = val
            continue
        if isinstance(val, (numbers.Number, np.ndarray)):
            dest[key] = dest[key] + val
            continue
        cur = dest[key]
        if not isinstance(cur, list):
            dest[key] = [cur, val]
        elif isinstance(val, list):
            cur.extend(val)
        else:
            cur.append(val)


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
