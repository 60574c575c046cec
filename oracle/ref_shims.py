"""Import the UNMODIFIED reference from /root/reference (this container only).

TEST INFRASTRUCTURE.  Used by ``oracle/gen_golden.py`` (fixture generation) and
by ``oracle/ref_baseline.py`` (the CPU baseline of record: the reference's own
``MultiProcessTrainer`` timed on the host cores; on the GPU box it runs from the
staged copy ``oracle/_ref``, see ``oracle/build_ref.py``).  Nothing here changes reference arithmetic; the
shims only make 2018-era code importable on python 3.12 / numpy 2.3 / torch 2.11
(SURVEY.md section 8(c)):

1. a ``gym`` stub package (registry + space descriptors only; no arithmetic
   lives in gym: predator_prey_env.py:25-27,95,107; traffic_junction_env.py:24-26,
   109,135-148; env_wrappers.py:21-50)
2. ``inspect.getargspec`` alias (trainer.py:2,28; env_wrappers.py:5,57)
3. ``_all_idx`` wrapper: numpy>=2 returns a tuple from ``np.ogrid``
   (predator_prey_env.py:302-305; traffic_junction_env.py:606-609)
4. ``CommNetMLP.get_agent_mask`` returns a clone (comm.py:175 does an in-place
   multiply on an expanded view, rejected by modern torch)
5. ``Optimizer.zero_grad(set_to_none=False)`` (multi_processing.py:60-72 caches
   grad tensors)

A *random tape* can be installed on the reference side so that its
``np.random.uniform`` / ``np.random.choice`` / ``torch.multinomial`` draws
consume externally supplied uniforms.
"""
import inspect
import os
import sys
import types

import numpy as np

def _find_ref_root():
    """/root/reference in the development container; the staged copy ``oracle/_ref`` (oracle/build_ref.py,
    git-ignored, shipped by gpurun) on the GPU box."""
    cands = [os.environ.get("IC3NET_REFERENCE"), "/root/reference",
             os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "ic3net-envs", "ic3net_envs")):
            return c
    return "/root/reference"


REF_ROOT = _find_ref_root()


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "ic3net-envs", "ic3net_envs"))


# --------------------------------------------------------------------------
# gym stub
# --------------------------------------------------------------------------
def _make_gym_stub():
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")

    class Env(object):
        pass

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
        def __init__(self, spaces_):
            self.spaces = tuple(spaces_)

    registry = {}

    def register(id, entry_point, **kw):
        registry[id] = entry_point

    def make(id):
        mod, cls = registry[id].split(":")
        m = __import__(mod, fromlist=[cls])
        return getattr(m, cls)()

    gym.Env = Env
    gym.make = make
    gym.spaces = spaces
    gym.envs = envs
    envs.registration = registration
    registration.register = register
    for c in (Box, Discrete, MultiDiscrete, MultiBinary, Tuple):
        setattr(spaces, c.__name__, c)
    return {"gym": gym, "gym.spaces": spaces, "gym.envs": envs,
            "gym.envs.registration": registration}


_installed = False


def install():
    """Make ``import ic3net_envs, comm, trainer, ...`` work.  Idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    import torch

    for name, mod in _make_gym_stub().items():
        sys.modules.setdefault(name, mod)
    if not hasattr(inspect, "getargspec"):
        inspect.getargspec = inspect.getfullargspec
    for p in (REF_ROOT, os.path.join(REF_ROOT, "ic3net-envs")):
        if p not in sys.path:
            sys.path.insert(0, p)

    import ic3net_envs  # noqa: F401  (registers the env ids)
    from ic3net_envs import predator_prey_env, traffic_junction_env

    def _all_idx(self, idx, axis):
        grid = list(np.ogrid[tuple(map(slice, idx.shape))])
        grid.insert(axis, idx)
        return tuple(grid)

    predator_prey_env.PredatorPreyEnv._all_idx = _all_idx
    traffic_junction_env.TrafficJunctionEnv._all_idx = _all_idx

    import comm

    _orig_mask = comm.CommNetMLP.get_agent_mask

    def get_agent_mask(self, batch_size, info):
        n_alive, mask = _orig_mask(self, batch_size, info)
        return n_alive, mask.clone().to(torch.get_default_dtype())

    comm.CommNetMLP.get_agent_mask = get_agent_mask

    _orig_zero = torch.optim.Optimizer.zero_grad

    def zero_grad(self, set_to_none=False):
        return _orig_zero(self, set_to_none=False)

    torch.optim.Optimizer.zero_grad = zero_grad
    _installed = True


# --------------------------------------------------------------------------
# random tape (reference side)
# --------------------------------------------------------------------------
class Tape(object):
    """Sequential uniforms in [0,1), each an exact multiple of 2**-24."""

    def __init__(self, u):
        self.u = np.asarray(u, dtype=np.float64).ravel()
        self.i = 0

    def next(self):
        v = self.u[self.i]
        self.i += 1
        return float(v)


class tape_patch(object):
    """Context manager: route the reference's numpy draws through ``tape``.

    Mappings:
      np.random.uniform()            -> u
      np.random.choice(k or array)   -> element floor(u*k)
    """

    def __init__(self, tape):
        self.tape = tape

    def __enter__(self):
        self._uniform, self._choice = np.random.uniform, np.random.choice
        tape = self.tape

        def uniform(*a, **k):
            assert not a and not k
            return tape.next()

        def choice(a, size=None, replace=True, p=None):
            assert size is None and p is None
            arr = np.arange(a) if np.isscalar(a) else np.asarray(a)
            return arr[int(np.floor(tape.next() * len(arr)))]

        np.random.uniform, np.random.choice = uniform, choice
        return self

    def __exit__(self, *exc):
        np.random.uniform, np.random.choice = self._uniform, self._choice
        return False


def make_args(**kw):
    """Namespace with the main.py defaults (main.py:25-109) + derived fields
    (main.py:115-155) for a CommNet/IC3Net run."""
    import argparse

    d = dict(num_epochs=100, epoch_size=10, batch_size=500, nprocesses=1, hid_size=128,
             recurrent=True, gamma=1.0, tau=1.0, seed=0, normalize_rewards=False, lrate=0.001,
             entr=0.0, value_coeff=0.01, env_name="predator_prey", max_steps=20, nactions="1",
             action_scale=1.0, plot=False, plot_env="main", save="", save_every=0, load="",
             display=False, random=False, commnet=True, ic3net=False, nagents=3, comm_mode="avg",
             comm_passes=1, comm_mask_zero=False, mean_ratio=1.0, rnn_type="MLP",
             detach_gap=10000, comm_init="uniform", hard_attn=False, comm_action_one=False,
             advantages_per_action=False, share_weights=False,
             # predator-prey flags (predator_prey_env.py:55-70)
             nenemies=1, dim=5, vision=0, moving_prey=False, no_stay=False, mode="mixed",
             enemy_comm=False,
             # traffic-junction flags (traffic_junction_env.py:60-77)
             add_rate_min=0.05, add_rate_max=0.2, curr_start=0, curr_end=0, difficulty="easy",
             vocab_type="bool")
    d.update(kw)
    a = argparse.Namespace(**d)
    if a.ic3net:
        a.commnet, a.hard_attn, a.mean_ratio = 1, 1, 0
        if a.env_name == "traffic_junction":
            a.comm_action_one = True
    a.nfriendly = a.nagents
    if getattr(a, "enemy_comm", False):          # main.py:124-131: the prey becomes one more agent of the policy
        a.nagents += a.nenemies
    return a


def make_ref_env(args):
    """data.init without importing data.py's optional envs (data.py:16-28)."""
    install()
    import gym
    from env_wrappers import GymWrapper

    env = gym.make({"predator_prey": "PredatorPrey-v0",
                    "traffic_junction": "TrafficJunction-v0"}[args.env_name])
    env.multi_agent_init(args)
    return GymWrapper(env)


def finish_args(args, env):
    """main.py:134-155."""
    install()
    from action_utils import parse_action_args

    args.num_inputs = env.observation_dim
    na = env.num_actions
    args.num_actions = [na] if not isinstance(na, (list, tuple)) else list(na)
    args.dim_actions = env.dim_actions
    if args.hard_attn and args.commnet:
        args.num_actions = [*args.num_actions, 2]
        args.dim_actions = env.dim_actions + 1
    if args.commnet and (args.recurrent or args.rnn_type == "LSTM"):
        args.recurrent = True
        args.rnn_type = "LSTM"
    parse_action_args(args)
    return args
