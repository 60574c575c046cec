"""Observation handle of a batched environment.

The reference's ``env.step`` returns the observation as a dense ``[N, obs_dim]`` array that ``CommNetMLP.forward``
immediately multiplies by the encoder weight (comm.py:119).  For the one-hot observations of this path that tensor is
> 99 % zeros (1.19 GB per step at 8192 predator-prey envs): materialising it and reading it back is most of a step's
HBM traffic.  With ``args.obs_api = 'handle'`` the environments therefore return a ``LazyObs`` instead: a handle on
the environment state the observation is a function of.  ``CommNetMLP.forward`` accepts it wherever it accepts the
tensor and evaluates the encoder straight from the state (same additions in the same order as the dense encoder:
bit-identical ``x``); anything else that treats it as a tensor (``.dense()``, indexing, ``.cpu()``, ``.shape`` ...)
gets the exact dense observation, gathered on demand.

A handle describes the state at the time it was returned: it is valid until the next ``step`` / ``reset`` of its
environment (``env.obs_version``); using a stale handle raises.
"""
import torch


class LazyObs(object):
    def __init__(self, env, shape=None):
        self.env = env
        self.version = env.obs_version
        self.shape = tuple(env.obs_shape) if shape is None else tuple(shape)
        self.device = env.device
        self.dtype = torch.float32

    # ---- what CommNetMLP.forward needs ---------------------------------------------------------------
    def check_current(self):
        if self.version != self.env.obs_version:
            raise RuntimeError("stale observation handle: the environment has stepped / reset since it was returned "
                               "(call .dense() right after env.step to keep a copy)")
        return self.env

    # ---- tensor-like surface: materialise on demand ----------------------------------------------------
    def dense(self):
        return self.check_current()._get_obs().reshape(self.shape)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return len(self.shape)

    def numel(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    def reshape(self, *shape):
        """A reshaped handle (still lazy): GymWrapper._flatten_obs turns [nenvs, N, W, W, V] into [nenvs, N, obs_dim]."""
        shape = tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else tuple(shape)
        prod = 1
        for s in shape:
            if s != -1:
                prod *= s
        full = tuple(s if s != -1 else self.numel() // max(1, prod) for s in shape)
        n = 1
        for s in full:
            n *= s
        if n != self.numel():
            raise RuntimeError("shape %s is invalid for an observation of %d elements" % (shape, self.numel()))
        out = LazyObs(self.env, full)
        out.version = self.version
        return out

    view = reshape

    def to(self, *a, **k):
        return self.dense().to(*a, **k)

    def float(self):
        return self.dense()

    def contiguous(self):
        return self.dense()

    def cpu(self):
        return self.dense().cpu()

    def numpy(self):
        return self.dense().cpu().numpy()

    def __getitem__(self, idx):
        return self.dense()[idx]

    def __array__(self, dtype=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __repr__(self):
        return "LazyObs(shape=%s, env=%s, version=%d)" % (self.shape, type(self.env).__name__, self.version)
