"""Traffic-junction environment, CPU restatement.  TEST INFRASTRUCTURE (oracle).

Follows ``ic3net-envs/ic3net_envs/traffic_junction_env.py`` of the reference:
  multi_agent_init :80-158   dims (+1 for easy), BASE/OUTSIDE/CAR ids, vocab, npath
  reset            :160-204  (+ curriculum :620-626)
  step             :206-252  order: is_completed=0, _take_action, _add_cars, obs, reward
  _take_action     :540-581
  _add_cars        :369-393  (+ _choose_dead :614-618) sequential over arrival groups
  _get_obs         :321-366  (act, route id, window one-hot + car COUNT incl. dead cars at (0,0))
  _get_reward      :585-595
and the flattening of ``env_wrappers.py:88-100`` ([N, 2 + W*W*V]).

The static tables (road-id grid, routes) are INPUTS: in tests they come from the
golden fixtures produced by the unmodified reference (tests/golden/tj_tables_*.npz),
so this file never re-derives ``traffic_helper.get_routes``.

Randomness: every arrival group g visited at env step ``tick`` owns three 24-bit
draws (spawn test, dead-slot choice, path choice).  They come from the Philox
stream (oracle/philox.py) or from an explicit per-step array ``draws[G,3]``.
"""
import math

import numpy as np

from . import philox

TIMESTEP_PENALTY = -0.01   # traffic_junction_env.py:44
CRASH_PENALTY = -10.0      # :45


def constants(difficulty, dim):
    """(dims, BASE, OUTSIDE, CAR, vocab, npath) per :103-133."""
    dims = (dim + 1, dim + 1) if difficulty == "easy" else (dim, dim)     # :112-115
    nroad = {"easy": 2, "medium": 4, "hard": 8}[difficulty]
    base = {"easy": 1, "medium": 2, "hard": 4}[difficulty] * (dim + dim)    # :121-124 (original dim)
    npath = math.factorial(nroad) // math.factorial(nroad - 2)             # :126
    return dims, base, base, base + 2, base + 3, npath


class TrafficJunctionOracle(object):
    def __init__(self, nagents, dim, vision, difficulty, tables,
                 add_rate_min=0.05, add_rate_max=0.2, curr_start=0, curr_end=0):
        self.n, self.dim, self.vision, self.difficulty = int(nagents), int(dim), int(vision), difficulty
        (self.dims, self.BASE, self.OUTSIDE, self.CAR, self.vocab_size,
         self.npath) = constants(difficulty, self.dim)
        self.naction = 2
        self.grid = np.asarray(tables["grid"], dtype=np.int64)
        assert self.grid.shape == tuple(self.dims)
        self.routes = [[np.asarray(p, dtype=np.int64) for p in grp] for grp in tables["routes"]]
        assert sum(len(g) for g in self.routes) == self.npath               # :520
        self.W = 2 * self.vision + 1
        self.obs_dim = 2 + self.W * self.W * self.vocab_size                # env_wrappers.py:21-29
        self.add_rate_min, self.add_rate_max = add_rate_min, add_rate_max
        self.curr_start, self.curr_end = curr_start, curr_end
        self.exact_rate = self.add_rate = add_rate_min                      # :103
        self.epoch_last_update = 0
        self.pad = np.pad(self.grid, self.vision, "constant", constant_values=self.OUTSIDE)
        self.tick = 0
        self.stat = {}
        self.episode_over = False

    # ---- reset / curriculum ------------------------------------------------
    def reset(self, epoch=None):
        n = self.n
        self.episode_over = False
        self.has_failed = 0
        self.alive = np.zeros(n, dtype=np.int64)
        self.wait = np.zeros(n, dtype=np.int64)
        self.cars_in_sys = 0
        self.route_id = np.full(n, -1, dtype=np.int64)
        self.path_grp = np.zeros(n, dtype=np.int64)
        self.path_idx = np.zeros(n, dtype=np.int64)
        self.car_loc = np.zeros((n, 2), dtype=np.int64)
        self.last_act = np.zeros(n, dtype=np.int64)
        self.route_loc = np.full(n, -1, dtype=np.int64)
        self.is_completed = np.zeros(n, dtype=np.int64)
        self.stat = {}
        rng_e = self.curr_end - self.curr_start
        rng_r = self.add_rate_max - self.add_rate_min
        if epoch is not None and rng_e > 0 and rng_r > 0 and epoch > self.epoch_last_update:   # :197
            self.curriculum(epoch)
            self.epoch_last_update = epoch
        return self.get_obs()

    def curriculum(self, epoch):
        step = (self.add_rate_max - self.add_rate_min) / (self.curr_end - self.curr_start)
        if self.curr_start <= epoch < self.curr_end:
            self.exact_rate = self.exact_rate + step
            self.add_rate = 0.01 * (self.exact_rate // 0.01)                # :626

    def spawn_threshold(self):
        """u <= add_rate  <=>  u24 <= floor(add_rate * 2**24)  for u = u24 * 2**-24."""
        return int(math.floor(self.add_rate * (2.0 ** 24)))

    # ---- dynamics --------------------------------------------------------------
    def _advance(self, i, a):
        if not self.alive[i]:                                     # :542-543
            return
        self.wait[i] += 1                                         # :546
        if a == 1:                                                # BRAKE :549-551
            self.last_act[i] = 1
            return
        if a == 0:                                                # GAS :554
            self.route_loc[i] += 1
            path = self.routes[self.path_grp[i]][self.path_idx[i]]
            k = self.route_loc[i]
            if k == len(path):                                    # :560-568
                self.cars_in_sys -= 1
                self.alive[i] = 0
                self.wait[i] = 0
                self.car_loc[i] = 0
                self.is_completed[i] = 1
                return
            if k > len(path):
                raise RuntimeError("Out of boud car path")        # :570-572
            self.car_loc[i] = path[k]
            self.last_act[i] = 0

    def _spawn(self, draws, seed, env_id):
        thr = self.spawn_threshold()
        for g, paths in enumerate(self.routes):
            if self.cars_in_sys >= self.n:                        # :371-372
                return
            if draws is not None:
                w = [int(x) for x in draws[g]]
            else:
                w = [int(x) for x in philox.draw_u24(seed, env_id, self.tick, philox.STREAM_TJ_SPAWN, g)]
            if w[0] <= thr:                                       # :375
                dead = np.flatnonzero(self.alive == 0)            # :614-618
                idx = int(dead[(w[1] * len(dead)) >> 24])
                self.alive[idx] = 1
                p = (w[2] * len(paths)) >> 24                     # :383
                self.route_id[idx] = p + g * len(paths)           # :385
                self.path_grp[idx], self.path_idx[idx] = g, p
                self.route_loc[idx] = 0
                self.car_loc[idx] = paths[p][0]
                self.cars_in_sys += 1

    def _reward(self):
        reward = TIMESTEP_PENALTY * self.wait.astype(np.float64)
        for i in range(self.n):
            l = self.car_loc[i]
            if l.any():
                same = np.all(self.car_loc == l, axis=1)
                same[i] = False
                if same.any():
                    reward[i] += CRASH_PENALTY
                    self.has_failed = 1
        return self.alive * reward

    def step(self, action, draws=None, seed=None, env_id=0):
        if self.episode_over:
            raise RuntimeError("Episode is done")
        action = np.asarray(action).squeeze()
        assert np.all(action <= self.naction)
        assert len(action) == self.n
        self.is_completed = np.zeros(self.n, dtype=np.int64)
        for i, a in enumerate(action):
            self._advance(i, int(a))
        self._spawn(draws, seed, env_id)
        self.tick += 1
        obs = self.get_obs()
        reward = self._reward()
        info = {"car_loc": self.car_loc, "alive_mask": self.alive.astype(np.float64).copy(),
                "wait": self.wait, "cars_in_sys": self.cars_in_sys,
                "is_completed": self.is_completed.astype(np.float64).copy()}
        self.stat["success"] = 1 - self.has_failed
        self.stat["add_rate"] = self.add_rate
        return obs, reward, self.episode_over, info

    def reward_terminal(self):
        return np.zeros_like(self._reward())

    # ---- observation -------------------------------------------------------------
    def get_obs(self):
        """[N, 2 + W*W*V] float64 (already flattened like env_wrappers.py:88-98)."""
        n, v, W, V = self.n, self.vision, self.W, self.vocab_size
        h, w = self.dims
        cnt = np.zeros((h + 2 * v, w + 2 * v), dtype=np.int64)
        for r, c in self.car_loc:                                 # every slot, dead ones sit at (0,0)
            cnt[r + v, c + v] += 1
        out = np.zeros((n, self.obs_dim), dtype=np.float64)
        for i in range(n):
            if not self.alive[i]:                                 # :352-356
                continue
            r, c = self.car_loc[i]
            win = np.zeros((W, W, V), dtype=np.float64)
            for dy in range(W):
                for dx in range(W):
                    win[dy, dx, self.pad[r + dy, c + dx]] = 1
                    win[dy, dx, self.CAR] += cnt[r + dy, c + dx]
            out[i, 0] = self.last_act[i] / (self.naction - 1)
            out[i, 1] = self.route_id[i] / (self.npath - 1)
            out[i, 2:] = win.reshape(-1)
        return out
