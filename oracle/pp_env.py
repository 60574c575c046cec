"""Predator-prey environment, CPU restatement.  TEST INFRASTRUCTURE (oracle).

Follows ``ic3net-envs/ic3net_envs/predator_prey_env.py`` of the reference:
  multi_agent_init :72-110   class ids / vocab / naction
  reset            :146-168  (+ _get_cordinates :173-175)
  _take_action     :212-252  clamped moves, reached predators freeze, act 4 = stay
  _get_obs         :188-210  one-hot cell id + PREDATOR/PREY *counts* per window cell
  _get_reward      :254-290  mixed / cooperative / competitive; reached; episode_over; success
  step             :112-144
  reward_terminal  :292-293  zeros, but re-runs _get_reward (side effects)
and the flattening of ``env_wrappers.py:88-100`` ([N, W*W*V], window-major).

One instance is ONE environment (the reference cannot batch).  Spawn positions
come either from the caller (``reset(locs=...)``) or from the shared Philox
stream (``reset(seed=..., env_id=..., episode=...)``), see oracle/philox.py.
"""
import numpy as np

from . import philox

MODES = ("mixed", "cooperative", "competitive")


class PredatorPreyOracle(object):
    TIMESTEP_PENALTY = -0.05   # predator_prey_env.py:40
    PREY_REWARD = 0.0          # :41
    POS_PREY_REWARD = 0.05     # :42

    def __init__(self, nagents, dim, vision, mode="mixed", nenemies=1, no_stay=False, enemy_comm=False):
        """``nagents`` = predators (args.nfriendly).  ``enemy_comm``: the prey is one more AGENT of the policy -- it
        gets an observation row (:203-207) and a reward entry (:255, :276-281), its action is ignored (:214-217)."""
        if nenemies != 1:
            raise NotImplementedError("reference reward logic only works for one prey (:258)")
        if mode not in MODES:
            raise RuntimeError("Incorrect mode, Available modes: [cooperative|competitive|mixed]")
        self.n, self.dim, self.vision, self.mode = int(nagents), int(dim), int(vision), mode
        self.enemy_comm = bool(enemy_comm)
        self.na = self.n + (1 if self.enemy_comm else 0)          # rows of obs / reward (:203-207, :255)
        self.naction = 4 if no_stay else 5                       # :88-92
        base = self.dim * self.dim                                # :96
        self.OUTSIDE, self.PREY, self.PREDATOR = base + 1, base + 2, base + 3   # :97-99
        self.vocab_size = base + 4                                # :102
        self.W = 2 * self.vision + 1
        self.obs_dim = self.W * self.W * self.vocab_size          # env_wrappers.py:30-31
        self.episode_over = False
        self.stat = {}

    # ---- spawn -------------------------------------------------------------
    def sample_cells(self, seed, env_id, episode):
        """N+1 distinct cells by rejection from the Philox stream (same law as
        np.random.choice(D*D, N+1, replace=False) at :174, different stream)."""
        need, cells, blk = self.n + 1, [], 0
        ncell = self.dim * self.dim
        assert need <= ncell
        while len(cells) < need:
            for w in philox.draw_u24(seed, env_id, episode, philox.STREAM_PP_RESET, blk):
                cell = int((int(w) * ncell) >> 24)
                if cell not in cells:
                    cells.append(cell)
                    if len(cells) == need:
                        break
            blk += 1
        return np.array([[c // self.dim, c % self.dim] for c in cells], dtype=np.int64)

    def reset(self, locs=None, seed=None, env_id=0, episode=0):
        if locs is None:
            locs = self.sample_cells(seed, env_id, episode)
        locs = np.array(locs, dtype=np.int64).reshape(self.n + 1, 2)
        self.predator_loc = locs[: self.n].copy()
        self.prey_loc = locs[self.n:].copy()
        self.reached = np.zeros(self.n, dtype=np.int64)
        self.episode_over = False
        self.stat = {}
        return self.get_obs()

    # ---- dynamics ------------------------------------------------------------
    def _move(self, i, a):
        if self.reached[i]:                       # :221-222
            return
        r, c = self.predator_loc[i]
        d = self.dim - 1
        if a == 0:                                # UP    :228-232
            r = max(0, r - 1)
        elif a == 1:                              # RIGHT :234-239
            c = min(d, c + 1)
        elif a == 2:                              # DOWN  :241-246
            r = min(d, r + 1)
        elif a == 3:                              # LEFT  :248-252
            c = max(0, c - 1)
        # a == 4 (STAY) falls through every branch; the act==5 test at :225 never fires
        self.predator_loc[i] = (r, c)

    def _reward(self):
        reward = np.full(self.na, self.TIMESTEP_PENALTY)                # :255-256
        on = np.zeros(self.na, dtype=bool)
        on[: self.n] = np.all(self.predator_loc == self.prey_loc[0], axis=1)     # :258
        n_on = int(on.sum())
        if self.mode == "cooperative":
            reward[on] = self.POS_PREY_REWARD * n_on
        elif self.mode == "competitive":
            if n_on:
                reward[on] = self.POS_PREY_REWARD / n_on
        else:
            reward[on] = self.PREY_REWARD
        on = on[: self.n]
        reward[self.n:] = -1 * self.TIMESTEP_PENALTY if n_on == 0 else 0      # prey reward :276-281
        self.reached[on] = 1                                            # :271
        if self.mode == "mixed" and np.all(self.reached == 1):          # :273-274
            self.episode_over = True
        if self.mode != "competitive":                                  # :284-288
            self.stat["success"] = 1 if n_on == self.n else 0
        return reward

    def step(self, action):
        if self.episode_over:
            raise RuntimeError("Episode is done")                       # :129-130
        action = np.atleast_1d(np.asarray(action).squeeze())
        assert np.all(action <= self.naction)                           # :137 (sic)
        for i, a in enumerate(action[: self.n]):
            self._move(i, int(a))
        self.episode_over = False
        obs = self.get_obs()
        reward = self._reward()
        info = {"predator_locs": self.predator_loc, "prey_locs": self.prey_loc}
        return obs, reward, self.episode_over, info

    def reward_terminal(self):
        return np.zeros_like(self._reward())

    # ---- observation -----------------------------------------------------------
    def get_obs(self):
        """[N, W, W, V] int64 exactly like :188-210 (counts, not booleans)."""
        v, W, V, D = self.vision, self.W, self.vocab_size, self.dim
        obs = np.zeros((self.na, W, W, V), dtype=np.int64)
        pred_cnt = np.zeros((D, D), dtype=np.int64)
        prey_cnt = np.zeros((D, D), dtype=np.int64)
        for r, c in self.predator_loc:
            pred_cnt[r, c] += 1
        for r, c in self.prey_loc:
            prey_cnt[r, c] += 1
        rows = list(self.predator_loc) + (list(self.prey_loc) if self.enemy_comm else [])     # :198-207
        for i, (r, c) in enumerate(rows):
            for dy in range(W):
                for dx in range(W):
                    rr, cc = r - v + dy, c - v + dx
                    if 0 <= rr < D and 0 <= cc < D:
                        obs[i, dy, dx, rr * D + cc] = 1
                        obs[i, dy, dx, self.PREDATOR] = pred_cnt[rr, cc]
                        obs[i, dy, dx, self.PREY] = prey_cnt[rr, cc]
                    else:
                        obs[i, dy, dx, self.OUTSIDE] = 1
        return obs

    def flat_obs(self, obs=None):
        """env_wrappers.py:98-99: [N, O] float64."""
        obs = self.get_obs() if obs is None else obs
        return obs.reshape(self.na, -1).astype(np.float64)
