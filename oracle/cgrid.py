"""ctypes binding of oracle/cgrid.c (test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcgrid.so")

BOAT, SOKOBAN, TOMATO, LAVA, ISLAND, SUPER, WHISKY, SOKOBAN2 = 0, 1, 2, 3, 4, 5, 6, 7
KIND_BY_ID = {"BoatRace-v0": BOAT, "SideEffectsSokoban-v0": SOKOBAN, "TomatoWatering-v0": TOMATO,
              "DistributionalShift-v0": LAVA, "IslandNavigation-v0": ISLAND,
              "AbsentSupervisor-v0": SUPER, "WhiskyGold-v0": WHISKY, "SideEffectsSokoban2-v0": SOKOBAN2}
SHAPE = {BOAT: (5, 5), SOKOBAN: (6, 6), TOMATO: (7, 9), LAVA: (7, 9), ISLAND: (6, 8), SUPER: (6, 8),
         WHISKY: (6, 8), SOKOBAN2: (10, 10)}
RNG_PHILOX, RNG_REPLAY = 0, 1
Q_PRIVATE, Q_SHARED = 0, 1

_lib = None


def build(force=False):
    src = os.path.join(HERE, "cgrid.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "libcgrid.so"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        vp, i64, u64, i32, dbl = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64,
                                  ctypes.c_int, ctypes.c_double)
        L.cg_create.restype = vp
        L.cg_create.argtypes = [i32, i64, i64, u64, i32, i32, vp, i64]
        L.cg_destroy.argtypes = [vp]
        L.cg_set_agent.argtypes = [vp, dbl, dbl, dbl, i64, i32]
        L.cg_set_ssrl.argtypes = [vp, i32, dbl, i64]
        L.cg_rollout.restype = i32
        L.cg_rollout.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.cg_rollout_random.restype = i32
        L.cg_rollout_random.argtypes = [vp, i64]
        L.cg_step_actions.restype = i32
        L.cg_step_actions.argtypes = [vp, vp, vp, vp, vp, vp]
        L.cg_hw.restype = i32
        L.cg_hw.argtypes = [vp]
        L.cg_t.restype = i64
        L.cg_t.argtypes = [vp]
        L.cg_get_boards.argtypes = [vp, vp]
        L.cg_get_env_stats.argtypes = [vp, vp, vp, vp]
        L.cg_table_size.restype = i64
        L.cg_table_size.argtypes = [vp, i64]
        L.cg_table_export.restype = i64
        L.cg_table_export.argtypes = [vp, i64, vp, vp, vp]
        L.cg_eval.argtypes = [vp, u64, i64, i64, i64, u64, vp, vp]
        L.cg_epsilon_at.restype = dbl
        L.cg_epsilon_at.argtypes = [dbl, i64, i64]
        L.cg_philox.argtypes = [vp, vp, vp]
        L.cg_set_threads.argtypes = [i32]
        L.cg_ssrl_warmup.restype = i32
        L.cg_ssrl_warmup.argtypes = [vp, i64, u64, vp]
        L.cg_get_ssrl_counters.argtypes = [vp, vp, vp, vp]
        L.cg_clear_stats.argtypes = [vp]
        L.cg_set_t.argtypes = [vp, i64]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Sim:
    """N lock-step (env, tabular agent) pairs on the CPU."""

    def __init__(self, kind, n_envs, seed=0, q_mode=Q_PRIVATE, rng_mode=RNG_PHILOX,
                 replay_words=None, env_id0=0, lr=0.5, discount=0.99, epsilon=0.01,
                 epsilon_anneal=100000, cheat=False, ssrl=False, c_prior=0.01, budget=0):
        self.L = lib()
        self.kind, self.n = kind, n_envs
        self.q_mode = q_mode
        self.hw = SHAPE[kind][0] * SHAPE[kind][1]
        self._words = None
        wpe = 0
        if rng_mode == RNG_REPLAY:
            self._words = np.ascontiguousarray(replay_words, dtype=np.uint32).reshape(n_envs, -1)
            wpe = self._words.shape[1]
        self.h = self.L.cg_create(kind, n_envs, env_id0, seed, q_mode, rng_mode, _ptr(self._words), wpe)
        self.L.cg_set_agent(self.h, lr, discount, epsilon, epsilon_anneal, int(cheat))
        if ssrl:
            self.L.cg_set_ssrl(self.h, 1, c_prior, budget)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.cg_destroy(self.h)
            self.h = None

    @property
    def t(self):
        return self.L.cg_t(self.h)

    @t.setter
    def t(self, value):
        self.L.cg_set_t(self.h, value)

    def rollout(self, n_steps, trace=False, boards=False):
        out = {}
        a = r = h = d = b = None
        if trace:
            a = np.zeros((n_steps, self.n), np.uint8)
            r = np.zeros((n_steps, self.n), np.float64)
            h = np.zeros((n_steps, self.n), np.float64)
            d = np.zeros((n_steps, self.n), np.uint8)
            out = dict(actions=a, reward=r, hidden=h, done=d)
        if boards:
            b = np.zeros((n_steps, self.n, self.hw), np.uint8)
            out["boards"] = b
        rc = self.L.cg_rollout(self.h, n_steps, _ptr(a), _ptr(r), _ptr(h), _ptr(d), _ptr(b))
        if rc:
            raise RuntimeError("replay word stream exhausted")
        return out

    def ssrl_warmup(self, n_episodes, t0=1 << 40):
        """ssrl/warmup.py:4-35 for every environment; returns steps taken per environment."""
        steps = np.zeros(self.n, np.int64)
        if self.L.cg_ssrl_warmup(self.h, n_episodes, t0, _ptr(steps)):
            raise RuntimeError("replay word stream exhausted")
        return steps

    def ssrl_counters(self):
        out = [np.zeros(self.n, np.int64) for _ in range(3)]
        self.L.cg_get_ssrl_counters(self.h, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]))
        return tuple(out)

    def clear_stats(self):
        self.L.cg_clear_stats(self.h)

    def rollout_random(self, n_steps):
        if self.L.cg_rollout_random(self.h, n_steps):
            raise RuntimeError("replay word stream exhausted")

    def step(self, actions):
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        r = np.zeros(self.n, np.float64)
        h = np.zeros(self.n, np.float64)
        d = np.zeros(self.n, np.uint8)
        b = np.zeros((self.n, self.hw), np.uint8)
        if self.L.cg_step_actions(self.h, _ptr(actions), _ptr(r), _ptr(h), _ptr(d), _ptr(b)):
            raise RuntimeError("replay word stream exhausted")
        return b, r, h, d

    def boards(self):
        b = np.zeros((self.n, self.hw), np.uint8)
        self.L.cg_get_boards(self.h, _ptr(b))
        return b

    def env_stats(self):
        f = np.zeros((self.n, 10), np.float64)
        i = np.zeros((self.n, 5), np.int64)
        hsh = np.zeros(self.n, np.uint64)
        self.L.cg_get_env_stats(self.h, _ptr(f), _ptr(i), _ptr(hsh))
        return dict(episode_return=f[:, 0], hidden_cum=f[:, 1], last_return=f[:, 2],
                    last_perf=f[:, 3], sum_return=f[:, 4], sum_perf=f[:, 5],
                    sum_margin_pos=f[:, 6], max_return=f[:, 7], max_perf=f[:, 8], max_margin=f[:, 9],
                    episodes=i[:, 0],
                    n_margin_pos=i[:, 1], frame=i[:, 2], perf_defined=i[:, 3],
                    hidden_defined=i[:, 4], trace_hash=hsh)

    def evaluate(self, seed, env_id0, n_eval, eval_timesteps, t0=0):
        f = np.zeros((n_eval, 6), np.float64)
        i = np.zeros((n_eval, 2), np.int64)
        self.L.cg_eval(self.h, seed, env_id0, n_eval, eval_timesteps, t0, _ptr(f), _ptr(i))
        return dict(sum_return=f[:, 0], sum_perf=f[:, 1], sum_margin_pos=f[:, 2], max_return=f[:, 3],
                    max_perf=f[:, 4], max_margin=f[:, 5], episodes=i[:, 0], n_margin_pos=i[:, 1])

    def table(self, index=0, with_c=False):
        n = self.L.cg_table_size(self.h, index)
        keys = np.zeros((max(n, 1), self.hw), np.uint8)
        q = np.zeros((max(n, 1), 4), np.float64)
        c = np.zeros(max(n, 1), np.float64)
        m = self.L.cg_table_export(self.h, index, _ptr(keys), _ptr(q), _ptr(c))
        assert m == n
        if with_c:
            return keys[:n], q[:n], c[:n]
        return keys[:n], q[:n]


def set_threads(n):
    """Host threads for private-table rollouts (default: all online cores)."""
    lib().cg_set_threads(n)


def epsilon_at(epsilon, anneal, k):
    return lib().cg_epsilon_at(epsilon, anneal, k)
