"""Batched lock-step environments and tabular agents on one GPU.

Thin host objects over the C ABI (include/sgk.h).  torch is used only for
device buffers and streams; all computation happens in libsgk.so's kernels.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import (ENV_BOAT, ENV_ISLAND, ENV_LAVA, ENV_SOKOBAN, ENV_SOKOBAN2, ENV_SUPER, ENV_TOMATO, ENV_WHISKY, Q_PRIVATE, Q_SHARED,
                   RNG_PHILOX, RNG_REPLAY, EnvStats, SgkError, check)

# ENV_MAP values at safe_grid_agents/parsing/parse.py:25,29,31
KIND_BY_ID = {"BoatRace-v0": ENV_BOAT, "SideEffectsSokoban-v0": ENV_SOKOBAN,
              "TomatoWatering-v0": ENV_TOMATO, "DistributionalShift-v0": ENV_LAVA,
              "IslandNavigation-v0": ENV_ISLAND, "AbsentSupervisor-v0": ENV_SUPER, "WhiskyGold-v0": ENV_WHISKY,
              "SideEffectsSokoban2-v0": ENV_SOKOBAN2}     # level 1 of the sokoban module (no ENV_MAP alias in the reference)
KIND_BY_ALIAS = {"boat": ENV_BOAT, "sokoban": ENV_SOKOBAN, "tomato": ENV_TOMATO, "lava": ENV_LAVA,
                 "island": ENV_ISLAND, "super": ENV_SUPER, "whisky": ENV_WHISKY, "sokoban2": ENV_SOKOBAN2}     # parse.py:22-37


TOTAL_KEYS = ("episodes", "sum_return", "sum_performance", "sum_margin_pos", "n_margin_pos",
              "max_return", "running_return", "max_performance", "max_margin")


def summarize_totals(tot):
    """avg / max of returns, safeties, margins, margins_support -- the payload
    of the reference's Evaluation/* scalars (common/utils/meters.py:96-106)."""
    n = tot["episodes"]
    if n <= 0:
        return {}
    out = {"returns": {"avg": tot["sum_return"] / n, "max": tot["max_return"]},
           "safeties": {"avg": tot["sum_performance"] / n, "max": tot["max_performance"]},
           "margins": {"avg": (tot["sum_return"] - tot["sum_performance"]) / n, "max": tot["max_margin"]}}
    if tot["n_margin_pos"] > 0:
        out["margins_support"] = {"avg": tot["sum_margin_pos"] / tot["n_margin_pos"], "max": tot["max_margin"]}
    out["episodes"] = n
    return out


def _kind(kind):
    if isinstance(kind, str):
        if kind in KIND_BY_ID:
            return KIND_BY_ID[kind]
        if kind in KIND_BY_ALIAS:
            return KIND_BY_ALIAS[kind]
        raise ValueError("unknown environment %r (in scope: %s)" % (kind, sorted(KIND_BY_ID)))
    return int(kind)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    """torch's current stream as a C pointer.  The two raw accessors skip the
    Python-side device bookkeeping of torch.cuda.current_stream() (15 us per
    call, several calls per episode on the single-environment drop-in path)."""
    if _raw_stream is not None and _raw_device is not None:
        return ctypes.c_void_p(_raw_stream(_raw_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class BatchedEnv:
    """`n_envs` lock-step copies of one gridworld (replaces N x gym.make,
    train.py:51-52).  Environment i has global id env_id0 + i."""

    def __init__(self, kind, n_envs, seed=0, env_id0=0, device=0):
        if not torch.cuda.is_available():
            raise SgkError("gridfast needs a CUDA device; there is no CPU fallback")
        self.L = _lib.load()
        self.kind = _kind(kind)
        self.n = int(n_envs)
        self.seed, self.env_id0 = int(seed), int(env_id0)
        self.device = torch.device("cuda", device)
        self.h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self.L.sgk_env_create(self.kind, self.n, self.env_id0, self.seed, device, ctypes.byref(self.h)))
        c, h, w, a = (ctypes.c_int() for _ in range(4))
        check(self.L.sgk_env_shape(self.h, c, h, w, a))
        self.shape = (c.value, h.value, w.value)
        self.hw = h.value * w.value
        self.n_actions = a.value
        self._replay = None
        self.t = 0   # agent-step index of the next lock-step

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self.L.sgk_env_destroy(h)

    # -- buffers -------------------------------------------------------------
    def _u8(self, *shape):
        return torch.empty(shape, dtype=torch.uint8, device=self.device)

    def _f64(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=self.device)

    # -- configuration -------------------------------------------------------
    def set_replay_words(self, words):
        """Replay raw MT19937 words ([n_envs, L] uint32) instead of Philox."""
        w = torch.as_tensor(np.ascontiguousarray(words, dtype=np.uint32).view(np.int32)).to(self.device)
        w = w.reshape(self.n, -1).contiguous()
        self._replay = w
        check(self.L.sgk_env_set_replay(self.h, _p(w), w.shape[1], _stream()))

    def set_trace(self, enabled=True):
        check(self.L.sgk_env_set_trace(self.h, int(enabled)))

    # -- the unfused env API ---------------------------------------------------
    def reset(self, mask=None, step=None, want_boards=True):
        boards = self._u8(self.n, self.hw) if want_boards else None
        step = self.t if step is None else step
        check(self.L.sgk_env_reset(self.h, _p(mask), step, _p(boards), _stream()))
        return boards

    def step(self, actions, step=None, out=None):
        """One lock-step.  actions: uint8 cuda tensor [n].  Returns
        (boards u8 [n,HW], reward f64 [n], hidden f64 [n] (NaN = None), done u8 [n])."""
        if out is None:
            out = (self._u8(self.n, self.hw), self._f64(self.n), self._f64(self.n), self._u8(self.n))
        boards, reward, hidden, done = out
        step = self.t if step is None else step
        check(self.L.sgk_env_step(self.h, _p(actions), step, _p(boards), _p(reward), _p(hidden), _p(done), _stream()))
        self.t = step + 1
        return boards, reward, hidden, done

    def actual_actions(self):
        """info["extra_observations"]["actual_actions"] of the last `step`
        (learn.py:42-47,74-78): uint8 cuda tensor [n]."""
        out = self._u8(self.n)
        check(self.L.sgk_env_actual_actions(self.h, _p(out), _stream()))
        return out

    def render(self):
        boards = self._u8(self.n, self.hw)
        check(self.L.sgk_env_render(self.h, _p(boards), _stream()))
        return boards

    def boards_to_f32(self, boards):
        n = boards.shape[0]
        out = torch.empty((n,) + self.shape, dtype=torch.float32, device=self.device)
        check(self.L.sgk_board_to_f32(self.h, _p(boards), _p(out), n, _stream()))
        return out

    def board_keys(self, boards):
        boards = boards.contiguous()
        n = boards.shape[0]
        out = torch.empty(n, dtype=torch.int64, device=self.device)
        check(self.L.sgk_board_to_key(self.h, _p(boards), _p(out), n, _stream()))
        return out

    def keys_to_boards(self, keys):
        """The boards table keys stand for (keys are lossless): int64 cuda
        tensor [n] -> uint8 [n, HW]."""
        keys = keys.contiguous()
        n = keys.shape[0]
        out = self._u8(n, self.hw)
        check(self.L.sgk_key_to_board(self.h, _p(keys), _p(out), n, _stream()))
        return out

    @property
    def max_iterations(self):
        return self.L.sgk_env_max_iterations(self.h)

    # -- bookkeeping -----------------------------------------------------------
    def stats(self):
        """Per-environment episode bookkeeping as a dict of cuda tensors."""
        f = {k: self._f64(self.n) for k in ("episode_return", "last_return", "last_performance",
                                            "sum_return", "sum_performance", "sum_margin_pos", "max_return",
                                            "max_performance", "max_margin")}
        i = {k: torch.empty(self.n, dtype=torch.int64, device=self.device)
             for k in ("episodes", "n_margin_pos", "trace_hash")}
        st = EnvStats(**{k: v.data_ptr() for k, v in {**f, **i}.items()})
        check(self.L.sgk_env_get_stats(self.h, ctypes.byref(st), _stream()))
        return {**f, **i}

    def stats_brief(self):
        """[3, n] float64: episode_return, last_return, last_performance (NaN =
        None) -- what track_metrics reads through env._env (meters.py:76-80)."""
        out = self._f64(3, self.n)
        st = EnvStats(episode_return=out[0].data_ptr(), last_return=out[1].data_ptr(),
                      last_performance=out[2].data_ptr())
        check(self.L.sgk_env_get_stats(self.h, ctypes.byref(st), _stream()))
        return out[:, 0] if self.n == 1 else out

    def totals(self):
        """Deterministic totals over all copies (synchronises)."""
        buf = (ctypes.c_double * 9)()
        check(self.L.sgk_env_totals_host(self.h, ctypes.byref(buf), _stream()))
        return dict(zip(TOTAL_KEYS, list(buf)))

    def clear_stats(self):
        check(self.L.sgk_env_clear_stats(self.h, _stream()))

    def totals_device(self, out=None):
        """The same 7 totals as a cuda float64 tensor, no synchronisation."""
        out = self._f64(9) if out is None else out
        check(self.L.sgk_env_totals(self.h, _p(out), _stream()))
        return out

    def core(self):
        out = torch.empty(self.n, dtype=torch.int64, device=self.device)
        check(self.L.sgk_env_get_core(self.h, _p(out), _stream()))
        return out

    def rollout_random(self, n_steps):
        check(self.L.sgk_rollout_random(self.h, n_steps, self.t, _stream()))
        self.t += n_steps


class BatchedTabularQ:
    """Tabular Q agent(s) for a BatchedEnv: private (one table per
    environment = N copies of the reference's TabularQAgent,
    common/agents/value.py:15-58) or one shared table."""

    def __init__(self, env, q_mode=Q_PRIVATE, capacity=0, lr=0.5, discount=0.99,
                 epsilon=0.01, epsilon_anneal=100000):
        self.L = env.L
        self.env = env
        self.q_mode = q_mode
        self.h = ctypes.c_void_p()
        with torch.cuda.device(env.device):
            check(self.L.sgk_tabq_create(env.h, q_mode, capacity, ctypes.byref(self.h)))
        self.n_tables = self.L.sgk_tabq_tables(self.h)
        self.configure(lr, discount, epsilon, epsilon_anneal)

    @property
    def capacity(self):
        """Slots per table right now (hashed private tables grow on demand)."""
        return self.L.sgk_tabq_capacity(self.h)

    def set_auto_grow(self, enabled=True):
        """On (default): tables grow like the reference's dict.  Off: fixed
        capacity, an overflow raises SgkError (SGK_EFULL) at the next check()."""
        check(self.L.sgk_tabq_set_auto_grow(self.h, int(enabled)))

    def max_fill(self):
        """Key count of the fullest table (synchronises)."""
        out = ctypes.c_int64()
        check(self.L.sgk_tabq_max_fill(self.h, ctypes.byref(out), _stream()))
        return out.value

    def grow(self, new_capacity):
        check(self.L.sgk_tabq_grow(self.h, new_capacity, _stream()))

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self.L.sgk_tabq_destroy(h)

    def configure(self, lr, discount, epsilon, epsilon_anneal):
        self.lr, self.discount, self.epsilon, self.epsilon_anneal = lr, discount, epsilon, epsilon_anneal
        check(self.L.sgk_tabq_configure(self.h, lr, discount, epsilon, epsilon_anneal))

    def enable_ssrl(self, c_prior, budget, max_episode_steps=100):
        """TabularSSQAgent.__init__ (ssrl/agents.py:16-27): C prior, query budget."""
        check(self.L.sgk_tabq_enable_ssrl(self.h, c_prior, budget, max_episode_steps))

    def ssrl_warmup(self, n_episodes, t0=None, want_steps=False):
        """ssrl.random_warmup (ssrl/warmup.py:4-35), batched: `n_episodes`
        random-policy episodes per environment, query_H + learn_C after each.
        Leaves every environment finished-and-not-reset, like the reference.
        The warm-up's lock-steps are indexed from `t0` (default: a region of
        the counter space training never reaches, so Philox draws are not
        reused when training then starts from agent-step 0)."""
        t0 = (1 << 40) if t0 is None else t0
        steps = torch.zeros(self.env.n, dtype=torch.int64, device=self.env.device) if want_steps else None
        check(self.L.sgk_ssrl_warmup(self.env.h, self.h, n_episodes, t0, _p(steps), _stream()))
        return steps

    def ssrl_counters(self):
        """(budget, episodes, corrupt_episodes) per environment, int64 cuda tensors."""
        dev = self.env.device
        out = [torch.empty(self.env.n, dtype=torch.int64, device=dev) for _ in range(3)]
        check(self.L.sgk_ssrl_get_counters(self.h, _p(out[0]), _p(out[1]), _p(out[2]), _stream()))
        return tuple(out)

    def epsilon_at(self, k):
        return self.L.sgk_tabq_epsilon_at(self.h, k)

    def act(self, boards, step, explore=False, out=None):
        n = boards.shape[0]
        out = self.env._u8(n) if out is None else out
        check(self.L.sgk_tabq_act(self.h, self.env.h, _p(boards), n, step, int(explore), _p(out), _stream()))
        return out

    def learn(self, boards, actions, rewards, successors):
        n = boards.shape[0]
        check(self.L.sgk_tabq_learn(self.h, _p(boards), _p(actions), _p(rewards), _p(successors), n, _stream()))

    def rollout(self, n_steps, cheat=False):
        """The fused hot path: n_steps lock-steps of act/step/learn/reset."""
        check(self.L.sgk_rollout_tabq(self.env.h, self.h, n_steps, self.env.t, int(cheat), _stream()))
        self.env.t += n_steps

    def rollout_episodes(self, max_episodes=1, max_steps=None, cheat=False, t0=None):
        """The fused path episode-wise (one reference call of tabq_learn =
        whiler.stepbystep, common/learn.py:13-24, per episode): every
        environment runs `max_episodes` episodes and is left un-reset.  Returns
        (steps int64 [n], last reward f64 [n], last hidden f64 [n] (NaN = None))
        as cuda tensors; the agent-step clock advances by max(steps) when N = 1."""
        dev = self.env.device
        max_steps = max_episodes * self.env.max_iterations if max_steps is None else max_steps
        steps = torch.zeros(self.env.n, dtype=torch.int64, device=dev)
        reward = torch.zeros(self.env.n, dtype=torch.float64, device=dev)
        hidden = torch.zeros(self.env.n, dtype=torch.float64, device=dev)
        t0 = self.env.t if t0 is None else t0
        check(self.L.sgk_rollout_tabq_episodes(self.env.h, self.h, max_episodes, max_steps, t0, int(cheat),
                                               _p(steps), _p(reward), _p(hidden), _stream()))
        return steps, reward, hidden

    def check(self):
        check(self.L.sgk_check(self.env.h, self.h, _stream()))

    def evaluate(self, eval_env, eval_timesteps=2000):
        """default_eval (common/eval.py:8-56) on `eval_env` (a BatchedEnv of the
        same kind; same size as the training set for private tables): greedy,
        read-only.  Returns avg/max of returns, safeties, margins,
        margins_support over the evaluation episodes."""
        eval_env.clear_stats()
        eval_env.reset(step=eval_env.t, want_boards=False)
        check(self.L.sgk_eval_tabq(eval_env.h, self.h, eval_timesteps, eval_env.t, _stream()))
        eval_env.t += eval_timesteps + eval_env.max_iterations
        return summarize_totals(eval_env.totals())

    def evaluate_logged(self, eval_env, eval_timesteps, insert_on_miss=True):
        """default_eval for the drop-in adapters: `eval_env` must already be
        reset; act() inserts unseen boards like the reference's defaultdict
        (value.py:31,35); returns the (return, performance) pairs of the
        evaluation episodes of environment 0 in order (host numpy [m, 2])."""
        cap = eval_timesteps + eval_env.max_iterations
        log = torch.full((eval_env.n, cap, 2), float("nan"), dtype=torch.float64, device=eval_env.device)
        check(self.L.sgk_eval_tabq_ex(eval_env.h, self.h, eval_timesteps, eval_env.t, int(insert_on_miss),
                                      _p(log), cap, _stream()))
        eval_env.t += cap
        rows = log[0].cpu().numpy()
        return rows[~np.isnan(rows[:, 0])]

    def export(self, table=0, with_corruption=False):
        """(keys int64 [m], rows f64 [m,4]) of the occupied slots (host numpy)."""
        dev = self.env.device
        keys = torch.empty(self.capacity, dtype=torch.int64, device=dev)
        rows = torch.empty(self.capacity, 4, dtype=torch.float64, device=dev)
        corr = torch.empty(self.capacity, dtype=torch.float64, device=dev) if with_corruption else None
        check(self.L.sgk_tabq_export(self.h, table, _p(keys), _p(rows), _p(corr), _stream()))
        keys, rows = keys.cpu().numpy().view(np.uint64), rows.cpu().numpy()
        used = keys != 0
        if with_corruption:
            return keys[used], rows[used], corr.cpu().numpy()[used]
        return keys[used], rows[used]

    def import_table(self, table, keys_full, rows_full):
        dev = self.env.device
        k = torch.as_tensor(np.ascontiguousarray(keys_full).view(np.int64)).to(dev)
        r = torch.as_tensor(np.ascontiguousarray(rows_full, dtype=np.float64)).to(dev)
        assert k.numel() == self.capacity and r.shape == (self.capacity, 4)
        check(self.L.sgk_tabq_import(self.h, table, _p(k), _p(r), _stream()))
        torch.cuda.current_stream().synchronize()

    # -- replica sync of a shared table (see gridfast.distributed) -------------
    def delta_export(self):
        dev = self.env.device
        keys = torch.empty(self.capacity, dtype=torch.int64, device=dev)
        delta = torch.empty(self.capacity, 4, dtype=torch.float64, device=dev)
        check(self.L.sgk_tabq_delta_export(self.h, _p(keys), _p(delta), _stream()))
        return keys, delta

    def delta_apply(self, keys, delta, scale):
        check(self.L.sgk_tabq_delta_apply(self.h, _p(keys), _p(delta), float(scale), _stream()))

    def rebase(self):
        check(self.L.sgk_tabq_rebase(self.h, _stream()))

    def dense_size(self):
        """Entries of the canonical dense delta array (0: none, use the key/delta records)."""
        return self.L.sgk_tabq_dense_size(self.h)

    def delta_export_dense(self, out=None):
        n = self.dense_size()
        out = torch.empty(n, 5, dtype=torch.float64, device=self.env.device) if out is None else out
        check(self.L.sgk_tabq_delta_export_dense(self.h, _p(out), _stream()))
        return out

    def delta_apply_dense(self, delta_sum, scale):
        check(self.L.sgk_tabq_delta_apply_dense(self.h, _p(delta_sum), float(scale), _stream()))

    def restore_base(self):
        check(self.L.sgk_tabq_restore_base(self.h, _stream()))

    def rollout_host(self, n_steps, core_in, core_out, boards_out, cheat=False):
        """Host-buffer form: pinned numpy/torch CPU buffers in and out."""
        totals = (ctypes.c_double * 9)()
        check(self.L.sgk_rollout_tabq_host(
            self.env.h, self.h, n_steps, self.env.t, int(cheat),
            None if core_in is None else ctypes.c_void_p(core_in.data_ptr()),
            None if core_out is None else ctypes.c_void_p(core_out.data_ptr()),
            None if boards_out is None else ctypes.c_void_p(boards_out.data_ptr()),
            ctypes.byref(totals), _stream()))
        self.env.t += n_steps
        return list(totals)
