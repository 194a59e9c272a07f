"""Single-environment adapters: the duck-typed surface the reference drives.

`GridworldEnv` stands where `gym.make(ENV_MAP[alias])` stands in the reference
(train.py:51-52) and answers every member the reference touches (SURVEY.md
section 8b): seed, reset, step -> (board float32 (1,H,W), reward, done, info),
action_space.n, observation_space.shape, render(mode="rgb_array"),
_env.episode_return, _env.get_last_performance().

`GpuTabularQAgent` stands where `TabularQAgent(env, args)` stands
(common/agents/value.py:15-58): act, act_explore, learn, update_epsilon,
attributes Q, epsilon, action_n, discount, lr.

Both are N == 1 views of the batched engine; all dynamics, table lookups and
updates run in the CUDA kernels.  What stays on the host is only the plumbing
the reference also keeps on the host: the random draws come from numpy's
global stream when rng="numpy" (exactly the calls the reference makes:
np.random.sample / np.random.choice at value.py:38-39), so that a run seeded
with np.random.seed(s) (train.py:32) follows the reference's own stream.
"""
import numpy as np
import torch

from . import batched
from ._lib import Q_PRIVATE, SgkError

_MAX_WORDS_PER_CALL = 32      # >= 2 words x 13 tomatoes
# raw words one fused lock-step can consume: act_explore 2 + 1 (value.py:38-39),
# the environment up to 2 x 13 (tomato drying draws) or 2 + 1 (whisky)
_WORDS_PER_FUSED_STEP = 3 + 26

# board values of the tiles that end an episode (the wrapper reports discount 0.0 there)
_TERMINAL_VALUES = {batched.ENV_SOKOBAN: (5,), batched.ENV_SOKOBAN2: (), batched.ENV_LAVA: (3, 4), batched.ENV_ISLAND: (3, 4),
                    batched.ENV_SUPER: (5,), batched.ENV_WHISKY: (4,)}
# environments whose step / reset consume draws of the global numpy stream
_STOCHASTIC = (batched.ENV_TOMATO, batched.ENV_SUPER, batched.ENV_WHISKY)

# colours for render(mode="rgb_array") only; not part of any parity claim
_PALETTE = np.array([[152, 152, 152], [219, 219, 219], [0, 180, 255], [0, 210, 50],
                     [153, 102, 51], [255, 0, 255]], dtype=np.uint8)


class _Discrete:
    def __init__(self, n):
        self.n = n

    def sample(self):
        return int(np.random.randint(0, self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class _Box:
    def __init__(self, shape):
        self.shape = shape
        self.dtype = np.float32


def _as_int_action(action):
    # numpy.int64 from argmax, int from np.random.choice, or a 1-element tensor
    # from DeepQAgent.act (value.py:35,39,92)
    if hasattr(action, "item"):
        return int(action.item())
    return int(action)


class _InnerEnv:
    """What the reference reaches through `env._env` (meters.py:67-80)."""

    def __init__(self, outer):
        self._outer = outer

    @property
    def episode_return(self):
        return self._outer._episode_return

    def get_last_performance(self):
        return self._outer._last_performance


class GridworldEnv:
    metadata = {"render.modes": ["human", "ansi", "rgb_array"]}

    def __init__(self, env_id, seed=0, device=0, rng="numpy", env_index=0):
        if rng not in ("numpy", "philox"):
            raise ValueError("rng must be 'numpy' or 'philox'")
        self._env_id = env_id
        self._device = device
        self._rng = rng
        self._env_index = env_index
        self._make(seed)
        self._env = _InnerEnv(self)
        self.action_space = _Discrete(self.batched.n_actions)
        self.observation_space = _Box(self.batched.shape)

    def _make(self, seed):
        self.batched = batched.BatchedEnv(self._env_id, 1, seed=seed, env_id0=self._env_index,
                                          device=self._device)
        self._t = 0
        self._episode_return = 0
        self._last_performance = None
        self._board = None
        self._actions = torch.zeros(1, dtype=torch.uint8, device=self.batched.device)
        self._out = (self.batched._u8(1, self.batched.hw), self.batched._f64(1),
                     self.batched._f64(1), self.batched._u8(1))
        self._cursor = torch.zeros(1, dtype=torch.int64, device=self.batched.device)
        self._stochastic = self.batched.kind in _STOCHASTIC
        self._terminal = None
        self._water = None

    def seed(self, seed=None):
        """train.py:52.  Philox mode re-keys the streams; numpy mode seeds the
        global numpy stream like the wrapper the reference uses."""
        if self._rng == "philox":
            self._make(0 if seed is None else seed)
        else:
            np.random.seed(seed)
        return [seed]

    # -- numpy-stream plumbing: lend the kernel the next raw words of the global
    # stream, then advance the stream by exactly what the kernel consumed
    def _lend_words(self, n_words=_MAX_WORDS_PER_CALL, always=False):
        if self._rng != "numpy" or not (self._stochastic or always):
            return None
        state = np.random.get_state()
        words = np.random.randint(0, 2 ** 32, size=n_words, dtype=np.uint32)
        self.batched.set_replay_words(words.reshape(1, -1))
        return state

    def _settle_words(self, state):
        if state is None:
            return
        batched.check(self.batched.L.sgk_env_replay_cursor(
            self.batched.h, batched._p(self._cursor), batched._stream()))
        used = int(self._cursor.item())
        np.random.set_state(state)
        if used:
            np.random.randint(0, 2 ** 32, size=used, dtype=np.uint32)

    def _observation(self, boards):
        self._board = boards[0].cpu().numpy()
        c, h, w = self.batched.shape
        return self._board.astype(np.float32).reshape(c, h, w)

    def reset(self):
        state = self._lend_words()
        boards = self.batched.reset(step=self._t)
        self._settle_words(state)
        self._episode_return = 0
        obs = self._observation(boards)
        self._read_static_tiles(self._board)
        return obs

    def _read_static_tiles(self, board):
        # once, off a board of a running episode (the agent never covers one of these tiles then)
        if self._terminal is None:
            self._terminal = np.isin(board, _TERMINAL_VALUES.get(self.batched.kind, ()))
            if self.batched.kind == batched.ENV_ISLAND:
                w = self.batched.shape[2]
                self._water = np.array([(c // w, c % w) for c in np.flatnonzero(board == 3)])

    def _extras(self, action, known=False):
        """info["extra_observations"] (learn.py:42-47,74-78)."""
        extra = {"actual_actions": action}
        if self.batched.kind == batched.ENV_WHISKY and not known:
            extra["actual_actions"] = int(self.batched.actual_actions()[0].item())
        if self._water is not None:
            w = self.batched.shape[2]
            cell = int(np.flatnonzero(self._board == 2)[0])
            extra["safety"] = int(np.min(np.abs(self._water[:, 0] - cell // w) + np.abs(self._water[:, 1] - cell % w)))
        return extra

    def step(self, action):
        action = _as_int_action(action)
        if not 0 <= action < self.batched.n_actions:
            raise ValueError("action %r outside the action space" % (action,))
        self._actions[0] = action
        if self._terminal is None:
            self._read_static_tiles(self.batched.render()[0].cpu().numpy())
        state = self._lend_words()
        boards, reward, hidden, done = self.batched.step(self._actions, step=self._t, out=self._out)
        self._settle_words(state)
        self._t += 1
        host = torch.stack([reward[0], hidden[0], done[0].double()]).cpu().numpy()
        reward, hidden, done = float(host[0]), float(host[1]), bool(host[2])
        hidden = None if hidden != hidden else hidden
        stats = self.batched.stats()
        self._episode_return = float(stats["episode_return"][0].item())
        if done:
            self._last_performance = float(stats["last_performance"][0].item())
        obs = self._observation(boards)
        return obs, reward, done, self._info(reward, hidden, done, None, action)

    def _info(self, reward, hidden, done, actual_action, action=None):
        """The wrapper's info dict (learn.py:42-47,72-78) for the current board."""
        if self._terminal is None:
            raise SgkError("step() before the first reset()")
        terminated = done and bool(self._terminal[np.flatnonzero(self._board == 2)[0]])
        info = {"hidden_reward": hidden, "observed_reward": reward,
                "discount": 0.0 if terminated else 1.0,
                "extra_observations": self._extras(actual_action if action is None else action,
                                                   known=actual_action is not None)}
        if done:
            info["extra_observations"]["termination_reason"] = 0 if terminated else 1
        return info

    # -- after a fused call (gridfast.loops): refresh what the reference reads
    def _after_fused(self, steps, done):
        """The kernels advanced the environment by `steps` steps on their own;
        bring the host-side mirror (board, _env.episode_return,
        _env.get_last_performance()) up to date.  Returns the observation."""
        self._t += int(steps)
        st = self.batched.stats_brief().cpu().numpy()      # [episode_return, last_return, last_performance]
        self._episode_return = float(st[1] if done else st[0])
        if not np.isnan(st[2]):
            self._last_performance = float(st[2])
        return self._observation(self.batched.render())

    def render(self, mode="human", close=False):
        if self._board is None:
            self._observation(self.batched.render())
        c, h, w = self.batched.shape
        if mode == "rgb_array":
            return np.moveaxis(_PALETTE[self._board.reshape(h, w)], -1, 0)
        text = "\n".join("".join("# A*XG"[v] for v in row) for row in self._board.reshape(h, w))
        if mode == "ansi":
            return text
        print(text)


def make(env_id, **kwargs):
    """Stand-in for gym.make (train.py:51)."""
    return GridworldEnv(env_id, **kwargs)


class _QView:
    """Read-only dict-like view of the device table, keyed like the
    reference's Q (tuple of the flattened float32 board, value.py:34)."""

    def __init__(self, agent):
        self._agent = agent

    def _snapshot(self):
        keys, rows = self._agent.table.export(0)
        if len(keys) == 0:
            return {}
        env = self._agent.env.batched
        boards = env.keys_to_boards(torch.as_tensor(keys.view(np.int64)).to(env.device)).cpu().numpy()
        return {tuple(np.float32(v) for v in boards[i]): rows[i] for i in range(len(keys))}

    def __len__(self):
        return len(self._snapshot())

    def __iter__(self):
        return iter(self._snapshot())

    def __getitem__(self, key):
        snap = self._snapshot()
        key = tuple(np.float32(v) for v in key)
        if key not in snap:
            return np.zeros(self._agent.action_n)
        return snap[key]

    def items(self):
        return self._snapshot().items()

    def keys(self):
        return self._snapshot().keys()


class GpuTabularQAgent:
    """Drop-in for TabularQAgent (common/agents/value.py:15-58); constructor
    signature (env, args) as AGENT_MAP classes have (train.py:54)."""

    def __init__(self, env, args):
        if not isinstance(env, GridworldEnv):
            raise SgkError("GpuTabularQAgent needs a gridfast GridworldEnv")
        self.env = env
        self.action_n = env.action_space.n
        self.discount = args.discount
        self.lr = args.lr
        self._final_epsilon = args.epsilon
        self._anneal = args.epsilon_anneal
        self._rng = env._rng
        self.table = batched.BatchedTabularQ(
            env.batched, q_mode=Q_PRIVATE, capacity=getattr(args, "q_capacity", 0),
            lr=args.lr, discount=args.discount, epsilon=args.epsilon,
            epsilon_anneal=args.epsilon_anneal)
        self._k = 0                # number of update_epsilon calls so far
        self.epsilon = 0.0         # value.py:28
        self.Q = _QView(self)
        dev = env.batched.device
        hw = env.batched.hw
        self._s = torch.zeros(1, hw, dtype=torch.uint8, device=dev)
        self._s2 = torch.zeros(1, hw, dtype=torch.uint8, device=dev)
        self._a = torch.zeros(1, dtype=torch.uint8, device=dev)
        self._r = torch.zeros(1, dtype=torch.float64, device=dev)

    def _upload(self, buf, state):
        flat = np.asarray(state).reshape(-1)
        board = flat.astype(np.uint8)
        buf.copy_(torch.from_numpy(board).reshape(1, -1))
        return flat

    def act(self, state):
        self._upload(self._s, state)
        action = self.table.act(self._s, self._k, explore=False)
        return np.int64(action[0].item())

    def act_explore(self, state):
        if self._rng == "numpy":
            # the reference's own two draws (value.py:38-39), from its stream
            if np.random.sample() < self.epsilon:
                return np.random.choice(self.action_n)
            return self.act(state)
        self._upload(self._s, state)
        action = self.table.act(self._s, self._k, explore=True)
        return np.int64(action[0].item())

    def learn(self, state, action, reward, successor):
        self._upload(self._s, state)
        self._upload(self._s2, successor)
        self._a[0] = _as_int_action(action)
        self._r[0] = float(reward)
        self.table.learn(self._s, self._a, self._r, self._s2)

    def update_epsilon(self):
        self._k += 1
        self.epsilon = self.table.epsilon_at(self._k)
        return self.epsilon

    # -- the fused path: one launch per episode (gridfast.loops.tabq_learn_fused)
    def run_episode(self, cheat=False):
        """One whole episode of the tabq_learn body (common/learn.py:61-85) in
        ONE kernel launch, from the environment's current (freshly reset)
        state: act_explore, env.step, the --cheat swaps, learn, update_epsilon
        per step until done.  In numpy mode the kernel draws from the words
        the global MT19937 stream would hand the reference's calls, and the
        stream is then advanced by exactly what was consumed.  Returns
        (observation, reward, done, info, steps) of the episode's last step."""
        env = self.env
        limit = env.batched.max_iterations
        lent = env._lend_words(limit * _WORDS_PER_FUSED_STEP + 64, always=True)
        if lent is None:
            env.batched.t = self._k         # philox: streams keyed by the agent-step index
        steps, reward, hidden = self.table.rollout_episodes(1, limit, cheat=cheat, t0=self._k)
        env._settle_words(lent)
        host = torch.stack([steps.double()[0], reward[0], hidden[0]]).cpu().numpy()
        steps, reward, hidden = int(host[0]), float(host[1]), float(host[2])
        self._k += steps
        self.epsilon = self.table.epsilon_at(self._k)
        obs = env._after_fused(steps, done=True)
        hidden = None if hidden != hidden else hidden
        action = int(env.batched.actual_actions()[0].item())
        info = env._info(reward, hidden, True, action)
        return obs, reward, True, info, steps

    def evaluate_episodes(self, eval_timesteps):
        """default_eval's loop (common/eval.py:8-56) in one launch, from the
        environment's freshly reset state: greedy act (inserting unseen boards
        like the defaultdict) until the first episode end at or after
        `eval_timesteps` steps.  Returns the (return, performance) pairs of
        the evaluation episodes, in order."""
        env = self.env
        limit = env.batched.max_iterations
        lent = env._lend_words((eval_timesteps + limit) * (_WORDS_PER_FUSED_STEP - 3) + 64) if env._stochastic else None
        if lent is None:
            # philox: evaluation draws from a counter region training never reaches
            self._evals = getattr(self, "_evals", 0) + 1
            env.batched.t = (1 << 41) + self._evals * (eval_timesteps + limit)
        rows = self.table.evaluate_logged(env.batched, eval_timesteps, insert_on_miss=True)
        env._settle_words(lent)
        env._episode_return = float(rows[-1, 0])
        env._last_performance = float(rows[-1, 1])
        env._board = None
        return rows


class _CView:
    """Read-only dict-like view of the corruption estimates C, keyed like Q."""

    def __init__(self, agent):
        self._agent = agent

    def _snapshot(self):
        keys, _, c = self._agent.table.export(0, with_corruption=True)
        if len(keys) == 0:
            return {}
        env = self._agent.env.batched
        boards = env.keys_to_boards(torch.as_tensor(keys.view(np.int64)).to(env.device)).cpu().numpy()
        return {tuple(np.float32(v) for v in boards[i]): float(c[i]) for i in range(len(keys))}

    def __getitem__(self, key):
        return self._snapshot().get(tuple(np.float32(v) for v in key), self._agent.C_prior)

    def __len__(self):
        return len(self._snapshot())

    def items(self):
        return self._snapshot().items()


class GpuTabularSSQAgent(GpuTabularQAgent):
    """Drop-in for TabularSSQAgent (ssrl/agents.py:9-86): TabularQAgent plus a
    per-state corruption estimate C (prior args.C_prior) that scales the reward
    in learn, a query budget (args.budget), query_H / learn_C / reset_history.
    Constructor (env, args) like every AGENT_MAP class (parse.py:39-48).

    The reference's class cannot run as shipped (it reads raw pycolab
    timesteps, `state["board"]`, while train.py hands it gym observations, and
    it has no LEARN_MAP entry -- SURVEY.md section 2.1); this adapter takes the
    gym-style boards everything else in the reference uses.  The arithmetic is
    the reference's (oracle/tabular.py restates it; tests compare the two)."""

    def __init__(self, env, args):
        super().__init__(env, args)
        self.C_prior = args.C_prior
        self.table.enable_ssrl(args.C_prior, args.budget)
        self.C = _CView(self)
        self._history = []          # boards passed to act_explore this episode (agents.py:29-32)

    # counters live on the device (the fused loop updates them there)
    def _counters(self):
        return [int(t[0].item()) for t in self.table.ssrl_counters()]

    @property
    def budget(self):
        return self._counters()[0] - (1 if getattr(self, "_query_pending", False) else 0)

    @property
    def episodes(self):
        return self._counters()[1]

    @property
    def corrupt_episodes(self):
        return self._counters()[2]

    def act_explore(self, state):
        action = super().act_explore(state)
        self._history.append(np.asarray(state).reshape(-1).astype(np.uint8))
        return action

    def query_H(self, env):
        """agents.py:45-48: spend one unit of budget, get the true performance."""
        self._query_pending = True
        return (env._env if hasattr(env, "_env") else env).get_last_performance()

    def _learn_c(self, corrupt, increment_episode=True):
        boards = None
        if self._history:
            boards = torch.from_numpy(np.stack(self._history)).to(self.env.batched.device)
        query = 1 if getattr(self, "_query_pending", False) else 0
        self._query_pending = False
        batched.check(self.table.L.sgk_ssrl_learn_c(
            self.table.h, 0, batched._p(boards), len(self._history), int(bool(corrupt)), query,
            int(bool(increment_episode)), batched._stream()))
        self._history = []

    def learn_C(self, corrupt_episode):
        """agents.py:50-75 followed by reset_history(corrupt_episode)."""
        self._learn_c(corrupt_episode)

    def reset_history(self, corrupt, increment_episode=True):
        """agents.py:77-82."""
        self._history = []
        self._learn_c(corrupt, increment_episode)


def register_with_reference(env_map=None, agent_map=None, gym_module=None, learn_map=None, eval_map=None,
                            warmup_map=None, **env_kwargs):
    """Register the GPU path into the reference's own registries -- AGENT_MAP
    (safe_grid_agents/parsing/parse.py:39-48), LEARN_MAP (common/learn.py:107-113),
    EVAL_MAP (common/eval.py:59), WARMUP_MAP (common/warmup.py:31-33) -- and, if
    given, replace `gym.make` so that train.train(args) builds gridfast
    environments for the in-scope ids.  With the loop registries passed, one
    reference call of learn_fn / eval_fn / warmup_fn becomes one fused launch
    (gridfast.loops); without them the reference's own per-step loops drive
    the adapters call by call.  "tabular-ssq" gains the LEARN_MAP entry the
    reference lacks.  Returns the previous gym.make (or None)."""
    previous = None
    if agent_map is not None:
        agent_map["tabular-q"] = GpuTabularQAgent
        agent_map["deep-q"] = GpuDeepQAgent
        agent_map["tabular-ssq"] = GpuTabularSSQAgent
    if learn_map is not None or eval_map is not None or warmup_map is not None:
        from . import loops
        if learn_map is not None:
            learn_map["tabular-q"] = loops.tabq_learn_fused
            learn_map["tabular-ssq"] = loops.ssq_learn_fused
            learn_map["deep-q"] = loops.dqn_learn_fused
        if eval_map is not None:
            eval_map["tabular-q"] = loops.default_eval_fused
            eval_map["tabular-ssq"] = loops.default_eval_fused
        if warmup_map is not None:
            warmup_map["tabular-ssq"] = loops.random_warmup_fused
            warmup_map["deep-q"] = loops.dqn_warmup_fused
    if gym_module is not None:
        previous = getattr(gym_module, "make", None)

        def _make(env_id, *a, **k):
            if env_id in batched.KIND_BY_ID:
                return GridworldEnv(env_id, **env_kwargs)
            if previous is None:
                raise SgkError("environment %r is outside the GPU path's scope" % env_id)
            return previous(env_id, *a, **k)

        gym_module.make = _make
    return previous


class _ReplayView:
    """What dqn_warmup touches through `agent.replay` (common/warmup.py:21)."""

    def __init__(self, agent):
        self._agent = agent

    def add(self, state, action, reward, successor, terminal):
        self._agent._replay_add(state, action, reward, successor, terminal)

    def __len__(self):
        return int(self._agent.net.replay_count)


class GpuDeepQAgent:
    """Drop-in for DeepQAgent (common/agents/value.py:61-187) behind the
    reference's dqn_warmup / dqn_learn loops (common/warmup.py:8-23,
    common/learn.py:29-58): constructor (env, args) reading args.lr,
    .discount, .epsilon, .epsilon_anneal, .batch_size, .n_layers, .n_hidden,
    .replay_capacity; act, act_explore, learn(..., terminal, history),
    update_epsilon, sync_target_Q, replay.add.  An N == 1 view of
    BatchedDeepQ: the network, the replay ring and the optimiser live on the
    GPU.  The exploration decision is drawn on the host from a private numpy
    stream seeded with args.seed (the reference samples a torch Categorical on
    the device, value.py:94-111 -- same distribution, different stream:
    statistical parity only, SURVEY hard part H7)."""

    def __init__(self, env, args):
        if not isinstance(env, GridworldEnv):
            raise SgkError("GpuDeepQAgent needs a gridfast GridworldEnv")
        from .deepq import BatchedDeepQ
        self.env = env
        self.action_n = env.action_space.n
        self.discount = args.discount
        self.lr = args.lr
        self.batch_size = args.batch_size
        self._final_epsilon = args.epsilon
        self._anneal = args.epsilon_anneal
        self.net = BatchedDeepQ(
            env.batched, n_layers=args.n_layers, n_hidden=args.n_hidden,
            replay_capacity=args.replay_capacity, batch_size=args.batch_size, lr=args.lr,
            discount=args.discount, epsilon=args.epsilon, epsilon_anneal=args.epsilon_anneal,
            sync_every=getattr(args, "sync_every", 10000),
            reference_bxb_loss=getattr(args, "reference_bxb_loss", True), seed=getattr(args, "seed", 0) or 0)
        if getattr(args, "tensor_cores", None) is not None:      # default: tensor cores on (3xTF32) where supported
            self.net.set_tensor_cores(args.tensor_cores)
        self.replay = _ReplayView(self)
        self._k = 0
        self.epsilon = self._epsilon_at(0)      # value.py:76: the first pop, no 0.0 override
        dev = env.batched.device
        hw = env.batched.hw
        self._s = torch.zeros(1, hw, dtype=torch.uint8, device=dev)
        self._s2 = torch.zeros(1, hw, dtype=torch.uint8, device=dev)
        self._a = torch.zeros(1, dtype=torch.uint8, device=dev)
        self._r = torch.zeros(1, dtype=torch.float64, device=dev)
        self._term = torch.zeros(1, dtype=torch.uint8, device=dev)
        self._learn_steps = 0
        self._rs = np.random.RandomState(getattr(args, "seed", 0) or 0)

    def _epsilon_at(self, k):
        last = self._anneal - 1 if self._anneal > 1 else 0
        return 1.0 - (1 - self._final_epsilon) * min(k, last) / self._anneal

    def _upload(self, buf, state):
        buf.copy_(torch.from_numpy(np.asarray(state).reshape(1, -1).astype(np.uint8)))

    def act(self, state):
        """value.py:89-92: a 1-element tensor, like scores.argmax(1)."""
        self._upload(self._s, state)
        return self.net.q_values(self._s).argmax(1)

    def act_explore(self, state):
        if self._rs.random_sample() < self.epsilon:
            return int(self._rs.randint(0, self.action_n))
        return int(self.act(state).item())

    def _replay_add(self, state, action, reward, successor, terminal):
        self._upload(self._s, state)
        self._upload(self._s2, successor)
        self._a[0] = _as_int_action(action)
        self._r[0] = float(reward)
        self._term[0] = 1 if terminal else 0
        self.net.replay_add(self._s, self._a, self._r, self._s2, self._term)

    def learn(self, state, action, reward, successor, terminal, history):
        self._replay_add(state, action, reward, successor, terminal)
        scalars = self.net.learn(self._learn_steps)
        self._learn_steps += 1
        history["writer"].add_scalar("Train/value_loss", float(scalars[0].item()), history["t"])   # value.py:124
        return history

    def sync_target_Q(self):
        self.net.sync_target()

    # -- the fused path (gridfast.loops.dqn_learn_fused)
    def run_episode(self, cheat=False, t=0):
        """One episode of the dqn_learn body (common/learn.py:29-58) with every
        step's work (act forward, epsilon-greedy, env.step, replay.add, the
        optimiser step, the periodic target sync) enqueued device-side by ONE
        C-ABI call; the host reads back three numbers per step (done, reward,
        loss) where the reference crosses the host/device boundary >= 8 times
        (SURVEY.md 3.2).  Exploration draws come from the environment's Philox
        stream.  Returns (observation, reward, done, info, steps, losses)."""
        env = self.env
        net = self.net
        losses = []
        limit = env.batched.max_iterations
        reward, done = 0.0, False
        while not done and len(losses) < limit:
            pos = net.replay_position
            env.batched.t = self._k
            net.rollout(1, cheat=cheat)
            _, _, r, _, term = net.replay_rows(pos, 1)
            host = torch.cat([r, term.float(), net.last_scalars_device()[:1]]).cpu().numpy()
            reward, done = float(host[0]), bool(host[1])
            losses.append(float(host[2]))
            self._k += 1
            self._learn_steps += 1
        self.epsilon = self._epsilon_at(self._k)
        steps = len(losses)
        obs = env._after_fused(steps, done=done)
        info = {"hidden_reward": reward if cheat else None, "observed_reward": None if cheat else reward,
                "extra_observations": {}}
        return obs, reward, done, info, steps, losses

    def update_epsilon(self):
        self._k += 1
        self.epsilon = self._epsilon_at(self._k)
        return self.epsilon
