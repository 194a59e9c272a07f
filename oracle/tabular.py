"""Tabular agents and their interaction loops, restated (test infrastructure).

Follows safe_grid_agents/common/agents/value.py:15-58 (TabularQAgent),
safe_grid_agents/ssrl/agents.py:9-86 (TabularSSQAgent),
safe_grid_agents/common/agents/dummy.py:7-16 (RandomAgent) and the loops in
safe_grid_agents/common/learn.py:8-85 (`whiler`, `tabq_learn`).  Unlike the
environment modules this half IS pinned: tests/golden/make_golden.py drives
the live reference classes and this restatement with the same numpy stream and
records that boards, actions, rewards and every Q row agree bit for bit.

Differences from the reference, all about plumbing, none about arithmetic:
  * random draws go through an injected stream (rng.py) instead of the global
    numpy functions -- with NumpyGlobalRng they are the same calls;
  * the epsilon schedule is evaluated in closed form instead of popping a
    100k-element list each step (value.py:23-28,54-58) -- same float64 values;
  * the SSRL agent takes boards (the reference's takes raw pycolab timesteps
    and cannot run as shipped) and its online loop is
    the one SURVEY.md section 8a row S defines.
"""
import numpy as np

from . import rng as rng_mod


def board_key(board):
    """value.py:34 -- the exact, collision-free dict key."""
    return tuple(np.asarray(board).flatten())


class EpsilonSchedule:
    """value.py:23-28,54-58.  `current` after k calls to `advance()`.

    The constructor's own `update_epsilon()` pops entry 0, then epsilon is
    overwritten with 0.0 (tabular agent only, value.py:28); every later call
    pops entry k (k = 1, 2, ...) until the list is empty, after which epsilon
    stays at the last entry, 1 - (1-eps)*(anneal-1)/anneal.
    """

    def __init__(self, epsilon, anneal, zero_first=True):
        self.final = epsilon
        self.anneal = anneal
        self.popped = 1 if anneal > 0 else 0
        self.current = 0.0 if zero_first else self.value_at(0)

    def value_at(self, t):
        return 1.0 - (1 - self.final) * t / self.anneal

    def advance(self):
        if self.popped < self.anneal:
            self.current = self.value_at(self.popped)
            self.popped += 1
        return self.current


class TabularQAgent:
    def __init__(self, n_actions, discount, epsilon, epsilon_anneal, lr, rng=None):
        self.action_n = n_actions
        self.discount = discount
        self.lr = lr
        self.schedule = EpsilonSchedule(epsilon, epsilon_anneal)
        self.Q = {}
        self.rng = rng or rng_mod.NumpyGlobalRng()

    @property
    def epsilon(self):
        return self.schedule.current

    def row(self, key):
        r = self.Q.get(key)
        if r is None:
            r = self.Q[key] = np.zeros(self.action_n)
        return r

    def act(self, state):
        return int(np.argmax(self.row(board_key(state))))

    def act_explore(self, state):
        if self.rng.agent_uniform() < self.epsilon:
            return int(self.rng.agent_choice(self.action_n))
        return self.act(state)

    def learn(self, state, action, reward, successor):
        s, s2 = board_key(state), board_key(successor)
        nxt = self.row(s2)
        target = reward + self.discount * nxt[int(np.argmax(nxt))]
        cur = self.row(s)
        cur[action] += self.lr * (target - cur[action])

    def update_epsilon(self):
        return self.schedule.advance()


class TabularSSQAgent(TabularQAgent):
    """ssrl/agents.py:9-86 on boards.  C defaults to C_prior per state."""

    def __init__(self, n_actions, discount, epsilon, epsilon_anneal, lr,
                 budget, C_prior, rng=None):
        super().__init__(n_actions, discount, epsilon, epsilon_anneal, lr, rng)
        self.budget = budget
        self.C_prior = C_prior
        self.C = {}
        self.C_support = {}
        self.corrupt_episodes = 0
        self.episodes = 0
        self._history = []

    def corruption(self, key):
        if key not in self.C:
            self.C[key] = self.C_prior
        return self.C[key]

    def act_explore(self, state):
        action = super().act_explore(state)
        self._history.append(board_key(state))
        return action

    def learn(self, state, action, reward, successor):
        s, s2 = board_key(state), board_key(successor)
        reward_estimate = reward * (1 - self.corruption(s))
        nxt = self.row(s2)
        target = reward_estimate + self.discount * nxt[int(np.argmax(nxt))]
        cur = self.row(s)
        cur[action] += self.lr * (target - cur[action])

    def query_H(self, env):
        self.budget -= 1
        return env.get_last_performance()

    def learn_C(self, corrupt_episode):
        for key in self._history:
            if not corrupt_episode:
                self.C[key] = self.corruption(key) * 0
                self.C_support.pop(key, None)
            else:
                self.C_support[key] = self.C_support.get(key, 0) + 1
                self.C[key] = self.corruption(key) * (
                    self.episodes / (self.corrupt_episodes + 1))
        self.reset_history(corrupt_episode)

    def reset_history(self, corrupt, increment_episode=True):
        if corrupt:
            self.corrupt_episodes += 1
        if increment_episode:
            self.episodes += 1
        self._history = []


class RandomAgent:
    def __init__(self, n_actions, rng=None):
        self.action_n = n_actions
        self.rng = rng or rng_mod.NumpyGlobalRng()

    def act(self, state):
        return int(self.rng.random_action(self.action_n))


def random_warmup(agent, env, n_episodes, rando=None, t0=1 << 40, env_id=0):
    """ssrl/warmup.py:4-35 on the gym-style environment (the reference's drives
    the raw pycolab API and cannot run as shipped):
    `n_episodes` = int(args.budget * args.warmup) random-policy episodes, each
    begun by env.reset(); after each, safety = agent.query_H(env), corrupt =
    episode_return - safety > 0, agent.learn_C(corrupt).  The warm-up never
    calls agent.act_explore, so agent._history is empty and learn_C only moves
    the episode counters -- exactly as in the reference.  Returns steps taken."""
    rando = rando or RandomAgent(agent.action_n, rng=agent.rng)
    stream = agent.rng
    t = t0
    for _ in range(n_episodes):
        stream.set_context(env_id, t)
        env.reset()
        done = False
        while not done:
            stream.set_context(env_id, t)
            _, _, done, _ = env.step(rando.act(None))
            t += 1
        safety = agent.query_H(env._env)
        agent.learn_C(env._env.episode_return - safety > 0)
    return t - t0


def run_tabq(agent, env, n_steps, cheat=False, t0=0, env_id=0, record=None,
             ssrl=False):
    """`n_steps` iterations of the tabq_learn body (learn.py:61-85) with the
    episode loop of train.py:62-70 around it (reset when done).  The per-step
    tensorboard scalar is dropped.  With `ssrl`, at every episode end the
    agent queries H while budget remains and updates C.
    Returns per-episode (return, performance) pairs."""
    episodes = []
    stream = agent.rng
    t = t0
    stream.set_context(env_id, t)
    state = env.reset()
    for _ in range(n_steps):
        stream.set_context(env_id, t)
        action = agent.act_explore(state)
        successor, reward, done, info = env.step(action)
        observed = reward
        if cheat:
            reward = info["hidden_reward"]
            reward = 0.0 if reward is None else reward
            try:
                action = info["extra_observations"]["actual_actions"]
            except KeyError:
                pass
        agent.learn(state, action, reward, successor)
        agent.update_epsilon()
        if record is not None:
            record(t, state, int(action), observed, info["hidden_reward"], done, successor)
        t += 1
        state = successor
        if done:
            ret = env._env.episode_return
            perf = env._env.get_last_performance()
            episodes.append((ret, perf))
            if ssrl:
                if agent.budget > 0:
                    safety = agent.query_H(env._env)
                    agent.learn_C(ret - safety > 0)
                else:
                    agent.reset_history(False)
            stream.set_context(env_id, t)
            state = env.reset()
    return episodes
