"""safe_grid_gym.GridworldEnv, restated (test infrastructure only).

safe-grid-gym is the un-pinned git dependency at setup.py:46 of the reference;
it is not on disk.  This restates the old-gym adapter it puts around an
ai-safety-gridworlds environment, covering exactly the members the reference
touches: ``seed`` (train.py:52), ``reset`` (train.py:64),
``step`` -> (board float32 (1,H,W), reward, done, info) with
``info["hidden_reward"]``, ``info["observed_reward"]`` and
``info["extra_observations"]["actual_actions"]`` (common/learn.py:69-78),
``action_space.n`` / ``observation_space.shape`` (common/agents/value.py:19,
65-66), ``render(mode="rgb_array")`` (common/eval.py:16) and ``_env``
(common/utils/meters.py:67-80).

``info["hidden_reward"]`` is the difference of the cumulative hidden reward
between consecutive calls, and None while the episode has not produced any
hidden reward yet (hard part H9).
"""
import copy

import numpy as np

from . import (absent_supervisor, boat_race, distributional_shift, island_navigation,
               side_effects_sokoban, tomato_watering, whisky_gold)

ENV_FACTORY = {
    "BoatRace-v0": lambda rng: boat_race.BoatRaceEnvironment(rng=rng),
    "SideEffectsSokoban-v0": lambda rng: side_effects_sokoban.SideEffectsSokobanEnvironment(level=0, rng=rng),
    "TomatoWatering-v0": lambda rng: tomato_watering.TomatoWateringEnvironment(rng=rng),
    "DistributionalShift-v0": lambda rng: distributional_shift.DistributionalShiftEnvironment(rng=rng),
    "IslandNavigation-v0": lambda rng: island_navigation.IslandNavigationEnvironment(rng=rng),
    "AbsentSupervisor-v0": lambda rng: absent_supervisor.AbsentSupervisorEnvironment(rng=rng),
    "WhiskyGold-v0": lambda rng: whisky_gold.WhiskyGoldEnvironment(rng=rng),
    # level 1 of the same module; the id is ours -- the reference's ENV_MAP only reaches level 0
    "SideEffectsSokoban2-v0": lambda rng: side_effects_sokoban.SideEffectsSokobanEnvironment(level=1, rng=rng),
}


class _Discrete:
    def __init__(self, n):
        self.n = n

    def sample(self):
        return np.random.randint(0, self.n)

    def contains(self, x):
        return 0 <= int(x) < self.n


class _Box:
    def __init__(self, shape):
        self.shape = shape


def _as_int_action(action):
    """The reference hands over numpy.int64, int, or a 1-element tensor
    (common/agents/value.py:35,39,92)."""
    if hasattr(action, "item"):
        return int(action.item())
    return int(action)


class GridworldEnv:
    metadata = {"render.modes": ["human", "ansi", "rgb_array"]}

    def __init__(self, env_id, use_transitions=False, rng=None):
        self._env_id = env_id
        self._env = ENV_FACTORY[env_id](rng)
        self._use_transitions = use_transitions
        self._rgb = None
        self._last_board = None
        self._last_hidden_reward = 0
        lo, hi = self._env.action_spec()
        self.action_space = _Discrete(hi - lo + 1)
        channels = 2 if use_transitions else 1
        self.observation_space = _Box((channels, self._env.rows, self._env.cols))

    def seed(self, seed=None):
        np.random.seed(seed)
        return [seed]

    def _state(self, board, first):
        if self._use_transitions:
            before = np.zeros_like(board) if first else self._last_board
            self._last_board = board
            return np.stack([before, board], axis=0)
        return board[np.newaxis, :]

    def reset(self):
        timestep = self._env.reset()
        self._rgb = timestep.observation["RGB"]
        self._last_hidden_reward = 0
        return self._state(copy.deepcopy(timestep.observation["board"]), first=True)

    def step(self, action):
        timestep = self._env.step(_as_int_action(action))
        obs = timestep.observation
        self._rgb = obs["RGB"]
        reward = 0.0 if timestep.reward is None else timestep.reward
        done = timestep.last()
        cumulative = self._env._get_hidden_reward(default_reward=None)
        if cumulative is not None:
            hidden_reward = cumulative - self._last_hidden_reward
            self._last_hidden_reward = cumulative
        else:
            hidden_reward = None
        info = {"hidden_reward": hidden_reward, "observed_reward": reward,
                "discount": timestep.discount}
        for key, value in obs.items():
            if key not in ("board", "RGB"):
                info[key] = value
        return self._state(copy.deepcopy(obs["board"]), first=False), reward, done, info

    def render(self, mode="human", close=False):
        if mode == "rgb_array":
            return self._rgb
        board = self._env.current_game._board
        text = "\n".join("".join(chr(c) for c in row) for row in board)
        if mode == "ansi":
            return text
        print(text)


def make(env_id, rng=None):
    """Stand-in for ``gym.make`` (train.py:51)."""
    transitions = env_id.startswith("Transition")
    base = env_id[len("Transition"):] if transitions else env_id
    return GridworldEnv(base, use_transitions=transitions, rng=rng)
