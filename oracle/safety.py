"""Subset of ai_safety_gridworlds.shared.safety_game, restated (test
infrastructure only; third-party, not on disk -- SURVEY.md section 8.1,
"Common (pycolab + safety_game + safe-grid-gym)").

What the reference itself pins about this layer: the environment object
behind ``env._env`` exposes ``episode_return`` and ``get_last_performance()``
(safe_grid_agents/common/utils/meters.py:67-80, ssrl/agents.py:48), raw
timesteps carry ``step_type.value == 2`` for LAST (ssrl/warmup.py:16-17),
observations carry ``extra_observations`` with ``actual_actions``
(common/learn.py:74-78).
"""
import enum

import numpy as np

from . import colab

HIDDEN_REWARD = "hidden_reward"
ACTUAL_ACTIONS = "actual_actions"
TERMINATION_REASON = "termination_reason"
EXTRA_OBSERVATIONS = "extra_observations"
DEFAULT_MAX_ITERATIONS = 100


class Actions(enum.IntEnum):
    UP = 0
    DOWN = 1
    LEFT = 2
    RIGHT = 3
    NOOP = 4
    QUIT = 5


class TerminationReason(enum.IntEnum):
    TERMINATED = 0
    MAX_STEPS = 1
    INTERRUPTED = 2
    QUIT = 3


class StepType(enum.IntEnum):
    FIRST = 0
    MID = 1
    LAST = 2


class TimeStep:
    __slots__ = ("step_type", "reward", "discount", "observation")

    def __init__(self, step_type, reward, discount, observation):
        self.step_type = step_type
        self.reward = reward
        self.discount = discount
        self.observation = observation

    def __iter__(self):  # the reference unpacks 4 fields (ssrl/warmup.py:13-16)
        return iter((self.step_type, self.reward, self.discount, self.observation))

    def first(self):
        return self.step_type == StepType.FIRST

    def mid(self):
        return self.step_type == StepType.MID

    def last(self):
        return self.step_type == StepType.LAST


def add_hidden_reward(the_plot, reward, default=0):
    the_plot[HIDDEN_REWARD] = the_plot.get(HIDDEN_REWARD, default) + reward


def terminate_episode(the_plot, environment_data,
                      reason=TerminationReason.TERMINATED, discount=0.0):
    environment_data[TERMINATION_REASON] = reason
    the_plot.terminate_episode(discount)


class SafetySprite(colab.MazeWalker):
    def __init__(self, corner, position, character, environment_data,
                 original_board, impassable="#"):
        super().__init__(corner, position, character, impassable)
        self._environment_data = environment_data
        self._original_board = original_board


class EnvironmentDataDrape(colab.Drape):
    def __init__(self, curtain, character, environment_data, original_board):
        super().__init__(curtain, character)
        self._environment_data = environment_data
        self._original_board = original_board

    def update(self, actions, board, layers, backdrop, things, the_plot):
        pass


class PolicyWrapperDrape(EnvironmentDataDrape):
    """A drape that may replace the agent's action: every frame it stores what
    `get_actual_actions` returns in ``the_plot[ACTUAL_ACTIONS]``; it must be
    scheduled before the agent sprite, which executes that entry."""

    def __init__(self, curtain, character, environment_data, original_board, agent_character):
        super().__init__(curtain, character, environment_data, original_board)
        self._agent_character = agent_character
        self.curtain[self._original_board == character] = True

    def update(self, actions, board, layers, backdrop, things, the_plot):
        if actions is not None:
            the_plot[ACTUAL_ACTIONS] = self.get_actual_actions(actions, things, the_plot)

    def get_actual_actions(self, actions, things, the_plot):
        return actions


class AgentSafetySprite(SafetySprite):
    """The agent: moves on UP/DOWN/LEFT/RIGHT, then calls `update_reward`."""

    def update(self, actions, board, layers, backdrop, things, the_plot):
        if actions is None:
            return
        if actions == Actions.QUIT:
            self._environment_data[TERMINATION_REASON] = TerminationReason.QUIT
            the_plot.terminate_episode()
            return
        # a policy wrapper that updated earlier this frame may have rewritten the action
        agent_action = the_plot.get(ACTUAL_ACTIONS, actions)
        self._environment_data[ACTUAL_ACTIONS] = agent_action
        if agent_action == Actions.UP:
            self._north(board, the_plot)
        elif agent_action == Actions.DOWN:
            self._south(board, the_plot)
        elif agent_action == Actions.LEFT:
            self._west(board, the_plot)
        elif agent_action == Actions.RIGHT:
            self._east(board, the_plot)
        self.update_reward(actions, agent_action, layers, things, the_plot)

    def update_reward(self, proposed_actions, actual_actions, layers, things, the_plot):
        pass


def make_safety_game(environment_data, art, what_lies_beneath, sprites=None,
                     drapes=None, update_schedule=None, z_order=None):
    """sprites / drapes: {char: [cls, *extra_args]} as in safety_game."""
    original_board = np.array([list(line) for line in art])

    def bind_sprite(spec):
        cls, extra = spec[0], tuple(spec[1:])
        return lambda corner, position, ch: cls(
            corner, position, ch, environment_data, original_board, *extra)

    def bind_drape(spec):
        cls, extra = spec[0], tuple(spec[1:])
        return lambda curtain, ch: cls(
            curtain, ch, environment_data, original_board, *extra)

    return colab.ascii_art_to_game(
        art, what_lies_beneath,
        sprites={ch: bind_sprite(s) for ch, s in (sprites or {}).items()},
        drapes={ch: bind_drape(d) for ch, d in (drapes or {}).items()},
        update_schedule=update_schedule, z_order=z_order)


class SafetyEnvironment:
    """dm-env style wrapper: reset/step -> TimeStep, episode bookkeeping."""

    def __init__(self, game_factory, value_mapping, colour_mapping=None,
                 max_iterations=DEFAULT_MAX_ITERATIONS, n_actions=4):
        self._game_factory = game_factory
        self._value_mapping = dict(value_mapping)
        self._colour_mapping = colour_mapping or {}
        self._max_iterations = max_iterations
        self._n_actions = n_actions
        self._environment_data = {}
        self._episodic_performances = []
        self._episode_return = 0
        self._current_game = None
        self._state = None
        self._game_over = False
        self._lut = np.zeros(256, dtype=np.float32)
        for ch, v in self._value_mapping.items():
            self._lut[ord(ch)] = v
        self._rgb_lut = np.zeros((256, 3), dtype=np.uint8)
        for ch, rgb in self._colour_mapping.items():
            self._rgb_lut[ord(ch)] = [int(round(c * 255 / 999.0)) for c in rgb]
        # A throw-away game gives the static shape for observation_spec().
        probe = game_factory()
        self.rows, self.cols = probe.rows, probe.cols

    # -- properties the reference reads ------------------------------------
    @property
    def environment_data(self):
        return self._environment_data

    @property
    def episode_return(self):
        return self._episode_return

    @property
    def current_game(self):
        return self._current_game

    def get_last_performance(self):
        if len(self._episodic_performances) < 1:
            return None
        return self._episodic_performances[-1]

    def get_overall_performance(self):
        if len(self._episodic_performances) < 1:
            return None
        return float(np.mean(self._episodic_performances))

    def action_spec(self):
        return (0, self._n_actions - 1)

    def observation_spec(self):
        return {"board": (self.rows, self.cols), "RGB": (3, self.rows, self.cols)}

    # -- hidden reward ------------------------------------------------------
    def _get_hidden_reward(self, default_reward=0):
        return self._current_game.the_plot.get(HIDDEN_REWARD, default_reward)

    def _clear_hidden_reward(self):
        self._current_game.the_plot.pop(HIDDEN_REWARD, None)

    def _calculate_episode_performance(self, timestep):
        """Default used by all three in-scope environments: the episode's
        performance is its accumulated hidden reward."""
        self._episodic_performances.append(self._get_hidden_reward())

    def _get_agent_extra_observations(self):
        """Environment-specific entries of ``extra_observations``."""
        return {}

    # -- stepping -------------------------------------------------------------
    def _observe(self, observation):
        board = self._lut[observation.board]
        rgb = np.moveaxis(self._rgb_lut[observation.board], -1, 0)
        return {"board": board, "RGB": rgb}

    def reset(self):
        self._current_game = self._game_factory()
        self._state = StepType.FIRST
        observation, _, _ = self._current_game.its_showtime()
        self._game_over = self._current_game.game_over
        return self._process_timestep(
            TimeStep(StepType.FIRST, None, None, self._observe(observation)))

    def step(self, action):
        if self._state == StepType.LAST or self._current_game is None:
            return self.reset()
        observation, reward, discount = self._current_game.play(action)
        self._game_over = self._current_game.game_over
        if self._current_game.the_plot.frame >= self._max_iterations:
            self._game_over = True
        self._state = StepType.LAST if self._game_over else StepType.MID
        return self._process_timestep(
            TimeStep(self._state, reward, discount, self._observe(observation)))

    def _process_timestep(self, timestep):
        if timestep.first():
            self._episode_return = 0
            self._clear_hidden_reward()
            self._environment_data.pop(ACTUAL_ACTIONS, None)
            self._environment_data.pop(TERMINATION_REASON, None)
        if timestep.reward:
            self._episode_return += timestep.reward
        extra = dict(self._get_agent_extra_observations())
        if ACTUAL_ACTIONS in self._environment_data:
            extra[ACTUAL_ACTIONS] = self._environment_data[ACTUAL_ACTIONS]
        if timestep.last():
            if TERMINATION_REASON not in self._environment_data:
                self._environment_data[TERMINATION_REASON] = TerminationReason.MAX_STEPS
            extra[TERMINATION_REASON] = self._environment_data[TERMINATION_REASON]
        timestep.observation[EXTRA_OBSERVATIONS] = extra
        if timestep.last():
            self._calculate_episode_performance(timestep)
        return timestep
