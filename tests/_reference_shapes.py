"""The CALLER side of the hot path, restated for the GPU box (test infrastructure).

`/root/reference` does not exist where the -m gpu tests run, so the pieces of
the reference that CALL the path -- train.train (train.py:21-81), the meters
(common/utils/meters.py:9-63) and the registries' shapes (parsing/parse.py:22-48,
common/learn.py:107-113, common/eval.py:59, common/warmup.py:31-33) -- are
restated here, line for line in control flow, so that gridfast's registry
functions are driven exactly the way the reference drives them.  Where
/root/reference exists the tests use the real train.train instead.
"""
import bisect
import random
from collections import defaultdict

import numpy as np


class AverageMeter:
    """common/utils/meters.py:9-49"""

    def __init__(self, include_history=False):
        self.include_history = include_history
        self.reset(reset_history=True)

    def reset(self, reset_history=False):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self._max = -float("inf")
        self.count = 0
        if reset_history:
            self._history = None
            if self.include_history:
                self._history = []

    def update(self, val, n=1):
        self.val = val
        self._max = max(self._max, self.val)
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count
        if self._history is not None:
            for _ in range(n):
                bisect.insort(self._history, val)

    @property
    def max(self):
        return self._max


def make_meters(history):
    """common/utils/meters.py:52-63"""
    try:
        returns = history["returns"]
    except KeyError:
        returns = AverageMeter(include_history=True)
    return {"returns": returns, "safeties": AverageMeter(), "margins": AverageMeter(),
            "margins_support": AverageMeter()}


ENV_MAP = {"boat": "BoatRace-v0", "island": "IslandNavigation-v0", "lava": "DistributionalShift-v0",
           "sokoban": "SideEffectsSokoban-v0", "super": "AbsentSupervisor-v0", "tomato": "TomatoWatering-v0",
           "whisky": "WhiskyGold-v0"}      # parse.py:22-37, the in-scope ids


def noop_warmup(agent, env, history, args):
    return agent, env, history, args


def empty_registries():
    """(AGENT_MAP, LEARN_MAP, EVAL_MAP, WARMUP_MAP) shaped like the reference's, empty."""
    return {}, {}, {}, defaultdict(lambda: noop_warmup, {})


class RecordingWriter:
    def __init__(self, log_dir=None):
        self.events = []

    def add_scalar(self, tag, value, step=None):
        self.events.append(["scalar", tag, float(value), int(step)])

    def add_scalars(self, tag, values, step=None):
        self.events.append(["scalars", tag, {k: float(v) for k, v in values.items()}, int(step)])

    def add_text(self, *a, **k):
        pass

    add_video = add_histogram = add_text


def train(args, make_env, agent_map, learn_map, eval_map, warmup_map, writer):
    """train.py:21-81 with the registries and gym.make passed in.  Returns the agent and env."""
    import torch

    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    env_name = ENV_MAP[args.env_alias]
    agent_class = agent_map[args.agent_alias]
    warmup_fn = warmup_map[args.agent_alias]
    learn_fn = learn_map[args.agent_alias]
    eval_fn = eval_map[args.agent_alias]
    history, eval_history = make_meters({}), make_meters({})
    history["writer"] = writer
    eval_history["writer"] = writer
    env = make_env(env_name)
    env.seed(args.seed)
    agent = agent_class(env, args)
    agent, env, history, args = warmup_fn(agent, env, history, args)
    history["t"], history["t_learn"] = 0, 0
    history["episode"], eval_history["period"] = 0, 0
    for episode in range(args.episodes):
        env_state = (env.reset(), 0.0, False, {"hidden_reward": 0.0, "observed_reward": 0.0})
        history["episode"] += 1
        env_state, history, eval_next = learn_fn(agent, env, env_state, history, args)
        if eval_next:
            eval_history = eval_fn(agent, env, eval_history, args)
            eval_next = False
    eval_history = eval_fn(agent, env, eval_history, args)
    return agent, env, history
