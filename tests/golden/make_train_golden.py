"""Generate tests/golden/train_*.npz by running the reference's REAL train.train.

Run in the build container only (it imports /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_train_golden.py

`/root/reference/train.py::train(args)` (train.py:21-81) is executed
unmodified, with the reference's own TabularQAgent, tabq_learn / whiler,
default_eval and track_metrics, against the oracle's GridworldEnv (the real
environment stack is absent, profiles/r02_env_probe.log).  Only the three
third-party modules train.py imports are stubbed:

    gym.make            -> oracle.gridworld_env.make
    safe_grid_gym       -> empty module (imported for its registration side effect)
    tensorboardX.SummaryWriter -> a recorder of every add_scalar / add_scalars call

The fixture holds the complete ordered scalar log (Train/epsilon per step,
Train/returns|safeties|margins|margins_support per episode, Evaluation/* per
evaluation period), the agent's final Q table and the position of numpy's
global stream at the end.  tests/test_gpu_dropin.py replays the same run
through gridfast's fused LEARN_MAP / EVAL_MAP functions and must reproduce
all of it bit for bit -- and runs the real train.train itself whenever
/root/reference and a GPU are present together.
"""
import argparse
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


class RecordingWriter:
    """tensorboardX.SummaryWriter stand-in: keeps what the reference logs."""

    events = None

    def __init__(self, log_dir=None):
        self.log_dir = log_dir

    def add_scalar(self, tag, value, step=None):
        RecordingWriter.events.append(["scalar", tag, float(value), int(step)])

    def add_scalars(self, tag, values, step=None):
        RecordingWriter.events.append(["scalars", tag, {k: float(v) for k, v in values.items()}, int(step)])

    def add_text(self, *a, **k):
        pass

    add_video = add_histogram = add_text


def install_stubs():
    from oracle import gridworld_env

    crmdp = types.ModuleType("ai_safety_gridworlds.environments.tomato_crmdp")
    crmdp.REWARD_FACTOR = 0.02
    sys.modules["ai_safety_gridworlds"] = types.ModuleType("ai_safety_gridworlds")
    sys.modules["ai_safety_gridworlds.environments"] = types.ModuleType("ai_safety_gridworlds.environments")
    sys.modules["ai_safety_gridworlds.environments.tomato_crmdp"] = crmdp
    gym = types.ModuleType("gym")
    gym.make = gridworld_env.make
    sys.modules["gym"] = gym
    sys.modules["safe_grid_gym"] = types.ModuleType("safe_grid_gym")
    tbx = types.ModuleType("tensorboardX")
    tbx.SummaryWriter = RecordingWriter
    sys.modules["tensorboardX"] = tbx
    sys.path.insert(0, REF)
    os.chdir(REF)       # parsing/__init__.py opens its YAML by cwd-relative path


CASES = [
    # name, env alias, seed, episodes, eval_every, eval_timesteps, lr, epsilon_anneal, cheat
    ("train_boat", "boat", 3, 24, 10, 250, 0.5, 1500, False),
    ("train_sokoban", "sokoban", 5, 45, 15, 120, 0.5, 700, False),
    ("train_tomato", "tomato", 7, 11, 5, 220, 0.5, 500, False),
    ("train_island_cheat", "island", 4, 30, 12, 150, 0.25, 600, True),
    ("train_super", "super", 13, 30, 10, 130, 0.5, 600, False),
]


def make_args(alias, seed, episodes, eval_every, eval_timesteps, lr, anneal, cheat):
    return argparse.Namespace(
        env_alias=alias, agent_alias="tabular-q", seed=seed, log_dir=None, episodes=episodes,
        eval_every=eval_every, eval_timesteps=eval_timesteps, eval_visualize_episodes=0,
        discount=0.99, cheat=cheat, lr=lr, epsilon=0.01, epsilon_anneal=anneal, disable_cuda=True, device="cpu")


def run_case(name, alias, seed, episodes, eval_every, eval_timesteps, lr, anneal, cheat):
    import train as ref_train                       # /root/reference/train.py, unmodified
    from safe_grid_agents.parsing import AGENT_MAP
    from safe_grid_agents.common.agents.value import TabularQAgent

    captured = []

    class Capturing(TabularQAgent):
        def __init__(self, env, args):
            super().__init__(env, args)
            captured.append(self)

    AGENT_MAP["tabular-q"] = Capturing
    RecordingWriter.events = []
    args = make_args(alias, seed, episodes, eval_every, eval_timesteps, lr, anneal, cheat)
    ref_train.train(args)
    AGENT_MAP["tabular-q"] = TabularQAgent
    agent = captured[0]
    keys = sorted(agent.Q)
    hw = len(keys[0])
    # where the global stream stands: the next raw words it would hand out
    tail = np.random.randint(0, 2 ** 32, size=4, dtype=np.uint32)
    events = RecordingWriter.events
    n_eps = sum(1 for e in events if e[1] == "Train/epsilon")
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        env_alias=np.array(alias), seed=np.int64(seed), episodes=np.int64(episodes), eval_every=np.int64(eval_every),
        eval_timesteps=np.int64(eval_timesteps), lr=np.float64(lr), epsilon_anneal=np.int64(anneal),
        cheat=np.bool_(cheat), events=np.array(json.dumps(events)),
        q_keys=np.array(keys, dtype=np.uint8).reshape(len(keys), hw),
        q_rows=np.array([agent.Q[k] for k in keys], dtype=np.float64),
        final_epsilon=np.float64(agent.epsilon), stream_tail=tail)
    print("%-20s events=%6d (epsilon %5d) states=%5d  OK" % (name, len(events), n_eps, len(keys)))


def main():
    install_stubs()
    for case in CASES:
        run_case(*case)


if __name__ == "__main__":
    main()
