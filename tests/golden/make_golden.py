"""Generate tests/golden/*.npz by running the LIVE reference agent code.

Run in the build container only (it imports /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

For every in-scope environment it runs the reference's own
``TabularQAgent`` through the reference's own ``tabq_learn``/``whiler`` loop
(safe_grid_agents/common/agents/value.py:15-58, common/learn.py:8-85) and
``track_metrics`` (common/utils/meters.py:66-108) against the oracle's
GridworldEnv, all drawing from numpy's global MT19937 stream exactly as
``train.py:31-33,51-70`` sets it up.  It then replays the same raw MT words
through the oracle's restated agent (oracle/tabular.py, ReplayWordsRng) and
refuses to write a fixture unless boards, actions, rewards, hidden rewards,
done flags, episode metrics and every Q row agree bit for bit.

What this pins: the agent half of the path and the loop around it, against the
reference itself.  What it cannot pin: the environment rules (third-party,
absent -- "parity unpinned"); the env trajectories in the
fixtures are self-generated from the restated oracle and labelled as such.
"""
import argparse
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference():
    stub = types.ModuleType("ai_safety_gridworlds.environments.tomato_crmdp")
    stub.REWARD_FACTOR = 0.02
    sys.modules["ai_safety_gridworlds"] = types.ModuleType("ai_safety_gridworlds")
    sys.modules["ai_safety_gridworlds.environments"] = types.ModuleType(
        "ai_safety_gridworlds.environments")
    sys.modules["ai_safety_gridworlds.environments.tomato_crmdp"] = stub
    sys.path.insert(0, REF)
    os.chdir(REF)  # parsing/__init__.py opens its YAML by cwd-relative path
    from safe_grid_agents.common.agents.value import TabularQAgent
    from safe_grid_agents.common.learn import tabq_learn
    from safe_grid_agents.common.utils import make_meters
    return TabularQAgent, tabq_learn, make_meters


class NullWriter:
    def add_scalar(self, *a, **k):
        pass

    add_scalars = add_text = add_video = add_histogram = add_scalar


class RecordingEnv:
    """Forwards to the oracle env, logging every transition."""

    def __init__(self, env):
        self.env = env
        self._env = env._env
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.log = []
        self.resets = []

    def reset(self):
        board = self.env.reset()
        self.resets.append(board.copy())
        return board

    def step(self, action):
        out = self.env.step(action)
        board, reward, done, info = out
        self.log.append((int(action), board.copy(), reward, info["hidden_reward"], done))
        return out


CASES = [
    # name, env id, seed, episodes, lr, epsilon_anneal, cheat
    ("boat_tabq_seed3", "BoatRace-v0", 3, 12, 0.5, 300, False),
    ("boat_tabq_default_anneal", "BoatRace-v0", 11, 6, 0.5, 100000, False),
    ("sokoban_tabq_seed5", "SideEffectsSokoban-v0", 5, 40, 0.5, 600, False),
    ("sokoban_tabq_cheat", "SideEffectsSokoban-v0", 9, 30, 0.25, 400, True),
    ("tomato_tabq_seed7", "TomatoWatering-v0", 7, 8, 0.5, 300, False),
    ("tomato_tabq_cheat", "TomatoWatering-v0", 21, 6, 0.1, 200, True),
    # widened environments.  Whisky has no hidden reward, so the
    # reference's --cheat loop cannot run on it (reward = None); the swap to the
    # action really taken is pinned by the oracle-level tests instead.
    ("island_tabq_cheat", "IslandNavigation-v0", 4, 30, 0.5, 500, True),
    ("super_tabq_seed13", "AbsentSupervisor-v0", 13, 30, 0.5, 500, False),
    ("super_tabq_cheat", "AbsentSupervisor-v0", 29, 30, 0.25, 500, True),
    ("whisky_tabq_seed17", "WhiskyGold-v0", 17, 25, 0.5, 400, False),
]
N_WORDS = 1 << 16


def run_case(name, env_id, seed, episodes, lr, anneal, cheat, ref):
    TabularQAgent, tabq_learn, make_meters = ref
    from oracle import gridworld_env, rng, tabular

    args = argparse.Namespace(discount=0.99, epsilon=0.01, epsilon_anneal=anneal,
                              lr=lr, cheat=cheat, eval_every=10 ** 9, seed=seed)
    # ---- live reference agent + loop, global numpy stream (train.py:31-70) ----
    np.random.seed(seed)
    env = RecordingEnv(gridworld_env.make(env_id))
    env.env.seed(seed)
    agent = TabularQAgent(env, args)
    history = make_meters({})
    history["writer"] = NullWriter()
    history["t"], history["t_learn"], history["episode"] = 0, 0, 0
    ep_returns, ep_safeties = [], []
    for _ in range(episodes):
        env_state = (env.reset(), 0.0, False, {"hidden_reward": 0.0, "observed_reward": 0.0})
        history["episode"] += 1
        env_state, history, _ = tabq_learn(agent, env, env_state, history, args)
        ep_returns.append(history["returns"].val)
        ep_safeties.append(history["safeties"].val)
    n_steps = history["t"]
    ref_Q = {k: v.copy() for k, v in agent.Q.items()}
    ref_eps = agent.epsilon

    # ---- oracle restatement, replaying the same raw words ----
    words = rng.mt19937_words(seed, N_WORDS)
    stream = rng.ReplayWordsRng(words)
    o_env = gridworld_env.make(env_id, rng=stream)
    o_agent = tabular.TabularQAgent(4, args.discount, args.epsilon, anneal, lr, rng=stream)
    o_log = []
    o_eps = tabular.run_tabq(
        o_agent, o_env, n_steps, cheat=cheat,
        record=lambda t, s, a, r, h, d, s2: o_log.append((a, s2.copy(), r, h, d)))
    assert stream.cursor < N_WORDS

    # ---- bit-for-bit agreement, or no fixture ----
    assert len(o_log) == len(env.log) == n_steps
    for (a0, b0, r0, h0, d0), (a1, b1, r1, h1, d1) in zip(env.log, o_log):
        assert a0 == a1 and d0 == d1 and np.array_equal(b0, b1)
        assert r0 == r1 and (h0 == h1 or (h0 is None and h1 is None))
    assert len(o_eps) == episodes
    for (ret, perf), r_ref, s_ref in zip(o_eps, ep_returns, ep_safeties):
        assert ret == r_ref and perf == s_ref
    assert set(ref_Q) == set(o_agent.Q)
    for k in ref_Q:
        assert np.array_equal(ref_Q[k], o_agent.Q[k]), (k, ref_Q[k], o_agent.Q[k])
    assert ref_eps == o_agent.epsilon

    keys = sorted(ref_Q)
    hw = env.observation_space.shape[1] * env.observation_space.shape[2]
    out = dict(
        env_id=np.array(env_id), seed=np.int64(seed), lr=np.float64(lr),
        discount=np.float64(args.discount), epsilon=np.float64(args.epsilon),
        epsilon_anneal=np.int64(anneal), cheat=np.bool_(cheat),
        n_steps=np.int64(n_steps), words_used=np.int64(stream.cursor),
        actions=np.array([l[0] for l in env.log], dtype=np.uint8),
        boards=np.array([l[1].reshape(hw) for l in env.log], dtype=np.uint8),
        reset_boards=np.array([b.reshape(hw) for b in env.resets], dtype=np.uint8),
        rewards=np.array([l[2] for l in env.log], dtype=np.float64),
        hidden=np.array([np.nan if l[3] is None else l[3] for l in env.log], dtype=np.float64),
        done=np.array([l[4] for l in env.log], dtype=np.bool_),
        episode_returns=np.array(ep_returns, dtype=np.float64),
        episode_performance=np.array(ep_safeties, dtype=np.float64),
        q_keys=np.array(keys, dtype=np.uint8).reshape(len(keys), hw),
        q_rows=np.array([ref_Q[k] for k in keys], dtype=np.float64),
        final_epsilon=np.float64(ref_eps),
    )
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-28s steps=%5d episodes=%3d states=%4d words=%6d  OK" % (
        name, n_steps, episodes, len(keys), stream.cursor))


def main():
    ref = import_reference()
    for case in CASES:
        run_case(*case, ref=ref)


if __name__ == "__main__":
    main()
