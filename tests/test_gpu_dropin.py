"""The fused drop-in for BASELINE config 1 (`python main.py boat tabular-q --lr .5`):
gridfast's LEARN_MAP / EVAL_MAP / WARMUP_MAP functions -- one kernel launch
per reference call -- driven the way train.train (train.py:21-81) drives the
reference's own, on numpy's global stream.

Fixtures: tests/golden/train_*.npz hold the COMPLETE scalar log, final Q table
and final stream position of the reference's real train.train run in the build
container (make_train_golden.py); tests/golden/*_tabq_*.npz the per-step traces
of the live reference agent (make_golden.py).  Everything must be reproduced
bit for bit.  Where /root/reference and a GPU exist together the real
train.train itself is run against the GPU adapters.
"""
import argparse
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

import _reference_shapes as ref_shapes

pytestmark = pytest.mark.gpu

ALIAS_BY_ID = {v: k for k, v in ref_shapes.ENV_MAP.items()}


def _gf():
    import gridfast
    return gridfast


def _registries(gf):
    agent_map, learn_map, eval_map, warmup_map = ref_shapes.empty_registries()
    gf.register_with_reference(agent_map=agent_map, learn_map=learn_map, eval_map=eval_map, warmup_map=warmup_map)
    return agent_map, learn_map, eval_map, warmup_map


def _args(**kw):
    base = dict(agent_alias="tabular-q", log_dir=None, eval_visualize_episodes=0, discount=0.99, cheat=False,
                epsilon=0.01, eval_every=10 ** 9, eval_timesteps=100)
    base.update(kw)
    return argparse.Namespace(**base)


def _q_matches(agent, q_keys, q_rows):
    Q = dict(agent.Q.items())
    want = {tuple(np.float32(v) for v in k): r for k, r in zip(q_keys, q_rows)}
    assert set(Q) == set(want), "key sets differ: %d vs %d" % (len(Q), len(want))
    for k in want:
        assert np.array_equal(Q[k], want[k]), (k, Q[k], want[k])


def test_fused_loops_reproduce_the_real_train_train_logs(train_golden_files):
    """Scalar for scalar, in order: Train/epsilon per step, the episode metrics,
    Evaluation/* per period; then every Q row, epsilon, and where numpy's
    stream stands."""
    gf = _gf()
    for path in train_golden_files:
        g = np.load(path)
        args = _args(env_alias=str(g["env_alias"]), seed=int(g["seed"]), episodes=int(g["episodes"]),
                     eval_every=int(g["eval_every"]), eval_timesteps=int(g["eval_timesteps"]), lr=float(g["lr"]),
                     epsilon_anneal=int(g["epsilon_anneal"]), cheat=bool(g["cheat"]))
        writer = ref_shapes.RecordingWriter()
        agent, env, _ = ref_shapes.train(args, lambda env_id: gf.make(env_id, rng="numpy"), *_registries(gf), writer)
        want = json.loads(str(g["events"]))
        assert len(writer.events) == len(want), (path, len(writer.events), len(want))
        for i, (a, b) in enumerate(zip(writer.events, want)):
            assert a == b, (os.path.basename(path), i, a, b)
        _q_matches(agent, g["q_keys"], g["q_rows"])
        assert agent.epsilon == float(g["final_epsilon"])
        tail = np.random.randint(0, 2 ** 32, size=4, dtype=np.uint32)
        assert np.array_equal(tail, g["stream_tail"]), "numpy's global stream ended somewhere else"


def test_fused_episodes_reproduce_live_reference_golden(golden_files):
    """The ten per-step fixtures of the live reference agent, episode-wise
    through tabq_learn_fused (no evaluation in between)."""
    gf = _gf()
    for path in golden_files:
        g = np.load(path)
        episodes = len(g["episode_returns"])
        args = _args(env_alias=ALIAS_BY_ID[str(g["env_id"])], seed=int(g["seed"]), episodes=episodes, lr=float(g["lr"]),
                     epsilon_anneal=int(g["epsilon_anneal"]), cheat=bool(g["cheat"]), discount=float(g["discount"]),
                     epsilon=float(g["epsilon"]))
        np.random.seed(args.seed)
        env = gf.make(str(g["env_id"]), rng="numpy")
        env.seed(args.seed)
        agent = gf.GpuTabularQAgent(env, args)
        history = ref_shapes.make_meters({})
        history["writer"] = ref_shapes.RecordingWriter()
        history["t"], history["episode"] = 0, 0
        returns, safeties, boards = [], [], []
        for _ in range(episodes):
            first = env.reset()
            env_state = (first, 0.0, False, {"hidden_reward": 0.0, "observed_reward": 0.0})
            history["episode"] += 1
            env_state, history, _ = gf.tabq_learn_fused(agent, env, env_state, history, args)
            returns.append(history["returns"].val)
            safeties.append(history["safeties"].val)
            boards.append(env_state[0].reshape(-1).astype(np.uint8))
        assert history["t"] == int(g["n_steps"])
        assert returns == list(g["episode_returns"]) and safeties == list(g["episode_performance"])
        ends = np.flatnonzero(g["done"])
        assert np.array_equal(np.array(boards), g["boards"][ends])             # each episode's last observation
        assert env_state[1] == g["rewards"][-1] and env_state[2] is True
        h = env_state[3]["hidden_reward"]
        assert (h is None and np.isnan(g["hidden"][-1])) or h == g["hidden"][-1]
        _q_matches(agent, g["q_keys"], g["q_rows"])
        assert agent.epsilon == float(g["final_epsilon"])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="/root/reference is absent on the GPU box")
def test_the_real_train_train_drives_the_gpu_path(train_golden_files):      # pragma: no cover - needs both
    """Where the reference and a GPU exist together: the reference's own
    train.train(args), unmodified, with gridfast registered into its registries
    and gym.make replaced -- must log exactly what it logs with its own agent."""
    gf = _gf()
    here = os.getcwd()
    crmdp = types.ModuleType("ai_safety_gridworlds.environments.tomato_crmdp")
    crmdp.REWARD_FACTOR = 0.02
    for name, mod in (("ai_safety_gridworlds", types.ModuleType("ai_safety_gridworlds")),
                      ("ai_safety_gridworlds.environments", types.ModuleType("ai_safety_gridworlds.environments")),
                      ("ai_safety_gridworlds.environments.tomato_crmdp", crmdp),
                      ("gym", types.ModuleType("gym")), ("safe_grid_gym", types.ModuleType("safe_grid_gym")),
                      ("tensorboardX", types.ModuleType("tensorboardX"))):
        sys.modules[name] = mod
    holder = {}

    class Writer(ref_shapes.RecordingWriter):
        def __init__(self, log_dir=None):
            super().__init__(log_dir)
            holder["w"] = self

    sys.modules["tensorboardX"].SummaryWriter = Writer
    sys.path.insert(0, "/root/reference")
    os.chdir("/root/reference")
    try:
        import train as ref_train
        from safe_grid_agents.common.eval import EVAL_MAP
        from safe_grid_agents.common.learn import LEARN_MAP
        from safe_grid_agents.common.warmup import WARMUP_MAP
        from safe_grid_agents.parsing import AGENT_MAP
        gf.register_with_reference(agent_map=AGENT_MAP, learn_map=LEARN_MAP, eval_map=EVAL_MAP, warmup_map=WARMUP_MAP,
                                   gym_module=sys.modules["gym"], rng="numpy")
        for path in train_golden_files:
            g = np.load(path)
            args = _args(env_alias=str(g["env_alias"]), seed=int(g["seed"]), episodes=int(g["episodes"]),
                         eval_every=int(g["eval_every"]), eval_timesteps=int(g["eval_timesteps"]), lr=float(g["lr"]),
                         epsilon_anneal=int(g["epsilon_anneal"]), cheat=bool(g["cheat"]))
            ref_train.train(args)
            assert holder["w"].events == json.loads(str(g["events"]))
    finally:
        os.chdir(here)


# ------------------------------------------------------------------ SSRL
def _ssq_args(**kw):
    return _args(agent_alias="tabular-ssq", lr=0.5, epsilon_anneal=400, budget=12, warmup=0.5, C_prior=0.01, **kw)


def _oracle_ssq(env_id, args, n_warm, n_steps):
    """The restated TabularSSQAgent + random_warmup + SSRL loop on numpy's stream."""
    from oracle import gridworld_env, tabular
    np.random.seed(args.seed)
    env = gridworld_env.make(env_id)
    env.seed(args.seed)
    agent = tabular.TabularSSQAgent(4, args.discount, args.epsilon, args.epsilon_anneal, args.lr, args.budget, args.C_prior)
    if args.seed:
        np.random.seed(args.seed)               # RandomAgent.__init__ (dummy.py:10-13)
    tabular.random_warmup(agent, env, n_warm)
    eps = tabular.run_tabq(agent, env, n_steps, ssrl=True)
    return agent, eps


@pytest.mark.parametrize("env_id", ["TomatoWatering-v0", "BoatRace-v0", "AbsentSupervisor-v0"])
def test_ssq_agent_warmup_and_fused_loop_match_the_restated_reference(env_id):
    """C4's pieces at N = 1 on numpy's stream: GpuTabularSSQAgent(env, args) reading
    args.budget / .C_prior / .warmup, random_warmup_fused, then ssq_learn_fused
    episodes (query_H + learn_C at each episode end while budget lasts)."""
    gf = _gf()
    episodes = 14
    args = _ssq_args(env_alias=ALIAS_BY_ID[env_id], seed=5, episodes=episodes)
    writer = ref_shapes.RecordingWriter()
    agent, env, history = ref_shapes.train(args, lambda i: gf.make(i, rng="numpy"), *_registries(gf), writer)
    n_warm = int(args.budget * args.warmup)
    o_agent, o_eps = _oracle_ssq(env_id, args, n_warm, history["t"])
    # the final evaluation of train() touched states too: compare what learning wrote
    assert agent.episodes == o_agent.episodes == n_warm + episodes
    assert agent.corrupt_episodes == o_agent.corrupt_episodes
    assert agent.budget == o_agent.budget == max(args.budget - n_warm - episodes, 0)
    got_returns = [e[2] for e in writer.events if e[1] == "Train/returns"]
    assert got_returns == [r for r, _ in o_eps]
    Q = dict(agent.Q.items())
    for k, row in o_agent.Q.items():
        kk = tuple(np.float32(v) for v in k)
        assert kk in Q and np.array_equal(Q[kk], row), (k, Q.get(kk), row)
    C = dict(agent.C.items())
    for k, c in o_agent.C.items():
        assert C[tuple(np.float32(v) for v in k)] == c


def test_ssq_agent_methods_one_call_at_a_time():
    """query_H / learn_C / reset_history as individual calls (the reference's
    method surface, ssrl/agents.py:45-82) against the restated agent, driven by
    the same unfused loop on numpy's stream."""
    gf = _gf()
    from oracle import gridworld_env, tabular
    args = _ssq_args(env_alias="tomato", seed=9, episodes=5)

    def drive(env, agent, inner):
        np.random.seed(args.seed)
        out = []
        for _ in range(args.episodes):
            state, done = env.reset(), False
            while not done:
                action = agent.act_explore(state)
                successor, reward, done, info = env.step(action)
                agent.learn(state, action, reward, successor)
                agent.update_epsilon()
                state = successor
            if agent.budget > 0:
                safety = agent.query_H(inner(env))
                agent.learn_C(inner(env).episode_return - safety > 0)
            else:
                agent.reset_history(False)
            out.append((inner(env).episode_return, agent.budget, agent.episodes, agent.corrupt_episodes))
        return out

    o_env = gridworld_env.make("TomatoWatering-v0")
    o_agent = tabular.TabularSSQAgent(4, args.discount, args.epsilon, args.epsilon_anneal, args.lr, 3, args.C_prior)
    want = drive(o_env, o_agent, lambda e: e._env)
    args.budget = 3
    env = gf.make("TomatoWatering-v0", rng="numpy")
    agent = gf.GpuTabularSSQAgent(env, args)
    got = drive(env, agent, lambda e: e._env)
    assert got == want
    C = dict(agent.C.items())
    Q = dict(agent.Q.items())
    assert len(Q) == len(o_agent.Q)
    for k, c in o_agent.C.items():
        assert C[tuple(np.float32(v) for v in k)] == c
    for k, row in o_agent.Q.items():
        assert np.array_equal(Q[tuple(np.float32(v) for v in k)], row)


def test_batched_ssrl_warmup_and_c4_spec_match_the_c_oracle():
    """BASELINE config 4 as SURVEY 8(d) specifies it -- budget, warm-up fraction,
    random warm-up episodes, then SSRL learning -- on 4,096 tomato environments
    (the full 65,536 x budget 1,000 run is bench.py's C4 record), Philox streams,
    against the C oracle: trajectories, counters, Q and C tables."""
    gf = _gf()
    from oracle import cgrid
    n, seed, budget, warm = 4096, 6, 8, 0.5
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=300)
    env = gf.BatchedEnv("TomatoWatering-v0", n, seed=seed)
    env.set_trace(True)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, capacity=256, **hp)
    agent.enable_ssrl(c_prior=0.01, budget=budget)
    steps = agent.ssrl_warmup(int(budget * warm), want_steps=True)
    env.clear_stats()
    agent.rollout(700)
    agent.check()
    sim = cgrid.Sim(cgrid.TOMATO, n, seed=seed, ssrl=True, c_prior=0.01, budget=budget, **hp)
    o_steps = sim.ssrl_warmup(int(budget * warm))
    sim.clear_stats()
    sim.rollout(700)
    assert np.array_equal(steps.cpu().numpy(), o_steps)
    for got, want in zip(agent.ssrl_counters(), sim.ssrl_counters()):
        assert np.array_equal(got.cpu().numpy(), want)
    st = {k: v.cpu().numpy() for k, v in env.stats().items()}
    ref = sim.env_stats()
    assert np.array_equal(st["trace_hash"].view(np.uint64), ref["trace_hash"])
    assert np.array_equal(st["episodes"], ref["episodes"]) and np.array_equal(st["sum_return"], ref["sum_return"])
    assert agent.capacity > 256, "the tables were expected to grow"
    from test_gpu_parity import _cmp_table
    for i in (0, 1234, n - 1):
        _cmp_table(env, agent, sim, i, with_c=True)


# ------------------------------------------------------------------ tables grow like the dict
def test_tomato_adapter_past_ten_thousand_steps_never_fills():
    """the reachable tomato observations (29 * 2^13) exceed any
    fixed default; the drop-in must keep working past the point where a
    4,096-slot table would overflow.  130 fused episodes = 13,000 steps from a
    deliberately tiny table, against the restated agent on the same stream."""
    gf = _gf()
    from oracle import gridworld_env, tabular
    args = _args(env_alias="tomato", seed=2, episodes=130, lr=0.5, epsilon_anneal=100000, q_capacity=64)
    np.random.seed(args.seed)
    env = gf.make("TomatoWatering-v0", rng="numpy")
    agent = gf.GpuTabularQAgent(env, args)
    history = ref_shapes.make_meters({})
    history["writer"] = ref_shapes.RecordingWriter()
    history["t"], history["episode"] = 0, 0
    for _ in range(args.episodes):
        env_state = (env.reset(), 0.0, False, {})
        history["episode"] += 1
        gf.tabq_learn_fused(agent, env, env_state, history, args)
    agent.table.check()
    np.random.seed(args.seed)
    o_env = gridworld_env.make("TomatoWatering-v0")
    o_agent = tabular.TabularQAgent(4, args.discount, args.epsilon, args.epsilon_anneal, args.lr)
    tabular.run_tabq(o_agent, o_env, history["t"])
    assert history["t"] == 13000
    assert len(o_agent.Q) > 4096, "the scenario no longer exercises the overflow"
    Q = dict(agent.Q.items())
    assert len(Q) == len(o_agent.Q)
    assert agent.table.capacity >= 8192
    for k, row in o_agent.Q.items():
        assert np.array_equal(Q[tuple(np.float32(v) for v in k)], row)


def test_batched_tables_grow_mid_rollout_bit_exact():
    """2,048 private tomato tables starting at 128 slots through 2,500 lock-steps
    in ONE rollout call (cut into launches that cannot overflow, tables
    rehashed in between) -- with SSRL, whose visited-slot history must survive
    the rehash."""
    gf = _gf()
    from oracle import cgrid
    n, T, seed = 2048, 2500, 8
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=2000)
    env = gf.BatchedEnv("TomatoWatering-v0", n, seed=seed)
    env.set_trace(True)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, capacity=128, **hp)
    agent.enable_ssrl(c_prior=0.01, budget=10)
    agent.rollout(T)
    agent.check()
    assert agent.capacity >= 1024 and agent.max_fill() <= agent.capacity * 3 // 4
    sim = cgrid.Sim(cgrid.TOMATO, n, seed=seed, ssrl=True, c_prior=0.01, budget=10, **hp)
    sim.rollout(T)
    from test_gpu_parity import _cmp_stats, _cmp_table
    _cmp_stats(env, sim)
    for i in (0, 999, n - 1):
        _cmp_table(env, agent, sim, i, with_c=True)
