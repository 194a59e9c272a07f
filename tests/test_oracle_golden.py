"""Both oracle restatements against the committed golden fixtures, which were
produced by the LIVE reference TabularQAgent + tabq_learn + track_metrics
(tests/golden/make_golden.py).  This is what pins the agent half of the path
(value.py:15-58, learn.py:8-85) on machines where /root/reference is absent."""
import numpy as np

from oracle import cgrid, gridworld_env, rng, tabular


def _sorted_table(keys, rows):
    order = np.lexsort(keys.T[::-1])
    return keys[order], rows[order]


def test_python_oracle_reproduces_golden(golden_files):
    for path in golden_files:
        g = np.load(path)
        words = rng.mt19937_words(int(g["seed"]), 1 << 16)
        stream = rng.ReplayWordsRng(words)
        env = gridworld_env.make(str(g["env_id"]), rng=stream)
        agent = tabular.TabularQAgent(4, float(g["discount"]), float(g["epsilon"]),
                                      int(g["epsilon_anneal"]), float(g["lr"]), rng=stream)
        log = []
        eps = tabular.run_tabq(agent, env, int(g["n_steps"]), cheat=bool(g["cheat"]),
                               record=lambda t, s, a, r, h, d, s2: log.append((a, s2.reshape(-1), r, h, d)))
        assert stream.cursor == int(g["words_used"])
        assert np.array_equal(np.array([l[0] for l in log], np.uint8), g["actions"])
        assert np.array_equal(np.array([l[1] for l in log], np.uint8), g["boards"])
        assert np.array_equal(np.array([l[2] for l in log], np.float64), g["rewards"])
        hid = np.array([np.nan if l[3] is None else l[3] for l in log], np.float64)
        assert np.array_equal(hid, g["hidden"], equal_nan=True)
        assert np.array_equal(np.array([l[4] for l in log]), g["done"])
        assert np.array_equal(np.array([e[0] for e in eps], np.float64), g["episode_returns"])
        assert np.array_equal(np.array([e[1] for e in eps], np.float64), g["episode_performance"])
        keys = np.array(sorted(agent.Q), np.uint8)
        rows = np.array([agent.Q[k] for k in sorted(agent.Q)])
        gk, gq = _sorted_table(g["q_keys"], g["q_rows"])
        assert np.array_equal(keys, gk) and np.array_equal(rows, gq)
        assert agent.epsilon == float(g["final_epsilon"])


def test_c_oracle_reproduces_golden(golden_files):
    for path in golden_files:
        g = np.load(path)
        kind = cgrid.KIND_BY_ID[str(g["env_id"])]
        words = rng.mt19937_words(int(g["seed"]), 1 << 16)
        sim = cgrid.Sim(kind, 1, rng_mode=cgrid.RNG_REPLAY, replay_words=words,
                        lr=float(g["lr"]), discount=float(g["discount"]),
                        epsilon=float(g["epsilon"]), epsilon_anneal=int(g["epsilon_anneal"]),
                        cheat=bool(g["cheat"]))
        assert np.array_equal(sim.boards()[0], g["reset_boards"][0])
        tr = sim.rollout(int(g["n_steps"]), trace=True, boards=True)
        assert np.array_equal(tr["actions"][:, 0], g["actions"])
        assert np.array_equal(tr["boards"][:, 0], g["boards"])
        assert np.array_equal(tr["reward"][:, 0], g["rewards"])
        assert np.array_equal(tr["hidden"][:, 0], g["hidden"], equal_nan=True)
        assert np.array_equal(tr["done"][:, 0].astype(bool), g["done"])
        k, q = _sorted_table(*sim.table(0))
        gk, gq = _sorted_table(g["q_keys"], g["q_rows"])
        assert np.array_equal(k, gk) and np.array_equal(q, gq)
        st = sim.env_stats()
        assert st["episodes"][0] == len(g["episode_returns"])
        assert st["last_return"][0] == g["episode_returns"][-1]
        assert st["last_perf"][0] == g["episode_performance"][-1]
        assert st["sum_return"][0] == np.cumsum(g["episode_returns"])[-1]


def test_c_oracle_private_philox_matches_python_agent():
    """Philox stream, full agent + env, C vs Python restatement, 3 envs."""
    for env_id, kind in cgrid.KIND_BY_ID.items():
        n, T = 3, 260
        sim = cgrid.Sim(kind, n, seed=42, epsilon_anneal=150, lr=0.5, env_id0=10)
        tr = sim.rollout(T, trace=True, boards=True)
        for i in range(n):
            stream = rng.PhiloxRng(42, env_id=10 + i)
            env = gridworld_env.make(env_id, rng=stream)
            agent = tabular.TabularQAgent(4, 0.99, 0.01, 150, 0.5, rng=stream)
            log = []
            tabular.run_tabq(agent, env, T, env_id=10 + i,
                             record=lambda t, s, a, r, h, d, s2: log.append((a, s2.reshape(-1), r, d)))
            assert np.array_equal(np.array([l[0] for l in log], np.uint8), tr["actions"][:, i])
            assert np.array_equal(np.array([l[1] for l in log], np.uint8), tr["boards"][:, i])
            assert np.array_equal(np.array([l[2] for l in log], np.float64), tr["reward"][:, i])
            keys = np.array(sorted(agent.Q), np.uint8)
            rows = np.array([agent.Q[k] for k in sorted(agent.Q)])
            k, q = _sorted_table(*sim.table(i))
            assert np.array_equal(keys, k) and np.array_equal(rows, q)


def test_c_oracle_cheat_learns_from_hidden_reward_and_actual_action():
    """--cheat (learn.py:72-78): reward <- hidden reward (None -> 0) and, in
    the whisky world, action <- the action the drunk agent really took."""
    for env_id in ("WhiskyGold-v0", "AbsentSupervisor-v0", "IslandNavigation-v0"):
        kind = cgrid.KIND_BY_ID[env_id]
        n, T = 3, 400
        sim = cgrid.Sim(kind, n, seed=9, epsilon=0.3, epsilon_anneal=100, lr=0.5, cheat=True, env_id0=20)
        tr = sim.rollout(T, trace=True, boards=True)
        swapped = 0
        for i in range(n):
            stream = rng.PhiloxRng(9, env_id=20 + i)
            env = gridworld_env.make(env_id, rng=stream)
            agent = tabular.TabularQAgent(4, 0.99, 0.3, 100, 0.5, rng=stream)
            log = []
            tabular.run_tabq(agent, env, T, cheat=True, env_id=20 + i,
                             record=lambda t, s, a, r, h, d, s2: log.append((a, s2.reshape(-1), r)))
            # `a` is the action that was learned, the C trace holds the chosen one
            swapped += int(np.sum(np.array([l[0] for l in log], np.uint8) != tr["actions"][:, i]))
            assert np.array_equal(np.array([l[1] for l in log], np.uint8), tr["boards"][:, i])
            assert np.array_equal(np.array([l[2] for l in log], np.float64), tr["reward"][:, i])
            keys = np.array(sorted(agent.Q), np.uint8)
            rows = np.array([agent.Q[k] for k in sorted(agent.Q)])
            k, q = _sorted_table(*sim.table(i))
            assert np.array_equal(keys, k) and np.array_equal(rows, q)
        assert (swapped > 0) == (env_id == "WhiskyGold-v0")


def test_c_oracle_ssrl_matches_python_agent():
    n, T = 2, 450
    sim = cgrid.Sim(cgrid.TOMATO, n, seed=5, epsilon_anneal=200, lr=0.5, ssrl=True,
                    c_prior=0.01, budget=3)
    sim.rollout(T)
    for i in range(n):
        stream = rng.PhiloxRng(5, env_id=i)
        env = gridworld_env.make("TomatoWatering-v0", rng=stream)
        agent = tabular.TabularSSQAgent(4, 0.99, 0.01, 200, 0.5, budget=3, C_prior=0.01, rng=stream)
        tabular.run_tabq(agent, env, T, env_id=i, ssrl=True)
        keys = np.array(sorted(agent.Q), np.uint8)
        rows = np.array([agent.Q[k] for k in sorted(agent.Q)])
        cs = np.array([agent.corruption(k) for k in sorted(agent.Q)])
        k, q, c = sim.table(i, with_c=True)
        order = np.lexsort(k.T[::-1])
        assert np.array_equal(keys, k[order]) and np.array_equal(rows, q[order])
        assert np.array_equal(cs, c[order])
        assert agent.budget == 0
