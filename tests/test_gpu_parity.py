"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the
golden fixtures.  Bar: bit-exact boards, actions, rewards, hidden rewards,
done flags, episode metrics AND Q rows (the kernels do the reference's float64
arithmetic without FMA contraction, so the north-star 1e-5 tolerance on Q is
met with zero error)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ENVS = [("BoatRace-v0", 0), ("SideEffectsSokoban-v0", 1), ("TomatoWatering-v0", 2),
        ("DistributionalShift-v0", 3), ("IslandNavigation-v0", 4), ("AbsentSupervisor-v0", 5),
        ("WhiskyGold-v0", 6), ("SideEffectsSokoban2-v0", 7)]


def _gf():
    import gridfast
    return gridfast


def _cmp_stats(gpu_env, sim, with_hash=True):
    st = {k: v.cpu().numpy() for k, v in gpu_env.stats().items()}
    ref = sim.env_stats()
    if with_hash:
        assert np.array_equal(st["trace_hash"].view(np.uint64), ref["trace_hash"])
    assert np.array_equal(st["episodes"], ref["episodes"])
    assert np.array_equal(st["n_margin_pos"], ref["n_margin_pos"])
    for g, o in (("episode_return", "episode_return"), ("sum_return", "sum_return"),
                 ("sum_performance", "sum_perf"), ("sum_margin_pos", "sum_margin_pos")):
        assert np.array_equal(st[g], ref[o]), g
    done_any = ref["episodes"] > 0
    assert np.array_equal(st["last_return"][done_any], ref["last_return"][done_any])
    assert np.array_equal(st["last_performance"][done_any], ref["last_perf"][done_any])
    assert np.isnan(st["last_performance"][~done_any]).all()
    assert np.array_equal(st["max_return"][done_any], ref["max_return"][done_any])
    assert np.array_equal(gpu_env.render().cpu().numpy(), sim.boards())


def _cmp_table(env, agent, sim, table, with_c=False):
    got = agent.export(table, with_corruption=with_c)
    want = sim.table(table, with_c=with_c)
    wk = env.board_keys(torch.as_tensor(want[0]).to(env.device)).cpu().numpy().view(np.uint64)
    og, ow = np.argsort(got[0]), np.argsort(wk)
    assert np.array_equal(got[0][og], wk[ow]), "key sets differ"
    assert np.array_equal(got[1][og], want[1][ow]), "Q rows differ"
    if with_c:
        assert np.array_equal(got[2][og], want[2][ow]), "corruption estimates differ"


# ------------------------------------------------------------------ golden
def test_fused_rollout_reproduces_live_reference_golden(golden_files):
    """Replay the raw MT19937 words of np.random.seed(s) through the fused
    kernel: Q table, episode metrics and trajectory must equal what the LIVE
    reference TabularQAgent + tabq_learn produced (tests/golden)."""
    gf = _gf()
    from oracle import cgrid, rng
    for path in golden_files:
        g = np.load(path)
        kind = cgrid.KIND_BY_ID[str(g["env_id"])]
        words = rng.mt19937_words(int(g["seed"]), 1 << 16)
        hp = dict(lr=float(g["lr"]), discount=float(g["discount"]), epsilon=float(g["epsilon"]),
                  epsilon_anneal=int(g["epsilon_anneal"]))
        T = int(g["n_steps"])
        env = gf.BatchedEnv(str(g["env_id"]), 1)
        env.set_trace(True)
        env.set_replay_words(words.reshape(1, -1))
        first = env.reset(step=0).cpu().numpy()
        assert np.array_equal(first[0], g["reset_boards"][0])
        agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, **hp)
        # in two launches, to cover continuation across calls
        agent.rollout(T // 3, cheat=bool(g["cheat"]))
        agent.rollout(T - T // 3, cheat=bool(g["cheat"]))
        agent.check()
        # golden Q rows, bit for bit
        keys, rows = agent.export(0)
        gk = env.board_keys(torch.as_tensor(g["q_keys"]).to(env.device)).cpu().numpy().view(np.uint64)
        og, ow = np.argsort(keys), np.argsort(gk)
        assert np.array_equal(keys[og], gk[ow])
        assert np.array_equal(rows[og], g["q_rows"][ow])
        st = {k: v.cpu().numpy() for k, v in env.stats().items()}
        assert st["episodes"][0] == len(g["episode_returns"])
        assert st["last_return"][0] == g["episode_returns"][-1]
        assert st["last_performance"][0] == g["episode_performance"][-1]
        assert st["sum_return"][0] == np.cumsum(g["episode_returns"])[-1]
        # trajectory: the C oracle reproduces the golden trace exactly
        # (tests/test_oracle_golden.py), so equal hashes == equal traces
        sim = cgrid.Sim(kind, 1, rng_mode=cgrid.RNG_REPLAY, replay_words=words, cheat=bool(g["cheat"]), **hp)
        sim.rollout(T)
        assert st["trace_hash"].view(np.uint64)[0] == sim.env_stats()["trace_hash"][0]
        assert agent.epsilon_at(T) == float(g["final_epsilon"])


def test_unfused_calls_reproduce_golden(golden_files):
    """The one-call-per-reference-call API (env.step / agent.act / agent.learn
    as separate kernels) driven by the golden action stream."""
    gf = _gf()
    for path in golden_files:
        g = np.load(path)
        if str(g["env_id"]) in ("TomatoWatering-v0", "AbsentSupervisor-v0", "WhiskyGold-v0"):
            continue   # their env draws are interleaved with the agent's in the stream
        env = gf.BatchedEnv(str(g["env_id"]), 1)
        agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, lr=float(g["lr"]), discount=float(g["discount"]),
                                   epsilon=float(g["epsilon"]), epsilon_anneal=int(g["epsilon_anneal"]))
        board = env.reset(step=0)
        cheat = bool(g["cheat"])
        for t in range(int(g["n_steps"])):
            a = torch.tensor([g["actions"][t]], dtype=torch.uint8, device=env.device)
            nxt, r, h, d = env.step(a, step=t)
            assert np.array_equal(nxt.cpu().numpy()[0], g["boards"][t])
            assert r.item() == g["rewards"][t] and bool(d.item()) == bool(g["done"][t])
            hv = h.item()
            assert (np.isnan(hv) and np.isnan(g["hidden"][t])) or hv == g["hidden"][t]
            greedy = agent.act(board, t, explore=False)
            learn_r = torch.nan_to_num(h, nan=0.0) if cheat else r
            agent.learn(board, a, learn_r, nxt)
            board = nxt
            if d.item():
                board = env.reset(mask=d, step=t + 1)
            del greedy
        keys, rows = agent.export(0)
        gk = env.board_keys(torch.as_tensor(g["q_keys"]).to(env.device)).cpu().numpy().view(np.uint64)
        og, ow = np.argsort(keys), np.argsort(gk)
        assert np.array_equal(keys[og], gk[ow]) and np.array_equal(rows[og], g["q_rows"][ow])


# ------------------------------------------------------------------ oracle, Philox
@pytest.mark.parametrize("env_id,kind", ENVS)
def test_unfused_step_matches_oracle(env_id, kind):
    gf = _gf()
    from oracle import cgrid
    n, T, seed = 1000, 260, 11
    env = gf.BatchedEnv(env_id, n, seed=seed, env_id0=77)
    env.set_trace(True)
    sim = cgrid.Sim(kind, n, seed=seed, env_id0=77)
    assert np.array_equal(env.render().cpu().numpy(), sim.boards())
    rs = np.random.RandomState(5)
    for t in range(T):
        acts = rs.randint(0, 4, size=n).astype(np.uint8)
        b, r, h, d = env.step(torch.as_tensor(acts).to(env.device), step=t)
        ob, orr, oh, od = sim.step(acts)
        assert np.array_equal(b.cpu().numpy(), ob)
        assert np.array_equal(r.cpu().numpy(), orr)
        assert np.array_equal(h.cpu().numpy(), oh, equal_nan=True)
        assert np.array_equal(d.cpu().numpy(), od)
        if od.any():
            env.reset(mask=d, step=t + 1, want_boards=False)
    _cmp_stats(env, sim)
    obs = env.boards_to_f32(env.render())
    assert obs.dtype == torch.float32 and tuple(obs.shape[1:]) == env.shape
    assert np.array_equal(obs.cpu().numpy().reshape(n, -1), sim.boards().astype(np.float32))


@pytest.mark.parametrize("env_id,kind", ENVS)
@pytest.mark.parametrize("cheat", [False, True])
def test_fused_private_matches_oracle(env_id, kind, cheat):
    gf = _gf()
    from oracle import cgrid
    n, T, seed = (2048, 700, 3) if kind != 2 else (512, 500, 3)
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=400)
    env = gf.BatchedEnv(env_id, n, seed=seed)
    env.set_trace(True)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, **hp)
    for chunk in (1, 99, T - 100):
        agent.rollout(chunk, cheat=cheat)
    agent.check()
    sim = cgrid.Sim(kind, n, seed=seed, cheat=cheat, **hp)
    sim.rollout(T)
    _cmp_stats(env, sim)
    for i in (0, 1, 31, 32, n // 2 + 5, n - 1):
        _cmp_table(env, agent, sim, i)


@pytest.mark.parametrize("env_id,kind", ENVS)
def test_fused_shared_matches_oracle(env_id, kind):
    gf = _gf()
    from oracle import cgrid
    n, T, seed = (3000, 400, 9) if kind != 2 else (700, 300, 9)
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=200)
    env = gf.BatchedEnv(env_id, n, seed=seed)
    env.set_trace(True)
    agent = gf.BatchedTabularQ(env, gf.Q_SHARED, **hp)
    for chunk in (1, 120, T - 121):
        agent.rollout(chunk)
    agent.check()
    sim = cgrid.Sim(kind, n, seed=seed, q_mode=cgrid.Q_SHARED, **hp)
    sim.rollout(T)
    _cmp_stats(env, sim)
    _cmp_table(env, agent, sim, 0)


@pytest.mark.parametrize("q_mode", ["private", "shared"])
def test_whisky_cheat_learns_the_action_really_taken(q_mode):
    """learn.py:74-78: under --cheat a drunk agent's update goes to the action
    the environment executed; high epsilon so that the bottle is found often."""
    gf = _gf()
    from oracle import cgrid
    n, T, seed = 2048, 350, 21
    hp = dict(lr=0.5, discount=0.99, epsilon=0.4, epsilon_anneal=100)
    shared = q_mode == "shared"
    env = gf.BatchedEnv("WhiskyGold-v0", n, seed=seed)
    env.set_trace(True)
    agent = gf.BatchedTabularQ(env, gf.Q_SHARED if shared else gf.Q_PRIVATE, **hp)
    for chunk in (1, 99, T - 100):
        agent.rollout(chunk, cheat=True)
    agent.check()
    sim = cgrid.Sim(cgrid.WHISKY, n, seed=seed, q_mode=cgrid.Q_SHARED if shared else cgrid.Q_PRIVATE,
                    cheat=True, **hp)
    sim.rollout(T)
    _cmp_stats(env, sim)
    for i in ([0] if shared else [0, 1, n // 2, n - 1]):
        _cmp_table(env, agent, sim, i)
    # the swap matters: without it the tables differ
    plain = gf.BatchedEnv("WhiskyGold-v0", n, seed=seed)
    other = gf.BatchedTabularQ(plain, gf.Q_SHARED if shared else gf.Q_PRIVATE, **hp)
    other.rollout(T, cheat=False)
    assert not np.array_equal(other.export(0)[1], agent.export(0)[1])


def test_unfused_whisky_reports_actual_actions():
    gf = _gf()
    from oracle import gridworld_env, rng
    n, T, seed = 64, 120, 3
    env = gf.BatchedEnv("WhiskyGold-v0", n, seed=seed)
    env.reset(step=0)
    acts = np.random.RandomState(5).randint(0, 4, size=(T, n)).astype(np.uint8)
    got = np.zeros((T, n), np.uint8)
    for t in range(T):
        _, _, _, done = env.step(torch.as_tensor(acts[t]).to(env.device), step=t)
        got[t] = env.actual_actions().cpu().numpy()
        if done.any():
            env.reset(mask=done, step=t + 1)
    assert (got != acts).any()
    for i in (0, 17, 63):
        stream = rng.PhiloxRng(seed, env_id=i)
        o = gridworld_env.make("WhiskyGold-v0", rng=stream)
        stream.set_context(i, 0)
        o.reset()
        for t in range(T):
            stream.set_context(i, t)
            _, _, d, info = o.step(int(acts[t, i]))
            assert int(info["extra_observations"]["actual_actions"]) == got[t, i]
            if d:
                stream.set_context(i, t + 1)
                o.reset()


def test_shared_with_one_env_is_the_reference_agent(golden_files):
    """A shared table with N == 1 degenerates to the reference's sequential
    Q-learning: check it against the live-reference golden Q rows."""
    gf = _gf()
    from oracle import rng
    g = np.load([p for p in golden_files if "sokoban_tabq_seed5" in p][0])
    words = rng.mt19937_words(int(g["seed"]), 1 << 16)
    env = gf.BatchedEnv(str(g["env_id"]), 1)
    env.set_replay_words(words.reshape(1, -1))
    env.reset(step=0)
    agent = gf.BatchedTabularQ(env, gf.Q_SHARED, lr=float(g["lr"]), discount=float(g["discount"]),
                               epsilon=float(g["epsilon"]), epsilon_anneal=int(g["epsilon_anneal"]))
    agent.rollout(int(g["n_steps"]))
    agent.check()
    keys, rows = agent.export(0)
    gk = env.board_keys(torch.as_tensor(g["q_keys"]).to(env.device)).cpu().numpy().view(np.uint64)
    og, ow = np.argsort(keys), np.argsort(gk)
    assert np.array_equal(keys[og], gk[ow]) and np.array_equal(rows[og], g["q_rows"][ow])


@pytest.mark.parametrize("env_id,kind", [("TomatoWatering-v0", 2), ("BoatRace-v0", 0), ("AbsentSupervisor-v0", 5)])
def test_fused_ssrl_matches_oracle(env_id, kind):
    """SSRL over hashed tables (tomato, supervisor) and over the dense boat
    tables that live in shared memory during the rollout."""
    gf = _gf()
    from oracle import cgrid
    n, T, seed = 256, 650, 4
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=300)
    env = gf.BatchedEnv(env_id, n, seed=seed)
    env.set_trace(True)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, **hp)
    agent.enable_ssrl(c_prior=0.01, budget=4)
    agent.rollout(250)
    agent.rollout(T - 250)
    agent.check()
    sim = cgrid.Sim(kind, n, seed=seed, ssrl=True, c_prior=0.01, budget=4, **hp)
    sim.rollout(T)
    _cmp_stats(env, sim)
    for i in (0, 17, n - 1):
        _cmp_table(env, agent, sim, i, with_c=True)


def test_random_policy_rollout_matches_oracle():
    gf = _gf()
    from oracle import cgrid
    for env_id, kind in ENVS:
        env = gf.BatchedEnv(env_id, 777, seed=21)
        env.set_trace(True)
        env.rollout_random(333)
        sim = cgrid.Sim(kind, 777, seed=21)
        sim.rollout_random(333)
        _cmp_stats(env, sim)


def test_sharding_does_not_change_any_trajectory():
    """Global env ids key the streams: one object of 1024 environments equals
    two objects of 512 with env_id0 = 0 / 512."""
    gf = _gf()
    hp = dict(lr=0.5, epsilon_anneal=300)
    whole = gf.BatchedEnv("TomatoWatering-v0", 1024, seed=8)
    whole.set_trace(True)
    aw = gf.BatchedTabularQ(whole, gf.Q_PRIVATE, **hp)
    aw.rollout(400)
    hw = whole.stats()["trace_hash"].cpu().numpy()
    parts = []
    for k in range(2):
        e = gf.BatchedEnv("TomatoWatering-v0", 512, seed=8, env_id0=512 * k)
        e.set_trace(True)
        a = gf.BatchedTabularQ(e, gf.Q_PRIVATE, **hp)
        a.rollout(400)
        parts.append(e.stats()["trace_hash"].cpu().numpy())
    assert np.array_equal(hw, np.concatenate(parts))


def test_table_full_and_replay_dry_are_reported():
    gf = _gf()
    env = gf.BatchedEnv("TomatoWatering-v0", 64, seed=1)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, capacity=8, epsilon_anneal=50)
    agent.set_auto_grow(False)      # pinned capacity: overflow must be reported, never silently absorbed
    agent.rollout(300)
    with pytest.raises(gf.SgkError, match="ran out of slots"):
        agent.check()
    with pytest.raises(gf.SgkError, match="ran out of slots"):
        env.totals()                 # every synchronising call reports it
    env = gf.BatchedEnv("BoatRace-v0", 2, seed=1)
    env.set_replay_words(np.zeros((2, 10), np.uint32))
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE)
    agent.rollout(50)
    with pytest.raises(gf.SgkError, match="ran dry"):
        agent.check()


@pytest.mark.parametrize("env_id,kind", [("BoatRace-v0", 0), ("SideEffectsSokoban-v0", 1)])
@pytest.mark.parametrize("n", [1, 33, 129])
def test_trace_free_product_kernels_at_ragged_sizes(env_id, kind, n):
    """The two specialised product kernels (boat: dense tables + counted rewards,
    sokoban: perfect-index tables) with environment counts that leave a warp
    and a block partly empty, odd call lengths and an odd first step."""
    gf = _gf()
    from oracle import cgrid
    seed = 5 + n
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=300)
    env = gf.BatchedEnv(env_id, n, seed=seed)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, **hp)
    sim = cgrid.Sim(kind, n, seed=seed, **hp)
    for chunk in (3, 200, 1, 97):
        agent.rollout(chunk)
        sim.rollout(chunk)
    agent.check()
    _cmp_stats(env, sim, with_hash=False)
    for i in sorted({0, n // 2, n - 1}):
        _cmp_table(env, agent, sim, i)


def test_perfect_index_table_rejects_a_board_it_has_no_slot_for():
    """Sokoban's default private tables are addressed by the ranks of the agent
    and box cells among the non-wall cells.  A board that is not an observation
    of the level (here: the box painted onto a wall cell) has no slot: the
    unfused call must not touch any row and the next synchronising call must
    fail loudly -- never alias another state's row."""
    gf = _gf()
    env = gf.BatchedEnv("SideEffectsSokoban-v0", 4, seed=2)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE)
    boards = env.reset().to(torch.uint8).reshape(4, -1).clone()
    good = agent.act(boards, step=5)
    agent.check()
    before = [agent.export(i) for i in range(4)]
    bad = boards.clone()
    bad[bad == 4] = 1              # lift the box ...
    bad[:, 0] = 4                  # ... onto the corner wall
    agent.act(bad, step=6)
    with pytest.raises(gf.SgkError):
        agent.check()
    for i in range(4):             # the legitimate rows and keys are untouched
        keys, rows = agent.export(i)[:2]
        assert np.array_equal(keys, before[i][0]) and np.array_equal(rows, before[i][1])
    assert good.shape[0] == 4


# ------------------------------------------------------------------ full size
def test_full_size_boat_65536_envs_bit_exact():
    """BASELINE config 2 at full width: 65,536 lock-step boat races, private Q,
    1,000 lock-steps = 65.5M env-steps, every trajectory and every episode
    metric equal to the oracle's; Q tables compared for a sample."""
    gf = _gf()
    from oracle import cgrid
    n, T, seed = 65536, 1000, 0
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=100000)
    env = gf.BatchedEnv("BoatRace-v0", n, seed=seed)
    env.set_trace(True)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, **hp)
    agent.rollout(T)
    agent.check()
    sim = cgrid.Sim(cgrid.BOAT, n, seed=seed, **hp)
    sim.rollout(T)
    _cmp_stats(env, sim)
    for i in (0, 12345, n - 1):
        _cmp_table(env, agent, sim, i)
    tot = env.totals()
    ref = sim.env_stats()
    assert tot["episodes"] == ref["episodes"].sum() == n * 10
    assert tot["sum_return"] == ref["sum_return"].sum()       # integers: exact in any order
    assert tot["sum_performance"] == ref["sum_perf"].sum()
    # size-independent property: the trace-free fast kernel gives the same result
    env2 = gf.BatchedEnv("BoatRace-v0", n, seed=seed)
    agent2 = gf.BatchedTabularQ(env2, gf.Q_PRIVATE, **hp)
    agent2.rollout(T)
    assert torch.equal(env2.core(), env.core())
    assert torch.equal(env2.stats()["sum_return"], env.stats()["sum_return"])


def test_full_size_boat_annealed_phase_with_the_trace_free_kernel():
    """The kernel bench.py times (TRACE = 0 build, dense tables in shared memory)
    in the phase most of its launches run in: 65,536 environments started at
    agent-step 100,000 -- epsilon annealed to its floor, greedy lock-step --
    for 3,000 lock-steps, against the C oracle: every environment's episode
    statistics and board, and the key set and every Q row of 67 tables spread
    over the batch."""
    gf = _gf()
    from oracle import cgrid
    n, T, seed, t0 = 65536, 3000, 21, 100000
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=100000)
    env = gf.BatchedEnv("BoatRace-v0", n, seed=seed)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, **hp)
    env.t = t0
    agent.rollout(T)                      # trace off: the production instantiation
    agent.check()
    assert agent.epsilon_at(t0 + 5) == agent.epsilon_at(t0 + T)      # on the floor of the schedule
    sim = cgrid.Sim(cgrid.BOAT, n, seed=seed, **hp)
    sim.t = t0
    sim.rollout(T)
    _cmp_stats(env, sim, with_hash=False)
    for i in list(range(0, n, 997)) + [n - 1]:
        _cmp_table(env, agent, sim, i)
    tot = env.totals()
    ref = sim.env_stats()
    assert tot["episodes"] == ref["episodes"].sum() == n * (T // 100)
    assert tot["sum_return"] == ref["sum_return"].sum()              # integer-valued: exact in any order
    assert tot["sum_performance"] == ref["sum_perf"].sum()


@pytest.mark.parametrize("cheat", [False, True])
def test_trace_free_boat_kernel_in_the_exploring_phase_chunked(cheat):
    """The product instantiation again (TRACE = 0, dense tables, step pairs with
    look-ahead Philox, counted rewards, reward mode compiled in), this time
    where its special cases live: epsilon still annealing, calls that start on
    odd and even agent-steps, end mid-episode, on an episode boundary and right
    after one, single-step calls, with and without --cheat.  Against the C
    oracle: per-environment statistics and boards, key sets and Q rows."""
    gf = _gf()
    from oracle import cgrid
    n, seed = 4096, 17
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=900)
    chunks = (1, 1, 7, 90, 1, 100, 101, 250, 49, 2, 298)           # 900 lock-steps; boundaries at odd and even steps
    env = gf.BatchedEnv("BoatRace-v0", n, seed=seed)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, **hp)
    sim = cgrid.Sim(cgrid.BOAT, n, seed=seed, cheat=cheat, **hp)
    done = 0
    for j, chunk in enumerate(chunks):
        agent.rollout(chunk, cheat=cheat)                           # trace off
        sim.rollout(chunk)
        done += chunk
        if j in (2, 5, 6, len(chunks) - 1):                          # mid-episode, on a boundary, one past it, the end
            _cmp_stats(env, sim, with_hash=False)
            for i in (0, 31, 32, 1000, n - 1):
                _cmp_table(env, agent, sim, i)
    agent.check()
    assert done == 900
    tot = env.totals()
    assert tot["episodes"] == n * 9
    assert tot["sum_return"] == sim.env_stats()["sum_return"].sum()


def test_boat_hashed_and_dense_tables_agree_with_oracle():
    """Boat race private tables default to the minimal-perfect-hash layout
    (capacity 8); the generic hashed layout (capacity 16) must give the same
    trajectories, Q rows and key sets."""
    gf = _gf()
    from oracle import cgrid
    n, T, seed = 1500, 450, 6
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=300)
    sim = cgrid.Sim(cgrid.BOAT, n, seed=seed, **hp)
    sim.rollout(T)
    for capacity in (8, 16):
        env = gf.BatchedEnv("BoatRace-v0", n, seed=seed)
        env.set_trace(True)
        agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, capacity=capacity, **hp)
        assert agent.capacity == capacity
        for chunk in (7, 93, T - 100):       # ends mid-episode and on an episode boundary
            agent.rollout(chunk)
        agent.check()
        _cmp_stats(env, sim)
        for i in (0, 3, 700, n - 1):
            _cmp_table(env, agent, sim, i)


@pytest.mark.parametrize("cheat", [False, True])
def test_sokoban_hashed_and_perfect_index_tables_agree_with_oracle(cheat):
    """Sokoban level 0 private tables default to the perfect index
    rank(agent) * 11 + rank(box) (capacity 128, no key reads, no probing); the
    generic hashed layout (capacity 256) must give the same trajectories, Q
    rows and key sets -- with the trace kernels and with the trace-free ones,
    in calls that end mid-episode, on an episode end and after single steps."""
    gf = _gf()
    from oracle import cgrid
    n, seed = 3000, 8
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=500)
    chunks = (1, 7, 93, 1, 250, 148)
    sim = cgrid.Sim(cgrid.SOKOBAN, n, seed=seed, cheat=cheat, **hp)
    sim.rollout(sum(chunks))
    for capacity in (128, 256):
        for trace in (True, False):
            env = gf.BatchedEnv("SideEffectsSokoban-v0", n, seed=seed)
            env.set_trace(trace)
            agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, capacity=capacity, **hp)
            assert agent.capacity == capacity
            for chunk in chunks:
                agent.rollout(chunk, cheat=cheat)
            agent.check()
            _cmp_stats(env, sim, with_hash=trace)
            for i in (0, 3, 31, 32, 1700, n - 1):
                _cmp_table(env, agent, sim, i)


def test_replica_sync_of_two_shared_tables_on_one_gpu():
    """sgk_tabq_delta_export / restore_base / delta_apply / rebase: two
    replicas trained on different environment shards end bit-identical and
    equal to base + mean of their changes."""
    gf = _gf()
    hp = dict(lr=0.5, epsilon_anneal=200)
    reps = []
    for r in range(2):
        env = gf.BatchedEnv("SideEffectsSokoban-v0", 512, seed=4, env_id0=512 * r)
        reps.append((env, gf.BatchedTabularQ(env, gf.Q_SHARED, **hp)))
    for rnd in range(3):
        before = [dict(zip(*[x.tolist() for x in a.export(0)])) for _, a in reps]
        for _, a in reps:
            a.rollout(60)
        after = [dict(zip(*[x.tolist() for x in a.export(0)])) for _, a in reps]
        exported = [a.delta_export() for _, a in reps]
        for _, a in reps:
            a.restore_base()
            for keys, delta in exported:
                a.delta_apply(keys, delta, 0.5)
            a.rebase()
            a.check()
        merged = [dict(zip(*[x.tolist() for x in a.export(0)])) for _, a in reps]
        assert merged[0] == merged[1]
        assert before[0] == before[1] or rnd == 0
        for key, row in merged[0].items():
            base = np.array(before[0].get(key, [0.0] * 4))
            want = base.copy()
            for rep in after:
                want = want + 0.5 * (np.array(rep.get(key, base.tolist() if key in before[0] else [0.0] * 4)) - base)
            assert np.allclose(row, want, rtol=0, atol=1e-12), (key, row, want)


@pytest.mark.parametrize("env_id,kind", ENVS)
@pytest.mark.parametrize("shared", [False, True])
def test_batched_greedy_eval_matches_oracle(env_id, kind, shared):
    """default_eval semantics (eval.py:8-56) on fresh environments with the
    trained table(s): per-environment episode counts, sums and maxima."""
    gf = _gf()
    from oracle import cgrid
    from gridfast import metrics
    n, T, seed = 600, 350, 13
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=150)
    env = gf.BatchedEnv(env_id, n, seed=seed)
    agent = gf.BatchedTabularQ(env, gf.Q_SHARED if shared else gf.Q_PRIVATE, **hp)
    agent.rollout(T)
    sim = cgrid.Sim(kind, n, seed=seed, q_mode=cgrid.Q_SHARED if shared else cgrid.Q_PRIVATE, **hp)
    sim.rollout(T)
    eval_env = gf.BatchedEnv(env_id, n, seed=seed + 1000, env_id0=50)
    summary = agent.evaluate(eval_env, eval_timesteps=230)
    agent.check()
    ref = sim.evaluate(seed + 1000, 50, n, 230)
    st = {k: v.cpu().numpy() for k, v in eval_env.stats().items()}
    assert np.array_equal(st["episodes"], ref["episodes"]) and ref["episodes"].min() >= 1
    for g, o in (("sum_return", "sum_return"), ("sum_performance", "sum_perf"), ("sum_margin_pos", "sum_margin_pos"),
                 ("max_return", "max_return"), ("max_performance", "max_perf"), ("max_margin", "max_margin")):
        assert np.array_equal(st[g], ref[o]), g
    assert summary["returns"]["max"] == ref["max_return"].max()
    assert summary["safeties"]["max"] == ref["max_perf"].max()
    assert abs(summary["returns"]["avg"] - ref["sum_return"].sum() / ref["episodes"].sum()) < 1e-9
    # evaluation is read-only: a second evaluation gives the same numbers
    assert agent.evaluate(gf.BatchedEnv(env_id, n, seed=seed + 1000, env_id0=50), 230) == summary

    class Writer:
        def __init__(self):
            self.scalars = {}

        def add_scalar(self, name, value, step):
            self.scalars[name] = (value, step)

        def add_scalars(self, name, values, step):
            self.scalars[name] = (values, step)

    w = Writer()
    metrics.log_train(w, env, agent)
    metrics.log_eval(w, summary, 0)
    assert w.scalars["Train/epsilon"] == (agent.epsilon_at(T), T)
    assert {"Train/returns", "Train/safeties", "Train/margins", "Evaluation/returns",
            "Evaluation/safeties", "Evaluation/margins"} <= set(w.scalars)
    assert set(w.scalars["Evaluation/returns"][0]) == {"avg", "max"}


def test_bad_arguments_fail_loudly():
    """Error behaviour at the boundary: every misuse raises SgkError with the
    library's message; nothing is silently clamped."""
    gf = _gf()
    with pytest.raises(ValueError):
        gf.BatchedEnv("NoSuchEnv-v0", 4)
    with pytest.raises(gf.SgkError, match="n_envs"):
        gf.BatchedEnv("BoatRace-v0", 0)
    env = gf.BatchedEnv("BoatRace-v0", 3)          # ragged: far below one block
    assert env.render().shape == (3, 25)
    with pytest.raises(gf.SgkError, match="power of two"):
        gf.BatchedTabularQ(env, gf.Q_PRIVATE, capacity=12)
    with pytest.raises(gf.SgkError, match="epsilon_anneal"):
        gf.BatchedTabularQ(env, gf.Q_PRIVATE, epsilon_anneal=0)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE)
    other = gf.BatchedEnv("SideEffectsSokoban-v0", 3)
    with pytest.raises(gf.SgkError, match="different environment"):
        gf._lib.check(env.L.sgk_rollout_tabq(other.h, agent.h, 5, 0, 0, None))
    with pytest.raises(gf.SgkError, match="n_steps"):
        agent.rollout(0)
    shared = gf.BatchedTabularQ(env, gf.Q_SHARED)
    with pytest.raises(gf.SgkError, match="private"):
        shared.enable_ssrl(0.01, 5)
    with pytest.raises(gf.SgkError, match="shared"):
        agent.delta_export()
    # a single environment and a single lock-step still work
    one = gf.BatchedEnv("TomatoWatering-v0", 1, seed=3)
    a1 = gf.BatchedTabularQ(one, gf.Q_SHARED)
    a1.rollout(1)
    a1.check()
    assert one.t == 1


# ------------------------------------------------------------------ full size, configs 3 and 4
def test_full_size_sokoban_131072_envs_bit_exact():
    """BASELINE config 3, one GPU's share: 131,072 side-effects-sokoban
    environments, private Q, hidden-reward (safety performance) tracking;
    500 lock-steps = 65.5M env-steps against the (threaded) C oracle."""
    gf = _gf()
    from oracle import cgrid
    n, T, seed = 131072, 500, 1
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=100000)
    env = gf.BatchedEnv("SideEffectsSokoban-v0", n, seed=seed)
    env.set_trace(True)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, **hp)
    agent.rollout(T)
    agent.check()
    sim = cgrid.Sim(cgrid.SOKOBAN, n, seed=seed, **hp)
    sim.rollout(T)
    _cmp_stats(env, sim)
    for i in (0, 65535, n - 1):
        _cmp_table(env, agent, sim, i)
    tot, ref = env.totals(), sim.env_stats()
    assert tot["episodes"] == ref["episodes"].sum()
    assert tot["sum_performance"] == ref["sum_perf"].sum()      # integer-valued: exact in any order
    assert tot["max_margin"] == (ref["max_margin"][ref["episodes"] > 0]).max()


def test_full_size_tomato_65536_envs_ssrl_and_shared():
    """BASELINE config 4: 65,536 tomato-watering environments -- the SSRL agent
    (private tables, corruption estimates, query budget) and the shared-table
    agent, 200 lock-steps each, bit-exact against the C oracle."""
    gf = _gf()
    from oracle import cgrid
    n, T, seed = 65536, 200, 2
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=100000)
    env = gf.BatchedEnv("TomatoWatering-v0", n, seed=seed)
    env.set_trace(True)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, capacity=1024, **hp)
    agent.enable_ssrl(c_prior=0.01, budget=1)
    agent.rollout(T)
    agent.check()
    sim = cgrid.Sim(cgrid.TOMATO, n, seed=seed, ssrl=True, c_prior=0.01, budget=1, **hp)
    sim.rollout(T)
    _cmp_stats(env, sim)
    for i in (0, 40000, n - 1):
        _cmp_table(env, agent, sim, i, with_c=True)
    del agent, env, sim
    env = gf.BatchedEnv("TomatoWatering-v0", n, seed=seed)
    env.set_trace(True)
    agent = gf.BatchedTabularQ(env, gf.Q_SHARED, **hp)
    agent.rollout(T)
    agent.check()
    sim = cgrid.Sim(cgrid.TOMATO, n, seed=seed, q_mode=cgrid.Q_SHARED, **hp)
    sim.rollout(T)
    _cmp_stats(env, sim)
    _cmp_table(env, agent, sim, 0)


# ------------------------------------------------------------------ round-2 entry points
def test_dense_delta_sync_equals_the_record_sync():
    """The one-all-reduce form of the replica sync (dense canonical index, sgk_tabq_delta_export_dense /
    _apply_dense) must leave every replica with exactly the table the (key, delta) record form leaves."""
    gf = _gf()
    hp = dict(lr=0.5, epsilon_anneal=200)
    results = {}
    for mode in ("records", "dense"):
        reps = []
        for r in range(2):
            env = gf.BatchedEnv("SideEffectsSokoban-v0", 512, seed=4, env_id0=512 * r)
            reps.append((env, gf.BatchedTabularQ(env, gf.Q_SHARED, **hp)))
        assert reps[0][1].dense_size() == 36 * 36
        for rnd in range(3):
            for _, a in reps:
                a.rollout(60)
            if mode == "records":
                exported = [a.delta_export() for _, a in reps]
                for _, a in reps:
                    a.restore_base()
                    for keys, delta in exported:
                        a.delta_apply(keys, delta, 0.5)
                    a.rebase()
            else:
                total = reps[0][1].delta_export_dense() + reps[1][1].delta_export_dense()     # what the all-reduce computes
                for _, a in reps:
                    a.delta_apply_dense(total, 0.5)
            for _, a in reps:
                a.check()
        tables = []
        for _, a in reps:
            keys, rows = a.export(0)
            order = np.argsort(keys)
            tables.append((keys[order], rows[order]))
        assert np.array_equal(tables[0][0], tables[1][0]) and np.array_equal(tables[0][1], tables[1][1])
        results[mode] = tables[0]
    assert np.array_equal(results["records"][0], results["dense"][0])
    # the record form adds the two scaled deltas one after the other, the dense form scales their sum:
    # identical up to one rounding of the sum
    assert np.allclose(results["records"][1], results["dense"][1], rtol=1e-15, atol=1e-13)


@pytest.mark.parametrize("env_id,kind", ENVS)
def test_keys_decode_back_into_the_boards_they_stand_for(env_id, kind):
    """sgk_key_to_board is the inverse of sgk_board_to_key on every observation a rollout produces (what
    lets the adapters show the device table as the reference's dict keyed by board tuples)."""
    gf = _gf()
    env = gf.BatchedEnv(env_id, 2048, seed=3)
    seen = [env.render()]
    for _ in range(6):
        env.rollout_random(17)
        seen.append(env.render())
    boards = torch.cat(seen)
    keys = env.board_keys(boards)
    assert torch.equal(env.keys_to_boards(keys), boards)
    assert len(torch.unique(keys)) == len(torch.unique(boards, dim=0))        # lossless: one key per distinct board


def test_explicit_growth_keeps_every_table_entry():
    gf = _gf()
    from oracle import cgrid
    n, seed = 300, 12
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=200)
    env = gf.BatchedEnv("TomatoWatering-v0", n, seed=seed)
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, capacity=512, **hp)
    agent.rollout(300)
    fill = agent.max_fill()
    before = [agent.export(i) for i in (0, 150, n - 1)]
    agent.grow(2048)                      # slot-major 512 -> table-major 2048
    assert agent.capacity == 2048 and agent.max_fill() == fill
    for (k0, r0), i in zip(before, (0, 150, n - 1)):
        k1, r1 = agent.export(i)
        o0, o1 = np.argsort(k0), np.argsort(k1)
        assert np.array_equal(k0[o0], k1[o1]) and np.array_equal(r0[o0], r1[o1])
    agent.rollout(200)                    # and the run continues exactly as the oracle's
    agent.check()
    sim = cgrid.Sim(cgrid.TOMATO, n, seed=seed, **hp)
    sim.rollout(500)
    _cmp_stats(env, sim, with_hash=False)
    _cmp_table(env, agent, sim, 150)
