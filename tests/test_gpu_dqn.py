"""Deep-Q path against a plain PyTorch fp32 restatement of
DeepQAgent.learn (common/agents/value.py:113-136): same network, same batch,
same B x B-broadcast loss, clip_grad_norm_(10), Adam(amsgrad).  Floating point:
tolerance 1e-5 relative (north_star), stated per assertion."""
import warnings

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def build_Q(n_input, n_layers, n_hidden, n_actions):
    """value.py:148-158"""
    first = nn.Sequential(nn.Linear(n_input, n_hidden), nn.ReLU())
    hidden = nn.Sequential(*tuple(nn.Sequential(nn.Linear(n_hidden, n_hidden), nn.ReLU())
                                  for _ in range(n_layers - 1)))
    last = nn.Linear(n_hidden, n_actions)
    return nn.Sequential(first, hidden, last)


def reference_learn(Q, target_Q, optim, states, actions, rewards, successors, terminals, discount, bxb):
    """value.py:118-134 with the uint8 mask made boolean (torch >= 2 rejects uint8)."""
    Qs = Q(states).gather(1, actions.long().reshape(-1, 1))
    next_Qs = target_Q(successors).max(1)[0]
    next_Qs[terminals.bool()] = 0
    expected = discount * next_Qs + rewards
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loss = F.mse_loss(Qs, expected) if bxb else F.mse_loss(Qs.reshape(-1), expected)
    optim.zero_grad()
    loss.backward()
    norm = nn.utils.clip_grad_norm_(Q.parameters(), 10.0)
    optim.step()
    return loss.item(), float(norm)


def flat(module):
    return torch.cat([p.detach().reshape(-1) for m in module.modules() if isinstance(m, nn.Linear)
                      for p in (m.weight, m.bias)])


@pytest.mark.parametrize("bxb", [True, False])
@pytest.mark.parametrize("n_layers,n_hidden,batch", [(2, 100, 64), (1, 32, 200), (3, 48, 1000)])
def test_learn_step_matches_torch_fp32(bxb, n_layers, n_hidden, batch):
    import gridfast
    torch.manual_seed(7)
    dev = torch.device("cuda", 0)
    env = gridfast.BatchedEnv("SideEffectsSokoban-v0", 8, seed=1)
    agent = gridfast.BatchedDeepQ(env, n_layers=n_layers, n_hidden=n_hidden, batch_size=batch, lr=1e-3,
                                  discount=0.99, reference_bxb_loss=bxb)
    agent.set_tensor_cores(0)       # the fp32 FFMA path: parity reference for gradients and the optimiser
    Q = build_Q(env.hw, n_layers, n_hidden, 4).to(dev)
    T = build_Q(env.hw, n_layers, n_hidden, 4).to(dev)
    optim = torch.optim.Adam(Q.parameters(), lr=1e-3, amsgrad=True)      # value.py:87
    agent.load_torch_module(Q, 0)
    agent.load_torch_module(T, 1)
    assert agent.n_params == sum(p.numel() for p in Q.parameters())
    rs = np.random.RandomState(3)
    for step in range(4):
        s = torch.as_tensor(rs.randint(0, 6, size=(batch, env.hw)).astype(np.uint8)).to(dev)
        s2 = torch.as_tensor(rs.randint(0, 6, size=(batch, env.hw)).astype(np.uint8)).to(dev)
        a = torch.as_tensor(rs.randint(0, 4, size=batch).astype(np.uint8)).to(dev)
        r = torch.as_tensor(rs.choice([-1.0, 49.0, -11.0], size=batch)).to(dev)
        term = torch.as_tensor((rs.rand(batch) < 0.1).astype(np.uint8)).to(dev)
        # scores before the update (DeepQAgent.act)
        q_ours = agent.q_values(s)
        q_ref = Q(s.float())
        assert torch.allclose(q_ours, q_ref, rtol=1e-5, atol=1e-5)
        loss, norm = reference_learn(Q, T, optim, s.float(), a, r.float(), s2.float(), term, 0.99, bxb)
        ours = agent.learn_batch(s, a, r, s2, term).tolist()
        assert ours[0] == pytest.approx(loss, rel=1e-4), "loss"
        assert ours[1] == pytest.approx(norm, rel=1e-4), "gradient norm"
        # Adam's first steps move every weight by ~lr regardless of gradient scale,
        # so compare parameters absolutely: 1e-5 of the largest weight magnitude
        p_ours, p_ref = agent.get_params(0), flat(Q)
        assert (p_ours - p_ref).abs().max().item() <= 2e-5 * p_ref.abs().max().item() + 2e-6
    assert torch.equal(agent.get_params(1), flat(T)), "target network must not move"
    agent.sync_target()
    assert torch.equal(agent.get_params(1), agent.get_params(0))


def test_warmup_fills_ring_with_oracle_random_walk():
    """dqn_warmup semantics (warmup.py:8-23): random policy, ring filled in
    environment order; the environments follow the oracle's random walk."""
    import gridfast
    from oracle import cgrid
    n, T = 256, 150
    env = gridfast.BatchedEnv("SideEffectsSokoban-v0", n, seed=5)
    agent = gridfast.BatchedDeepQ(env, replay_capacity=n * 100, batch_size=64)
    agent.warmup(T)
    assert agent.replay_count == n * 100          # ring wrapped: 150 * 256 > capacity
    sim = cgrid.Sim(cgrid.SOKOBAN, n, seed=5)
    sim.rollout_random(T)
    ref = sim.env_stats()
    st = {k: v.cpu().numpy() for k, v in env.stats().items()}
    assert np.array_equal(st["episodes"], ref["episodes"])
    assert np.array_equal(st["sum_return"], ref["sum_return"])
    assert np.array_equal(env.render().cpu().numpy(), sim.boards())


@pytest.mark.parametrize("env_id", ["IslandNavigation-v0", "WhiskyGold-v0"])
def test_rollout_cheat_stores_hidden_reward_and_actual_action(env_id):
    """args.cheat in dqn_learn (learn.py:39-47).  With lr = 0 the network never
    changes, so a plain and a cheating run act identically; their rings must
    then differ exactly by reward <- hidden reward (None -> 0) and action <-
    the action the environment executed, as the unfused environment reports
    them when it is fed the plain run's actions."""
    import gridfast
    n, T = 512, 60
    rings = {}
    for cheat in (False, True):
        env = gridfast.BatchedEnv(env_id, n, seed=13)
        agent = gridfast.BatchedDeepQ(env, replay_capacity=n * T, batch_size=64, lr=0.0, epsilon=0.3,
                                      epsilon_anneal=10, seed=2)
        agent.rollout(T, cheat=cheat)
        rings[cheat] = [x.cpu().numpy().reshape(T, n, -1) for x in agent.replay_rows(0, n * T)]
    (s0, a0, r0, n0, t0), (s1, a1, r1, n1, t1) = rings[False], rings[True]
    assert np.array_equal(s0, s1) and np.array_equal(n0, n1) and np.array_equal(t0, t1)
    env = gridfast.BatchedEnv(env_id, n, seed=13)
    env.reset(step=0)
    swapped = 0
    for t in range(T):
        boards, reward, hidden, done = env.step(torch.as_tensor(a0[t, :, 0]).to(env.device), step=t)
        actual = env.actual_actions().cpu().numpy()
        assert np.array_equal(boards.cpu().numpy(), n0[t])
        assert np.array_equal(reward.cpu().numpy().astype(np.float32), r0[t, :, 0])
        assert np.array_equal(np.nan_to_num(hidden.cpu().numpy(), nan=0.0).astype(np.float32), r1[t, :, 0])
        assert np.array_equal(actual, a1[t, :, 0])
        swapped += int((actual != a0[t, :, 0]).sum())
        if done.any():
            env.reset(mask=done, step=t + 1)
    assert (swapped > 0) == (env_id == "WhiskyGold-v0")
    assert not np.array_equal(r0, r1)


@pytest.mark.parametrize("use_tc", [False, True])
def test_graph_replayed_lockstep_equals_direct_launches(use_tc):
    """Rollouts of >= 16 lock-steps replay one captured CUDA graph whose
    per-step scalars (step, epsilon threshold, ring position, ring fill, Adam
    bias corrections) come from a device table; shorter calls launch every
    kernel directly.  Both must produce the same network, ring and statistics,
    bit for bit -- across a ring wrap and several target syncs."""
    import gridfast
    n, T = 256, 48
    out = {}
    for label, chunks in (("graph", [T]), ("direct", [8] * (T // 8))):
        env = gridfast.BatchedEnv("SideEffectsSokoban-v0", n, seed=3)
        agent = gridfast.BatchedDeepQ(env, replay_capacity=n * 20, batch_size=512, lr=1e-3, epsilon=0.05,
                                      epsilon_anneal=30, sync_every=7, seed=5)
        agent.set_tensor_cores(use_tc)
        agent.warmup(10)
        for c in chunks:
            agent.rollout(c)
        ring = [x.cpu().numpy() for x in agent.replay_rows(0, n * 20)]
        out[label] = (agent.get_params(0).cpu().numpy(), agent.get_params(1).cpu().numpy(), ring,
                      {k: v.cpu().numpy() for k, v in env.stats().items()}, env.render().cpu().numpy())
    g, d = out["graph"], out["direct"]
    assert np.array_equal(g[0], d[0]) and np.array_equal(g[1], d[1])
    for a, b in zip(g[2], d[2]):
        assert np.array_equal(a, b)
    for k in g[3]:
        assert np.array_equal(g[3][k], d[3][k], equal_nan=True), k
    assert np.array_equal(g[4], d[4])


def _train_and_score(agent, env, train_steps=1500, score_steps=300):
    agent.warmup(40)
    base = env.totals()
    random_return = base["sum_return"] / base["episodes"]
    agent.rollout(train_steps)
    mid = env.totals()
    agent.rollout(score_steps)
    tot = env.totals()
    tail_return = (tot["sum_return"] - mid["sum_return"]) / (tot["episodes"] - mid["episodes"])
    return random_return, tail_return


def test_dqn_rollout_learns_sokoban():
    """End to end: warm-up, then act/step/learn lock-steps.  Statistical
    parity only (SURVEY hard part H7): after training, the epsilon-greedy
    policy must collect clearly more return per episode than the random policy
    (random ~28, optimum 45), and the bookkeeping must be exact."""
    import gridfast
    n = 512
    env = gridfast.BatchedEnv("SideEffectsSokoban-v0", n, seed=2)
    agent = gridfast.BatchedDeepQ(env, replay_capacity=n * 40, batch_size=1024, lr=1e-3, epsilon=0.05,
                                  epsilon_anneal=150, sync_every=25, reference_bxb_loss=False, seed=11)
    p0 = agent.get_params(0).clone()
    random_return, tail_return = _train_and_score(agent, env)
    loss, norm, clip = agent.last_scalars()
    assert np.isfinite(loss) and np.isfinite(norm) and 0 < clip <= 1
    assert not torch.equal(agent.get_params(0), p0)
    assert tail_return > random_return + 5, (random_return, tail_return)
    assert env.t == 40 + 1500 + 300
    boards = env.render()
    q = agent.q_values(boards)
    assert q.shape == (n, 4) and torch.isfinite(q).all()
    assert agent.act(boards).dtype == torch.uint8


# ---------------------------------------------------------------- tensor cores
@pytest.mark.parametrize("env_id", ["BoatRace-v0", "SideEffectsSokoban-v0", "TomatoWatering-v0"])
@pytest.mark.parametrize("rows", [1, 127, 128, 1000, 70001])
def test_tcgen05_forward_matches_fp32(env_id, rows):
    """The fused tcgen05 MLP forward (fp32 accumulate in TMEM, activations
    resident in tensor memory) against torch fp32 (float64 as the arbiter):
      * the DEFAULT mode, 3xTF32: within the 1e-5 north_star states for Q --
        no further from the float64 result than torch's own fp32 is, plus 1e-5;
      * single-pass TF32 (mode 1): 10 mantissa bits, 3e-3 of the largest |Q|
        (stated, not 1e-5)."""
    import gridfast
    torch.manual_seed(rows)
    dev = torch.device("cuda", 0)
    env = gridfast.BatchedEnv(env_id, 4, seed=1)
    agent = gridfast.BatchedDeepQ(env, n_layers=2, n_hidden=100)
    assert agent.tensor_core_mode == 3, "tensor cores (3xTF32) are the default for the reference's architecture"
    Q = build_Q(env.hw, 2, 100, 4).to(dev)
    agent.load_torch_module(Q, 0)
    agent.load_torch_module(Q, 1)
    boards = torch.randint(0, 6, (rows, env.hw), dtype=torch.uint8, device=dev)
    q_x3 = agent.q_values(boards)
    agent.set_tensor_cores(0)
    q_fp32 = agent.q_values(boards)
    agent.set_tensor_cores(1)
    q_tc = agent.q_values(boards)
    q_tc_target = agent.q_values(boards, which=1)
    q_ref = Q(boards.float())
    q_f64 = Q.double()(boards.double())
    Q.float()
    scale = q_ref.abs().max().item()
    assert torch.allclose(q_fp32, q_ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(q_x3, q_ref, rtol=1e-5, atol=1e-5 * scale), (q_x3 - q_ref).abs().max().item()
    err_x3 = (q_x3.double() - q_f64).abs().max().item()
    err_torch = (q_ref.double() - q_f64).abs().max().item()
    assert err_x3 <= err_torch + 1e-5 * scale, (err_x3, err_torch)
    assert (q_tc - q_ref).abs().max().item() <= 3e-3 * scale, (q_tc - q_ref).abs().max().item()
    assert torch.equal(q_tc, q_tc_target)
    # and the greedy action agrees wherever the fp32 margin is not a near-tie
    top2 = q_ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-2 * scale
    assert torch.equal(q_tc.argmax(1)[clear], q_ref.argmax(1)[clear])


@pytest.mark.parametrize("env_id", ["BoatRace-v0", "SideEffectsSokoban-v0", "TomatoWatering-v0"])
@pytest.mark.parametrize("batch", [64, 1000, 20000])
def test_tcgen05_backward_gradients_match_autograd(env_id, batch):
    """Backward on the tensor cores (error chain + sample-reduction weight
    gradients, TF32 operands) against torch autograd in fp32 and against the
    fp32 FFMA backward.  TF32 keeps 10 mantissa bits and the gradients are sums
    of cancelling terms, so the stated tolerance per parameter tensor is a
    relative L2 error of 2e-2 and 5e-2 of the largest |gradient| pointwise
    (measured: 3e-4 .. 1e-2, scripts/tc_grad_diag.py); overall cosine >= 0.9999."""
    import gridfast
    torch.manual_seed(batch)
    dev = torch.device("cuda", 0)
    env = gridfast.BatchedEnv(env_id, 4, seed=1)
    rs = np.random.RandomState(batch)
    s = torch.as_tensor(rs.randint(0, 6, size=(batch, env.hw)).astype(np.uint8)).to(dev)
    s2 = torch.as_tensor(rs.randint(0, 6, size=(batch, env.hw)).astype(np.uint8)).to(dev)
    a = torch.as_tensor(rs.randint(0, 4, size=batch).astype(np.uint8)).to(dev)
    r = torch.as_tensor(rs.choice([-1.0, 2.0, 49.0], size=batch)).to(dev)
    term = torch.as_tensor((rs.rand(batch) < 0.1).astype(np.uint8)).to(dev)
    Q = build_Q(env.hw, 2, 100, 4).to(dev)
    T = build_Q(env.hw, 2, 100, 4).to(dev)
    Qs = Q(s.float()).gather(1, a.long().reshape(-1, 1)).reshape(-1)
    nxt = T(s2.float()).max(1)[0]
    nxt[term.bool()] = 0
    loss = F.mse_loss(Qs, 0.99 * nxt + r.float())
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for m in Q.modules() if isinstance(m, nn.Linear) for p in (m.weight, m.bias)])
    grads = {}
    for use_tc in (False, True):
        agent = gridfast.BatchedDeepQ(env, batch_size=batch, reference_bxb_loss=False)
        agent.load_torch_module(Q, 0)
        agent.load_torch_module(T, 1)
        agent.set_tensor_cores(3 if use_tc else 0)
        agent.learn_batch(s, a, r, s2, term)
        grads[use_tc] = agent.get_grads()
    assert torch.allclose(grads[False], ref, rtol=1e-4, atol=1e-6 * ref.abs().max().item() + 1e-7)
    off = 0
    for m in Q.modules():
        if isinstance(m, nn.Linear):
            for prm in (m.weight, m.bias):
                n = prm.numel()
                g_ref, g_tc = ref[off:off + n], grads[True][off:off + n]
                assert ((g_tc - g_ref).norm() / g_ref.norm()).item() <= 2e-2, (env_id, batch, off)
                assert (g_tc - g_ref).abs().max().item() <= 5e-2 * g_ref.abs().max().item() + 1e-7, (env_id, batch, off)
                off += n
    cos = torch.dot(grads[True], ref) / (grads[True].norm() * ref.norm())
    assert cos.item() >= 0.9999


def test_tcgen05_learn_step_tracks_torch():
    import gridfast
    torch.manual_seed(5)
    dev = torch.device("cuda", 0)
    batch = 4096
    env = gridfast.BatchedEnv("SideEffectsSokoban-v0", 8, seed=1)
    agent = gridfast.BatchedDeepQ(env, batch_size=batch, lr=1e-3, reference_bxb_loss=False)
    agent.set_tensor_cores(True)
    Q = build_Q(env.hw, 2, 100, 4).to(dev)
    T = build_Q(env.hw, 2, 100, 4).to(dev)
    optim = torch.optim.Adam(Q.parameters(), lr=1e-3, amsgrad=True)
    agent.load_torch_module(Q, 0)
    agent.load_torch_module(T, 1)
    p0 = flat(Q).clone()
    rs = np.random.RandomState(4)
    for step in range(5):
        s = torch.as_tensor(rs.randint(0, 6, size=(batch, env.hw)).astype(np.uint8)).to(dev)
        s2 = torch.as_tensor(rs.randint(0, 6, size=(batch, env.hw)).astype(np.uint8)).to(dev)
        a = torch.as_tensor(rs.randint(0, 4, size=batch).astype(np.uint8)).to(dev)
        r = torch.as_tensor(rs.choice([-1.0, 49.0, -11.0], size=batch)).to(dev)
        term = torch.as_tensor((rs.rand(batch) < 0.1).astype(np.uint8)).to(dev)
        loss, norm = reference_learn(Q, T, optim, s.float(), a, r.float(), s2.float(), term, 0.99, False)
        ours = agent.learn_batch(s, a, r, s2, term).tolist()
        assert ours[0] == pytest.approx(loss, rel=5e-3)
        assert ours[1] == pytest.approx(norm, rel=5e-3)
    du, dr = agent.get_params(0) - p0, flat(Q) - p0
    cos = torch.dot(du, dr) / (du.norm() * dr.norm())
    assert cos.item() > 0.99, cos.item()


def test_dqn_rollout_learns_with_tensor_cores():
    import gridfast
    n = 512
    env = gridfast.BatchedEnv("SideEffectsSokoban-v0", n, seed=2)
    agent = gridfast.BatchedDeepQ(env, replay_capacity=n * 40, batch_size=1024, lr=1e-3, epsilon=0.05,
                                  epsilon_anneal=150, sync_every=25, reference_bxb_loss=False, seed=11)
    agent.set_tensor_cores(True)
    random_return, tail_return = _train_and_score(agent, env)
    assert tail_return > random_return + 5, (random_return, tail_return)
