#!/usr/bin/env python
"""Supplementary measurements (not the bench.py contract): every in-scope
workload of BASELINE.json configs 2-4 plus the HBM-honest data points.
Prints one JSON line per measurement; run on one B200:

    python scripts/bench_suite.py > gpurun_out/suite.jsonl
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "safe-grid-agents_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import gridfast  # noqa: E402

HP = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=100000)
# SURVEY 8(d) contract: 2 x board + 2 x scalars + 52 (Q traffic) + 10 (outputs)
B_ALG = {"BoatRace-v0": 132, "SideEffectsSokoban-v0": 158, "TomatoWatering-v0": 208,
         "DistributionalShift-v0": 208, "IslandNavigation-v0": 178, "AbsentSupervisor-v0": 180,
         "WhiskyGold-v0": 180}
BASE_ENVS = ("BoatRace-v0", "SideEffectsSokoban-v0", "TomatoWatering-v0")
WIDENED_ENVS = ("DistributionalShift-v0", "IslandNavigation-v0", "AbsentSupervisor-v0", "WhiskyGold-v0")
PEAK = 6550.7
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    FLUSH.zero_()


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(reps):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / reps / 1e3


def emit(**kw):
    print(json.dumps(kw), flush=True)


def fused(env_id, n, T, q_mode, reps=5, capacity=0, ssrl=False, label=None, warm=2):
    env = gridfast.BatchedEnv(env_id, n, seed=0)
    agent = gridfast.BatchedTabularQ(env, q_mode, capacity=capacity, **HP)
    if ssrl:
        agent.enable_ssrl(c_prior=0.01, budget=1000)
    sec = timed(lambda: agent.rollout(T), reps, warm)
    agent.check()
    rate = n * T / sec
    tot = env.totals()
    emit(measurement=label or "fused_rollout", env=env_id, n_envs=n, locksteps=T,
         q_mode="private" if q_mode == gridfast.Q_PRIVATE else "shared", ssrl=ssrl,
         capacity=agent.capacity, seconds_per_call=sec, env_steps_per_s=rate,
         us_per_lockstep=1e6 * sec / T,
         algorithmic_GBps=rate * B_ALG[env_id] / 1e9, frac_of_measured_hbm=rate * B_ALG[env_id] / 1e9 / PEAK,
         episodes=tot["episodes"], mean_return=tot["sum_return"] / max(tot["episodes"], 1),
         mean_performance=tot["sum_performance"] / max(tot["episodes"], 1))
    del agent, env
    torch.cuda.empty_cache()


def unfused_step(env_id, n, reps=10):
    """One lock-step per launch, state streamed from/to HBM: the honest
    HBM-bound form of env.step (inputs >> L2 at n = 2^24)."""
    env = gridfast.BatchedEnv(env_id, n, seed=0)
    acts = torch.randint(0, 4, (n,), dtype=torch.uint8, device="cuda")
    out = (env._u8(n, env.hw), env._f64(n), env._f64(n), env._u8(n))
    state = {"t": 0}

    def one():
        env.step(acts, step=state["t"], out=out)
        state["t"] += 1
    sec = timed(one, reps)
    # bytes actually moved per env-step: core+return+hidden r/w, action, board, reward, hidden, done
    moved = 8 * 2 * 3 + 1 + env.hw + 8 + 8 + 1
    emit(measurement="unfused_env_step", env=env_id, n_envs=n, seconds_per_call=sec,
         env_steps_per_s=n / sec, bytes_moved_per_env_step=moved,
         achieved_GBps=n * moved / sec / 1e9, frac_of_measured_hbm=n * moved / sec / 1e9 / PEAK)
    del env
    torch.cuda.empty_cache()


def unfused_agent(env_id, n, reps=10):
    """act + learn as separate launches over boards in HBM (private tables)."""
    env = gridfast.BatchedEnv(env_id, n, seed=0)
    agent = gridfast.BatchedTabularQ(env, gridfast.Q_PRIVATE, **HP)
    s = env.render()
    acts = torch.randint(0, 4, (n,), dtype=torch.uint8, device="cuda")
    s2, r, h, d = env.step(acts, step=0)
    a_out = env._u8(n)

    def one():
        agent.act(s, 5, explore=True, out=a_out)
        agent.learn(s, acts, r, s2)
    sec = timed(one, reps)
    moved = env.hw * 3 + 1 + 1 + 8 + (8 + 32) * 2 + 8 + 32
    emit(measurement="unfused_act_plus_learn", env=env_id, n_envs=n, seconds_per_call=sec,
         env_steps_per_s=n / sec, bytes_moved_per_env_step=moved,
         achieved_GBps=n * moved / sec / 1e9, frac_of_measured_hbm=n * moved / sec / 1e9 / PEAK)
    del agent, env
    torch.cuda.empty_cache()


def dqn(n, batch, T, use_tc, reps=3, label=None):
    """C5: side-effects sokoban, one shared 36-100-100-4 Q network, HBM replay ring."""
    env = gridfast.BatchedEnv("SideEffectsSokoban-v0", n, seed=0)
    agent = gridfast.BatchedDeepQ(env, replay_capacity=100 * n, batch_size=batch, lr=1e-3, epsilon=0.01,
                                  epsilon_anneal=100000, sync_every=10000, reference_bxb_loss=True)
    agent.set_tensor_cores(use_tc)
    agent.warmup(100)
    sec = timed(lambda: agent.rollout(T), reps, warm=1)
    flop = T * (n * 28000.0 + batch * 4 * 28000.0)
    emit(measurement=label or "dqn_rollout", env="SideEffectsSokoban-v0", n_envs=n, learn_batch=batch, locksteps=T,
         forward=agent.precision, seconds_per_call=sec, env_steps_per_s=n * T / sec,
         us_per_lockstep=1e6 * sec / T, samples_learned_per_s=batch * T / sec, model_TFLOPs=flop / sec / 1e12,
         loss_norm_clip=agent.last_scalars())
    del agent, env
    torch.cuda.empty_cache()


def mlp_forward(rows, use_tc, reps=10):
    """The forward kernel alone: rows x (36-100-100-4), 28.0 kFLOP per row."""
    env = gridfast.BatchedEnv("SideEffectsSokoban-v0", 4, seed=0)
    agent = gridfast.BatchedDeepQ(env)
    agent.set_tensor_cores(use_tc)
    boards = torch.randint(0, 6, (rows, env.hw), dtype=torch.uint8, device="cuda")
    sec = timed(lambda: agent.q_values(boards), reps)
    emit(measurement="mlp_forward", rows=rows, forward=agent.precision, seconds_per_call=sec,
         rows_per_s=rows / sec, TFLOPs=rows * 28000.0 / sec / 1e12,
         frac_of_measured_bf16_tensor_peak=rows * 28000.0 / sec / 1e12 / 1648.6 if use_tc else None)
    del agent, env
    torch.cuda.empty_cache()


def torch_context(rows=1 << 20, batch=262144, reps=10):
    """Context for the deep-Q kernels: the same 36-100-100-4 network in plain
    torch -- cuBLAS GEMMs with TF32 allowed, separate bias / ReLU kernels, autograd backward -- on the
    same shapes.  Not part of the product path."""
    import torch.nn as nn
    for allow in (True, False):
        torch.backends.cuda.matmul.allow_tf32 = allow
        torch.backends.cudnn.allow_tf32 = allow
        net = nn.Sequential(nn.Linear(36, 100), nn.ReLU(), nn.Linear(100, 100), nn.ReLU(), nn.Linear(100, 4)).cuda()
        x = torch.randint(0, 6, (rows, 36), device="cuda").float()
        with torch.no_grad():
            sec = timed(lambda: net(x), reps)
        emit(measurement="torch_mlp_forward", rows=rows, math="cuBLAS TF32" if allow else "cuBLAS fp32", seconds_per_call=sec,
             rows_per_s=rows / sec, TFLOPs=rows * 28000.0 / sec / 1e12)
        xb = torch.randint(0, 6, (batch, 36), device="cuda").float()
        a = torch.randint(0, 4, (batch, 1), device="cuda")
        y = torch.randn(batch, device="cuda")

        def step():
            net.zero_grad(set_to_none=True)
            q = net(xb).gather(1, a).reshape(-1)
            torch.nn.functional.mse_loss(q, y).backward()
        sec = timed(step, reps)
        emit(measurement="torch_mlp_forward_backward", batch=batch, math="cuBLAS TF32" if allow else "cuBLAS fp32",
             seconds_per_call=sec, samples_per_s=batch / sec, TFLOPs=batch * 3 * 28000.0 / sec / 1e12)
    torch.backends.cuda.matmul.allow_tf32 = False


def main():
    P, S = gridfast.Q_PRIVATE, gridfast.Q_SHARED
    which = sys.argv[1:] or ["fused", "tomato", "large", "unfused"]
    if "fused" in which:
        fused("BoatRace-v0", 65536, 10000, P, label="C2 boat private")
        fused("BoatRace-v0", 65536, 2000, S, label="C2 boat shared")
        fused("BoatRace-v0", 262144, 1000, S, label="boat shared 262144")
        fused("SideEffectsSokoban-v0", 131072, 5000, P, label="C3 sokoban private (per-GPU share)")
        fused("SideEffectsSokoban-v0", 131072, 1000, S, label="C3 sokoban shared")
    if "tomato" in which:
        # private tomato tables grow with every distinct board an agent sees:
        # 4 x 1000 lock-steps need <= 4000 slots -> capacity 8192 (21 GB of tables)
        fused("TomatoWatering-v0", 65536, 1000, P, capacity=8192, reps=3, warm=1, label="C4 tomato private")
        fused("TomatoWatering-v0", 65536, 1000, P, capacity=8192, reps=3, warm=1, ssrl=True, label="C4 tomato private + SSRL")
        fused("TomatoWatering-v0", 65536, 1000, S, label="C4 tomato shared")
    if "widened" in which:
        for env_id in WIDENED_ENVS:
            fused(env_id, 65536, 5000, P, label="8(f3) private")
            fused(env_id, 65536, 1000, S, label="8(f3) shared")
    if "large" in which:
        fused("BoatRace-v0", 1 << 24, 200, P, reps=3, label="large-N boat private 2^24 x 200")
        fused("SideEffectsSokoban-v0", 1 << 22, 200, P, reps=3, label="large-N sokoban private 2^22 x 200")
    if "torch" in which:
        torch_context()
    if "dqn" in which:
        for use_tc in (False, True):
            mlp_forward(1 << 20, use_tc)
        for use_tc in (False, True):
            dqn(4096, 4096, 200, use_tc, label="C5 sokoban DQN 4096 envs, batch 4096")
            dqn(4096, 64 * 4096, 50, use_tc, label="C5 sokoban DQN 4096 envs, batch 64 per env-step")
    if "unfused" in which:
        for env_id in BASE_ENVS:
            unfused_step(env_id, 1 << 24)
        unfused_agent("BoatRace-v0", 1 << 24)


if __name__ == "__main__":
    main()
