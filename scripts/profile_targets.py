#!/usr/bin/env python
"""Small drivers for ncu captures of single kernels (see profiles/).

    ncu --set full -k regex:k_env_step ... python scripts/profile_targets.py step
    ncu --set full -k regex:k_mlp_forward_tc ... python scripts/profile_targets.py mlp
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "safe-grid-agents_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import gridfast  # noqa: E402

what = sys.argv[1]
if what == "step":
    n = 1 << 24
    env = gridfast.BatchedEnv("BoatRace-v0", n, seed=0)
    acts = torch.randint(0, 4, (n,), dtype=torch.uint8, device="cuda")
    out = (env._u8(n, env.hw), env._f64(n), env._f64(n), env._u8(n))
    for t in range(4):
        env.step(acts, step=t, out=out)
elif what == "mlp":
    env = gridfast.BatchedEnv("SideEffectsSokoban-v0", 4, seed=0)
    agent = gridfast.BatchedDeepQ(env)
    agent.set_tensor_cores(True)
    boards = torch.randint(0, 6, (1 << 20, env.hw), dtype=torch.uint8, device="cuda")
    for _ in range(4):
        agent.q_values(boards)
elif what == "learn":
    env = gridfast.BatchedEnv("SideEffectsSokoban-v0", 4096, seed=0)
    agent = gridfast.BatchedDeepQ(env, replay_capacity=100 * 4096, batch_size=262144)
    agent.set_tensor_cores(True)
    agent.warmup(100)
    agent.rollout(2)
elif what == "shared":
    env = gridfast.BatchedEnv("BoatRace-v0", 65536, seed=0)
    agent = gridfast.BatchedTabularQ(env, gridfast.Q_SHARED)
    for _ in range(3):
        agent.rollout(200)
elif what == "rollout":
    # the bench.py workload: boat, 65,536 envs, 10,000 lock-steps per launch
    env = gridfast.BatchedEnv("BoatRace-v0", 65536, seed=0)
    agent = gridfast.BatchedTabularQ(env, gridfast.Q_PRIVATE)
    for _ in range(3):
        agent.rollout(10000)
    agent.check()
elif what == "sokoban":
    # config 3, one GPU's share: 131,072 envs, hashed private tables (capacity 128)
    env = gridfast.BatchedEnv("SideEffectsSokoban-v0", 131072, seed=0)
    agent = gridfast.BatchedTabularQ(env, gridfast.Q_PRIVATE)
    for _ in range(3):
        agent.rollout(5000)
    agent.check()
elif what == "tomato_shared":
    env = gridfast.BatchedEnv("TomatoWatering-v0", 65536, seed=0)
    agent = gridfast.BatchedTabularQ(env, gridfast.Q_SHARED)
    for _ in range(3):
        agent.rollout(300)
    agent.check()
elif what == "tomato":
    # C4 shape at a quarter of the environments (tables 5 GB instead of 21 GB: ncu replays)
    env = gridfast.BatchedEnv("TomatoWatering-v0", 16384, seed=0)
    agent = gridfast.BatchedTabularQ(env, gridfast.Q_PRIVATE, capacity=8192)
    for _ in range(3):
        agent.rollout(500)
    agent.check()
torch.cuda.synchronize()
