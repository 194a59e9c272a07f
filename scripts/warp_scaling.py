#!/usr/bin/env python
"""How the fused boat rollout scales with resident warps per scheduler.

n environments = 148 SMs x 4 schedulers x 32 lanes x w for w = 1..4 (plus the
bench's 65,536): if the time per launch stays flat as w grows, the kernel is
bound by the dependent chain of one lock-step (latency); if it grows with w,
by a shared resource (issue slots or a pipe).  Prints one JSON line per size.

    python scripts/warp_scaling.py [lock-steps per launch]
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "safe-grid-agents_b200"))
import torch

import gridfast

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
HP = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=100000)
for n in (148 * 128 * 1, 148 * 128 * 2, 148 * 128 * 3, 148 * 128 * 4, 65536):
    env = gridfast.BatchedEnv("BoatRace-v0", n, seed=0, device=0)
    agent = gridfast.BatchedTabularQ(env, gridfast.Q_PRIVATE, **HP)
    for _ in range(12):                      # past the annealing phase
        agent.rollout(T)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        agent.rollout(T)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(json.dumps({"n_envs": n, "warps_per_scheduler": n / (148 * 128), "ms_per_launch": ms,
                      "cycles_per_lockstep_at_1965MHz": ms * 1e-3 * 1.965e9 / T, "env_steps_per_s": n * T / (ms * 1e-3)}))
    del agent, env
