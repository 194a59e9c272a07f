#!/usr/bin/env python
"""Does the L2 fetch granularity (cudaLimitMaxL2FetchGranularity: 32 / 64 / 128 B) matter for the
hashed private tables?  Tomato tables are 10 GB of random 32-byte sectors; with the default
granularity every L2 miss may fetch a 64-byte pair from DRAM.

    python scripts/l2_granularity_probe.py [32|64|128]
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "safe-grid-agents_b200"))
import torch

import gridfast

want = int(sys.argv[1]) if len(sys.argv) > 1 else 0
torch.zeros(1, device="cuda")
rt = ctypes.CDLL([p for p in (os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart.so.12"), "libcudart.so.12", "libcudart.so") if os.path.exists(p) or "/" not in p][0])
LIMIT = 5    # cudaLimitMaxL2FetchGranularity
val = ctypes.c_size_t(0)
rt.cudaDeviceGetLimit(ctypes.byref(val), LIMIT)
before = val.value
rc = rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(want)) if want else 0
rt.cudaDeviceGetLimit(ctypes.byref(val), LIMIT)
HP = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=100000)
out = {"limit_before": before, "set_rc": rc, "limit_now": val.value}
for env_id, n, T, cap in (("TomatoWatering-v0", 65536, 1000, 8192), ("SideEffectsSokoban-v0", 131072, 5000, 0), ("BoatRace-v0", 65536, 10000, 0)):
    env = gridfast.BatchedEnv(env_id, n, seed=0, device=0)
    agent = gridfast.BatchedTabularQ(env, gridfast.Q_PRIVATE, capacity=cap, **HP)
    agent.rollout(T)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        agent.rollout(T)
    b.record()
    torch.cuda.synchronize()
    out[env_id] = n * T * 3 / (a.elapsed_time(b) * 1e-3)
    del agent, env
    torch.cuda.empty_cache()
print(json.dumps(out))
