import json, os, sys
sys.path.insert(0, "/root/repo/safe-grid-agents_b200")
import torch, gridfast
HP = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=100000)
for env_id, T in (("TomatoWatering-v0", 300), ("BoatRace-v0", 1000), ("SideEffectsSokoban-v0", 1000)):
    for n in (1024, 8192, 32768, 65536, 131072):
        env = gridfast.BatchedEnv(env_id, n, seed=0, device=0)
        agent = gridfast.BatchedTabularQ(env, gridfast.Q_SHARED, **HP)
        for _ in range(2): agent.rollout(T)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): agent.rollout(T)
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 3 / T * 1e3
        print(json.dumps({"env": env_id, "n": n, "us_per_lockstep": us, "env_steps_per_s": n / us * 1e6}), flush=True)
        del agent, env
