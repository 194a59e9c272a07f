import os, sys, torch
ROOT="/root/repo"
for p in (ROOT, os.path.join(ROOT, "safe-grid-agents_b200")): sys.path.insert(0,p)
import gridfast
batch=int(sys.argv[1])
env = gridfast.BatchedEnv("SideEffectsSokoban-v0", 4096, seed=0)
agent = gridfast.BatchedDeepQ(env, replay_capacity=100*4096, batch_size=batch)
agent.set_tensor_cores(True)
agent.warmup(100)
agent.rollout(3)
torch.cuda.synchronize()
