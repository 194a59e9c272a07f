"""Wall/GPU time of the deep-Q lock-step at a given learn batch (CUDA events)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "safe-grid-agents_b200")):
    sys.path.insert(0, p)
import gridfast
for batch in [int(x) for x in sys.argv[1:]] or [64, 4096]:
    env = gridfast.BatchedEnv("SideEffectsSokoban-v0", 4096, seed=0)
    agent = gridfast.BatchedDeepQ(env, replay_capacity=100 * 4096, batch_size=batch)
    agent.set_tensor_cores(True)
    agent.warmup(100)
    agent.rollout(50)
    torch.cuda.synchronize()
    T = 400
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    agent.rollout(T)
    t_launch = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    print("batch %d: %.1f us/lock-step on the GPU, host enqueue %.1f us/lock-step" %
          (batch, 1e3 * e0.elapsed_time(e1) / T, 1e6 * t_launch / T), flush=True)
