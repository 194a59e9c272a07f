import sys, argparse, cProfile, pstats, time
sys.path.insert(0, "/root/repo/safe-grid-agents_b200")
import numpy as np
import gridfast as gf
class NullWriter:
    def add_scalar(self, *a, **k): pass
    add_scalars = add_scalar
class Meter:
    val = avg = max = 0.0
    def update(self, v, n=1): self.val = v
args = argparse.Namespace(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=100000, cheat=False, eval_every=10 ** 9)
np.random.seed(0)
env = gf.make("BoatRace-v0", rng="numpy")
agent = gf.GpuTabularQAgent(env, args)
history = {"writer": NullWriter(), "t": 0, "episode": 0, "returns": Meter(), "safeties": Meter(), "margins": Meter(), "margins_support": Meter()}
def episodes(m):
    for _ in range(m):
        state = (env.reset(), 0.0, False, {})
        history["episode"] += 1
        gf.tabq_learn_fused(agent, env, state, history, args)
episodes(20)
pr = cProfile.Profile(); pr.enable(); episodes(200); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
