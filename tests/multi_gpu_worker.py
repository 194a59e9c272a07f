"""Worker for tests/test_gpu_multi.py: run under torchrun with >= 2 GPUs."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "safe-grid-agents_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import gridfast
    from gridfast import distributed as gd

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {}
    hp = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=300)

    # (a) private tables: sharding must not change any trajectory
    n_global, T = 8192, 350
    run = gd.ShardedRollout("TomatoWatering-v0", n_global, gridfast.Q_PRIVATE, seed=3, **hp)
    run.env.set_trace(True)
    run.rollout(T)
    run.agent.check()
    hashes = run.env.stats()["trace_hash"]
    gathered = [torch.empty_like(hashes) for _ in range(world)]
    dist.all_gather(gathered, hashes)
    totals = run.totals()
    if rank == 0:
        from oracle import cgrid
        sim = cgrid.Sim(cgrid.TOMATO, n_global, seed=3, **hp)
        sim.rollout(T)
        ref = sim.env_stats()
        got = torch.cat(gathered).cpu().numpy().view(np.uint64)
        out["private_traces_equal_oracle"] = bool(np.array_equal(got, ref["trace_hash"]))
        out["episodes"] = [totals["episodes"], float(ref["episodes"].sum())]
        out["sum_return_close"] = bool(abs(totals["sum_return"] - ref["sum_return"].sum()) < 1e-6)
        out["max_return_equal"] = bool(totals["max_return"] == ref["max_return"].max())

    # (b) shared tables: replicas identical after every sync
    run = gd.ShardedRollout("SideEffectsSokoban-v0", 4096, gridfast.Q_SHARED, seed=5, sync_interval=40, **hp)
    run.rollout(200)
    run.agent.check()
    keys, rows = run.agent.export(0)
    order = np.argsort(keys)
    blob = torch.as_tensor(np.concatenate([keys[order].astype(np.float64), rows[order].reshape(-1)])).cuda()
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([blob.numel()], device="cuda"))
    same_size = all(int(s) == blob.numel() for s in sizes)
    blobs = [torch.empty_like(blob) for _ in range(world)]
    if same_size:
        dist.all_gather(blobs, blob)
    if rank == 0:
        out["shared_replicas_identical"] = bool(same_size and all(torch.equal(b, blobs[0]) for b in blobs))
        out["shared_states"] = int(len(keys))
        out["shared_learned"] = bool(np.abs(rows).sum() > 0)
        print("MULTI_GPU_RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
