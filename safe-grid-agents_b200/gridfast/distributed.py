"""Multi-GPU plumbing: one process per GPU, environments sharded by global id.

The rollout path shards with NO data-path collective:
rank r owns global environment ids [r*n_local, (r+1)*n_local) and the random
streams are keyed by global id, so a trajectory does not depend on the world
size.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only

  * to merge the 9 episode-statistics totals (one all-gather, folded locally in
    rank order: sums and maxima from the same collective), and
  * in shared-Q mode, at sync intervals, to merge the replicas' Q changes so
    that every GPU's table stays bit-identical: ONE all-reduce (sum) of a
    dense delta-Q array where the level's keys have a small canonical index
    (boat, sokoban, lava, island, supervisor, whisky), an all-gather of
    (key, delta-Q) records for hashed tomato tables.

The table operations are reached through a small duck-typed interface
(delta_export / restore_base / delta_apply / rebase) so the orchestration can
be exercised on CPU with a stand-in table (tests/test_distributed_cpu.py).
"""
import torch
import torch.distributed as dist

TOTAL_KEYS = ("episodes", "sum_return", "sum_performance", "sum_margin_pos", "n_margin_pos",
              "max_return", "running_return", "max_performance", "max_margin")
_MAX_SLOTS = (5, 7, 8)   # maxima; the rest are sums


def shard(n_global, rank, world):
    """Contiguous split of the global id range: (env_id0, n_local)."""
    base, extra = divmod(n_global, world)
    n_local = base + (1 if rank < extra else 0)
    env_id0 = rank * base + min(rank, extra)
    return env_id0, n_local


def all_reduce_totals(totals, group=None):
    """totals: float64 tensor [..., 9] (sgk_env_totals layout).  Sums everywhere
    except the maxima slots; never-finished ranks carry -inf there.  ONE
    collective: the rows of all ranks are gathered and folded locally in rank
    order, so sums and maxima come out of the same exchange and every rank
    computes bit-identical results."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return totals
    world = dist.get_world_size(group)
    parts = [torch.empty_like(totals) for _ in range(world)]
    dist.all_gather(parts, totals.contiguous(), group=group)
    gathered = torch.stack(parts)
    folded = gathered[0].clone()
    for g in range(1, world):
        folded += gathered[g]
    idx = torch.tensor(_MAX_SLOTS, device=totals.device)
    folded[..., idx] = gathered[..., idx].max(dim=0).values
    totals.copy_(folded)
    return totals


def sync_shared_table(table, group=None):
    """Merge the replicas of a shared Q table.

    new = base + (1/G) * sum_g (replica_g - base), applied in rank order on
    every rank, so all replicas end bit-identical (periodic averaging of the
    replicas' changes since the last sync)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dense = getattr(table, "dense_size", lambda: 0)()
    if dense > 0:
        # canonical dense index: one all-reduce (sum) of [dense][delta-Q x 4, presence]
        buf = table.delta_export_dense()
        if world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        table.delta_apply_dense(buf, 1.0 / world)
        return
    keys, delta = table.delta_export()
    if world == 1:
        gathered_k, gathered_d = [keys], [delta]
    else:
        gathered_k = [torch.empty_like(keys) for _ in range(world)]
        gathered_d = [torch.empty_like(delta) for _ in range(world)]
        dist.all_gather(gathered_k, keys, group=group)
        dist.all_gather(gathered_d, delta, group=group)
    table.restore_base()
    for g in range(world):
        table.delta_apply(gathered_k[g], gathered_d[g], 1.0 / world)
    table.rebase()


class ShardedRollout:
    """The fused rollout over a global set of environments split across the
    ranks of the default process group."""

    def __init__(self, env_id, n_global, q_mode, seed=0, device=None, sync_interval=1000,
                 capacity=0, **hyper):
        import gridfast   # CUDA path; raises loudly without the library / a GPU
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.env_id0, self.n_local = shard(n_global, self.rank, self.world)
        device = torch.cuda.current_device() if device is None else device
        self.env = gridfast.BatchedEnv(env_id, self.n_local, seed=seed, env_id0=self.env_id0, device=device)
        self.agent = gridfast.BatchedTabularQ(self.env, q_mode, capacity=capacity, **hyper)
        self.shared = q_mode == gridfast.Q_SHARED
        self.sync_interval = sync_interval
        self._totals = torch.zeros(9, dtype=torch.float64, device=self.env.device)

    def rollout(self, n_steps, cheat=False):
        done = 0
        while done < n_steps:
            chunk = min(self.sync_interval, n_steps - done) if self.shared and self.world > 1 else n_steps - done
            self.agent.rollout(chunk, cheat=cheat)
            if self.shared and self.world > 1:
                sync_shared_table(self.agent)
            done += chunk

    def totals(self):
        self.env.totals_device(self._totals)
        all_reduce_totals(self._totals)
        return dict(zip(TOTAL_KEYS, self._totals.tolist()))
