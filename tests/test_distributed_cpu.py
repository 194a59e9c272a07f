"""Host-side multi-GPU logic on CPU: world_size 2 over gloo.  The CUDA table
is replaced by a dict-backed stand-in with the same four operations; what is
tested is the sharding arithmetic, the statistics all-reduce and the replica
sync protocol (rank-ordered, 1/G-scaled, identical result on every rank)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gridfast import distributed as gd


class FakeTable:
    """Fixed-capacity (keys, rows) arrays like sgk_tabq_export, dict inside."""

    def __init__(self, capacity=16):
        self.capacity = capacity
        self.live, self.base = {}, {}

    def delta_export(self):
        keys = torch.zeros(self.capacity, dtype=torch.int64)
        delta = torch.zeros(self.capacity, 4, dtype=torch.float64)
        for i, (k, row) in enumerate(sorted(self.live.items())):
            keys[i] = k
            delta[i] = torch.tensor(row - self.base.get(k, np.zeros(4)))
        return keys, delta

    def restore_base(self):
        self.live = {k: v.copy() for k, v in self.base.items()}

    def delta_apply(self, keys, delta, scale):
        for k, d in zip(keys.tolist(), delta.numpy()):
            if k:
                self.live[k] = self.live.get(k, np.zeros(4)) + scale * d

    def rebase(self):
        self.base = {k: v.copy() for k, v in self.live.items()}


class FakeDenseTable(FakeTable):
    """... with the canonical-index protocol (sgk_tabq_delta_export_dense /
    _apply_dense): keys ARE their dense index here."""

    def dense_size(self):
        return self.capacity

    def delta_export_dense(self):
        out = torch.zeros(self.capacity, 5, dtype=torch.float64)
        for k, row in self.live.items():
            out[k, :4] = torch.tensor(row - self.base.get(k, np.zeros(4)))
            out[k, 4] = 1.0
        return out

    def delta_apply_dense(self, delta_sum, scale):
        self.restore_base()
        for k in range(self.capacity):
            if delta_sum[k, 4] > 0:
                self.live[k] = self.live.get(k, np.zeros(4)) + scale * delta_sum[k, :4].numpy()
        self.rebase()


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # statistics: sums everywhere, max in slot 5
        totals = torch.tensor([10.0 + rank, -50.0 * (rank + 1), 3.0, 1.5, 2.0,
                               -7.0 if rank == 0 else -3.0, 4.0, 5.0 - rank, float("-inf") if rank else 2.5],
                              dtype=torch.float64)
        gd.all_reduce_totals(totals)
        # replica sync, two rounds
        table = FakeTable()
        table.live = {101: np.array([1.0, 0, 0, 0]) * (rank + 1), 200 + rank: np.full(4, 2.0)}
        gd.sync_shared_table(table)
        first = {k: v.copy() for k, v in table.live.items()}
        table.live[101] = table.live[101] + np.array([0, 4.0 * rank, 0, 0])
        gd.sync_shared_table(table)
        # the same two rounds through the dense one-all-reduce protocol
        dense = FakeDenseTable()
        dense.live = {1: np.array([1.0, 0, 0, 0]) * (rank + 1), 4 + rank: np.full(4, 2.0)}
        gd.sync_shared_table(dense)
        dense.live[1] = dense.live[1] + np.array([0, 4.0 * rank, 0, 0])
        gd.sync_shared_table(dense)
        out.put((rank, totals.tolist(), first, table.live, dense.live))
    finally:
        dist.destroy_process_group()


def test_two_rank_statistics_and_replica_sync():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, t0, first0, final0, dense0), (_, t1, first1, final1, dense1) = results
    for dense in (dense0, dense1):
        assert set(dense) == {1, 4, 5}
        assert np.array_equal(dense[1], [1.5, 2.0, 0, 0]) and np.array_equal(dense[4], np.full(4, 1.0))
    assert t0 == t1 == [21.0, -150.0, 6.0, 3.0, 4.0, -3.0, 8.0, 5.0, 2.5]
    # round 1: base empty -> mean of the replicas' values, keys unioned
    for first in (first0, first1):
        assert set(first) == {101, 200, 201}
        assert np.array_equal(first[101], [1.5, 0, 0, 0])
        assert np.array_equal(first[200], np.full(4, 1.0)) and np.array_equal(first[201], np.full(4, 1.0))
    # round 2: only rank 1 changed key 101 by +4 in column 1 -> +2 everywhere
    for final in (final0, final1):
        assert np.array_equal(final[101], [1.5, 2.0, 0, 0])
        assert np.array_equal(final[200], np.full(4, 1.0))
    assert all(np.array_equal(final0[k], final1[k]) for k in final0)


def test_shard_covers_the_global_range_contiguously():
    for n_global, world in ((1048576, 8), (65536, 1), (10, 4), (7, 8)):
        cursor = 0
        for rank in range(world):
            start, n = gd.shard(n_global, rank, world)
            assert start == cursor and n >= 0
            cursor += n
        assert cursor == n_global
