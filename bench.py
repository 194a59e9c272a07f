#!/usr/bin/env python
"""bench.py -- env-steps/s of the rollout hot path (step + tabular-Q update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): boat race, tabular Q-learning at the
reference's defaults (lr .5, discount .99, epsilon .01, epsilon-anneal 100000),
65,536 lock-step environments PER GPU (weak scaling), one private agent per
environment = N independent copies of the reference's (env, agent) pair.
One bench "step" = one fused rollout call of 10,000 lock-steps over the whole
batch (100 episodes per environment, auto-reset) = 655,360,000 env-steps/GPU.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "safe-grid-agents_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

ENV_ID = "BoatRace-v0"
N_ENVS_PER_GPU = 65536
LOCKSTEPS_PER_CALL = 10000
HP = dict(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=100000)
B_ALG = 132  # algorithmic bytes per env-step, boat race (SURVEY.md 8d / DESIGN.md)
WORKLOAD = "boat-race tabular-Q, 65536 lock-step envs/GPU, private Q, 10000 lock-steps per call"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu
    capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "rollout_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


# ----------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region
    runs (NVML; falls back to nvidia-smi)."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        if self.nv is None:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM))
            try:
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            for bit, name in {**self.BAD, **self.NOTE}.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        # NVML queries take a driver-wide lock: with one sampler per rank, poll
        # gently (and not at all when SGK_BENCH_NO_CLOCKS is set)
        if os.environ.get("SGK_BENCH_NO_CLOCKS"):
            return
        while not self._stop.is_set():
            self._once()
            self._stop.wait(0.02 if int(os.environ.get("WORLD_SIZE", "1")) == 1 else 0.05)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------- CPU port
def _cpu_worker(args):
    """One process: the reference's loop shape on the restated CPU path
    (oracle env + restated TabularQAgent), boat race, `n_steps` steps."""
    seed, n_steps = args
    import numpy as np
    from oracle import gridworld_env, tabular
    np.random.seed(seed)
    env = gridworld_env.make(ENV_ID)
    env.seed(seed)
    agent = tabular.TabularQAgent(4, HP["discount"], HP["epsilon"], HP["epsilon_anneal"], HP["lr"])
    t = time.perf_counter()
    tabular.run_tabq(agent, env, n_steps)
    return n_steps, time.perf_counter() - t


def cpu_port_rate(n_procs, n_steps):
    """Aggregate env-steps/s of `n_procs` independent (env, agent) pairs."""
    if n_procs == 1:
        n, dt = _cpu_worker((0, n_steps))
        return n / dt
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(n_procs) as pool:
        pool.map(_cpu_worker, [(s, 200) for s in range(n_procs)])      # warm imports
        t = time.perf_counter()
        res = pool.map(_cpu_worker, [(100 + s, n_steps) for s in range(n_procs)])
        wall = time.perf_counter() - t
    return sum(r[0] for r in res) / wall


def c_port_rate():
    """The plain-C restatement (oracle/cgrid.c, pthreads over environments) on
    all host cores: a stronger CPU yardstick than the Python port, reported
    beside it.  None if the oracle library is not built."""
    try:
        from oracle import cgrid
        n, T = 65536, 100
        sim = cgrid.Sim(cgrid.BOAT, n, seed=0, lr=HP["lr"], discount=HP["discount"], epsilon=HP["epsilon"],
                        epsilon_anneal=HP["epsilon_anneal"])
        sim.rollout(10)
        t = time.perf_counter()
        sim.rollout(T)
        dt = time.perf_counter() - t
        return {"value": n * T / dt, "unit": "env-steps/s", "cores": min(os.cpu_count() or 1, 64),
                "sample": "%d envs x %d lock-steps of boat-race tabular-Q, C oracle (pthreads), %.2f s" % (n, T, dt)}
    except Exception as exc:      # the C yardstick is optional; the Python port below is the contract value
        return {"unavailable": repr(exc)[:200]}


def run_reference(args, out):
    """--impl reference: the CPU implementation of the path on all host cores.
    /root/reference is pure Python whose env half is an absent third-party
    dependency, so it cannot be compiled into oracle/_ref; the arm times the
    oracle port (kind "port"): Python, pycolab-style engine + the restated
    TabularQAgent, one independent (env, agent) pair per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = 20000          # env-steps per process per bench step (~1.5 s)
    t_total, n_total = 0.0, 0
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(s, 500) for s in range(cores)])
        for w in range(args.warmup):
            pool.map(_cpu_worker, [(1000 * w + s, 2000) for s in range(cores)])
        for k in range(args.steps):
            t = time.perf_counter()
            res = pool.map(_cpu_worker, [(7000 + 100 * k + s, per_step) for s in range(cores)])
            t_total += time.perf_counter() - t
            n_total += sum(r[0] for r in res)
    value = n_total / t_total
    sample = "%d procs x %d env-steps x %d steps of boat-race tabular-Q (python oracle port)" % (cores, per_step, args.steps)
    line = {
        "impl": "reference", "metric": "env-steps/sec (step + tabular-Q update)", "value": value,
        "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_sample": sample},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "c_port": c_port_rate()},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=out, flush=True)


# ----------------------------------------------------------------- ours
def run_ours(args, out):
    import torch
    import torch.distributed as dist

    import gridfast
    from gridfast.distributed import _MAX_SLOTS
    MAX_COLS = list(_MAX_SLOTS)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = N_ENVS_PER_GPU
    T = LOCKSTEPS_PER_CALL
    K, W = args.steps, max(args.warmup, 3)

    env = gridfast.BatchedEnv(ENV_ID, n, seed=0, env_id0=rank * n, device=local)
    agent = gridfast.BatchedTabularQ(env, gridfast.Q_PRIVATE, **HP)
    totals = torch.zeros(9, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # 2x the 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    # Private-Q rollouts need no data-path collective.  Episode statistics are
    # kept per bench step and all-reduced (NCCL) once per sync interval -- here
    # the K timed steps -- so ranks never wait for each other inside a step; the
    # collective is timed as the last ("drain") segment of the region.
    def one_step(buf):
        agent.rollout(T)                      # 2 launches: thresholds + fused rollout
        env.totals_device(buf)                # 2 launches: partial + final reduction

    def sync_stats(block):                   # sums, and maxima for the max_* columns
        if world > 1:
            maxima = block[:, MAX_COLS].clone()
            dist.all_reduce(block)
            dist.all_reduce(maxima, op=dist.ReduceOp.MAX)
            block[:, MAX_COLS] = maxima

    for _ in range(W):
        one_step(totals)
    sync_stats(torch.zeros(K, 9, dtype=torch.float64, device=dev))   # same collectives as the timed ones
    agent.check()
    barrier()

    # ---- device-timed region: K steps, CUDA events on the launching stream
    stats = torch.zeros(K, 9, dtype=torch.float64, device=dev)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    kstart = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    kend = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    with ClockSampler(local) as clocks:
        for k in range(K):
            flush.zero_()                     # evict L2 between timed iterations (untimed)
            starts[k].record()
            kstart[k].record()
            agent.rollout(T)
            kend[k].record()
            env.totals_device(stats[k])
            ends[k].record()
        starts[K].record()                    # sync interval ends: one all-reduce of all K rows
        sync_stats(stats)
        ends[K].record()
        barrier()
    agent.check()
    step_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    drain_ms = starts[K].elapsed_time(ends[K])
    kernel_ms_by_step = [round(s.elapsed_time(e), 4) for s, e in zip(kstart, kend)]    # this rank's
    kernel_ms = sum(s.elapsed_time(e) for s, e in zip(kstart, kend))
    t = torch.tensor([step_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kernel_ms = float(t[0]), float(t[1])
    value = world * n * T * K / (step_ms / 1e3)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory)
    core_in = torch.empty(n, dtype=torch.int64).pin_memory()
    core_out = torch.empty(n, dtype=torch.int64).pin_memory()
    boards_out = torch.empty(n, env.hw, dtype=torch.uint8).pin_memory()
    core_in.copy_(env.core().cpu())
    for _ in range(2):
        agent.rollout_host(T, core_in, core_out, boards_out)
        core_in.copy_(core_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        host_totals = agent.rollout_host(T, core_in, core_out, boards_out)
        core_in.copy_(core_out)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * T * K / float(t[0])
    h2d = n * 8
    d2h = n * env.hw + n * 8 + 9 * 8

    if rank == 0:
        peak, peak_src = measured_peak()
        per_launch_s = kernel_ms / 1e3 / K
        achieved = B_ALG * n * T / per_launch_s / 1e9
        traffic = recorded_traffic()
        line = {
            "metric": "env-steps/sec (step + tabular-Q update)", "value": value, "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": step_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "locksteps_per_call": T,
                       "q_mode": "private", "rng": "philox4x32-10", "l2": "flushed between timed iterations (256 MiB memset)",
                       "episodes_finished": float(host_totals[0])},
            "roofline": {"bound": "hbm", "kernel": "k_rollout_private<boat,philox>",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch"),
                         "algorithmic_bytes_per_env_step": B_ALG, "peak_source": peak_src,
                         "kernel_ms_per_launch": kernel_ms / K,
                         # epsilon anneals over the first 100,000 agent steps: early launches explore
                         # (environments of a warp spread over the states), later ones run greedy
                         "kernel_ms_by_launch": kernel_ms_by_step},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": 4 * K, "stats_allreduce_ms": drain_ms,
            "clocks": clocks.summary(),
        }
        if world == 1:
            line["hbm_point"] = hbm_point(torch, gridfast, peak)
            t = time.perf_counter()
            rate = cpu_port_rate(1, 150000)
            line["cpu_baseline"] = {
                "value": rate, "unit": "env-steps/s", "cores": 1, "kind": "port",
                "sample": "150000 env-steps of boat-race tabular-Q, python oracle port, 1 process (%.1f s)" % (time.perf_counter() - t),
                "c_port": c_port_rate()}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """Keep stdout for the single JSON line: libraries (NCCL prints its version
    banner to stdout) are redirected to stderr for the whole run."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def hbm_point(torch, gridfast, peak):
    """The HBM-honest data point next to the fused headline: the UNFUSED
    env.step kernel (state in HBM, one lock-step per launch, boards / reward /
    hidden / done written out) at 2^24 boat-race environments -- inputs and
    outputs far larger than L2 -- reported in bytes the kernel actually moves
    (ncu: profiles/r01_env_step_boat_2p24_ncu_full.txt)."""
    n = 1 << 24
    env = gridfast.BatchedEnv(ENV_ID, n, seed=0)
    acts = torch.randint(0, 4, (n,), dtype=torch.uint8, device=env.device)
    out = (env._u8(n, env.hw), env._f64(n), env._f64(n), env._u8(n))
    for t in range(3):
        env.step(acts, step=t, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    a.record()
    for t in range(reps):
        env.step(acts, step=3 + t, out=out)
    b.record()
    torch.cuda.synchronize()
    sec = a.elapsed_time(b) / 1e3 / reps
    moved = 8 * 2 * 3 + 1 + env.hw + 8 + 8 + 1       # core/return/hidden r+w, action, board, reward, hidden, done
    gbs = n * moved / sec / 1e9
    return {"kernel": "k_env_step<boat,philox>", "n_envs": n, "env_steps_per_s": n / sec,
            "bytes_moved_per_env_step": moved, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    out = _claim_stdout()
    if args.impl == "reference":
        run_reference(args, out)
    else:
        run_ours(args, out)


if __name__ == "__main__":
    main()
