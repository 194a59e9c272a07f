#!/usr/bin/env python
"""bench.py -- env-steps/s of the rollout hot path (step + tabular-Q update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): boat race, tabular Q-learning at the
reference's defaults (lr .5, discount .99, epsilon .01, epsilon-anneal 100000),
65,536 lock-step environments PER GPU (weak scaling), one private agent per
environment = N independent copies of the reference's (env, agent) pair.
One bench "step" = one fused rollout call of 10,000 lock-steps over the whole
batch (100 episodes per environment, auto-reset) = 655,360,000 env-steps/GPU.

The same line carries sub-records for the other BASELINE configurations
(`configs`: C1 single-env drop-in, C3 sokoban 131,072 envs/GPU = 1,048,576 at
--gpus 8, C4 tomato + SSRL with budget 1,000 / warm-up .5, C5 sokoban deep-Q)
and, for every N, a shared-table leg (`shared_q`) whose replicas are merged
with one NCCL all-reduce of dense delta-Q arrays per sync interval.
Skip them with --main-only.

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


def recorded_issue():
    """Warp-instructions per launch of the dominant kernel, by launch index of
    this very command, from the committed ncu capture (smsp__inst_executed.sum;
    profiles/rollout_issue.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "rollout_issue.json")) as f:
            return json.load(f)
    except Exception:
        return None


B_ALG_BY_ENV = {"BoatRace-v0": 132, "SideEffectsSokoban-v0": 158, "TomatoWatering-v0": 212}   # SURVEY 8(d); tomato incl. C[s]


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
    from gridfast.distributed import all_reduce_totals, sync_shared_table

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

    def max_over_ranks(*values):
        t = torch.tensor(values, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    # Private-Q rollouts need no data-path collective.  Episode statistics are
    # kept per bench step and merged (NCCL, ONE all-gather folded locally) once
    # per sync interval -- here the K timed steps -- so ranks never wait for each
    # other inside a step; the collective is timed as the last ("drain") segment.
    def one_step(buf):
        agent.rollout(T)                      # 2 launches: thresholds + fused rollout
        env.totals_device(buf)                # 2 launches: partial + final reduction

    for _ in range(W):
        one_step(totals)
    all_reduce_totals(torch.zeros(K, 9, dtype=torch.float64, device=dev))   # same collective as the timed one
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
        # sync interval ends: one collective over all K rows.  The ranks are re-aligned first
        # (untimed): each rank's clock is the sum of its own steps, so what it waited for the
        # others' UNTIMED gaps (L2 flushes, launch jitter of eight Python processes) is not
        # work of the path; the slowest rank's steps still set the result (max over ranks).
        barrier()
        starts[K].record()
        all_reduce_totals(stats)
        ends[K].record()
        barrier()
    agent.check()
    step_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    drain_ms = starts[K].elapsed_time(ends[K])
    kernel_ms_by_step = [round(s.elapsed_time(e), 4) for s, e in zip(kstart, kend)]    # this rank's
    kernel_ms = sum(s.elapsed_time(e) for s, e in zip(kstart, kend))
    step_ms, kernel_ms = max_over_ranks(step_ms, kernel_ms)
    value = world * n * T * K / (step_ms / 1e3)
    episodes_finished = float(stats[-1, 0])

    # ---- end to end through the host-buffer C-ABI call (pinned host memory):
    # every step uploads the environment state words, runs the rollout, and
    # reads back boards, state and totals; L2 is flushed before each step like
    # in the device-timed region, and only the steps themselves are timed
    core_in = torch.empty(n, dtype=torch.int64).pin_memory()
    core_out = torch.empty(n, dtype=torch.int64).pin_memory()
    boards_out = torch.empty(n, env.hw, dtype=torch.uint8).pin_memory()
    core_in.copy_(env.core().cpu())
    for _ in range(2):
        agent.rollout_host(T, core_in, core_out, boards_out)
        core_in.copy_(core_out)
    e2e_s = 0.0
    for _ in range(K):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        host_totals = agent.rollout_host(T, core_in, core_out, boards_out)      # synchronises
        e2e_s += time.perf_counter() - t0
        core_in.copy_(core_out)
    (e2e_s,) = max_over_ranks(e2e_s)
    e2e_value = world * n * T * K / e2e_s
    h2d = n * 8
    d2h = n * env.hw + n * 8 + 9 * 8
    del host_totals

    # ---- the other BASELINE configurations and the shared-table leg
    extra = {}
    if not args.main_only:
        del agent, env
        torch.cuda.empty_cache()
        ctx = dict(torch=torch, dist=dist, gridfast=gridfast, world=world, rank=rank, local=local, dev=dev, flush=flush,
                   barrier=barrier, max_over_ranks=max_over_ranks, sync_shared_table=sync_shared_table)
        extra["configs"] = {"C3": bench_c3(ctx), "C4": bench_c4(ctx), "C5": bench_c5(ctx)}
        extra["shared_q"] = bench_shared(ctx)
        if world == 1:
            extra["configs"]["C1"] = bench_c1(ctx)

    if rank == 0:
        peak, peak_src = measured_peak()
        per_launch_s = kernel_ms / 1e3 / K
        clk = clocks.summary()
        f_sm = (clk.get("sm_mhz") or clk.get("sm_max_mhz") or 1965) * 1e6
        issue = recorded_issue()
        roof = {"kernel": "k_rollout_private<boat,philox>", "kernel_ms_per_launch": kernel_ms / K,
                # epsilon anneals over the first 100,000 agent steps: early launches explore
                # (environments of a warp spread over the states), later ones run greedy
                "kernel_ms_by_launch": kernel_ms_by_step}
        traffic = recorded_traffic()
        roof["traffic"] = None if traffic is None else traffic.get("dram_bytes_per_launch")
        # what the unfused one-step contract of SURVEY 8(d) would move; the fused kernel keeps
        # state in registers and tables in shared memory, so this is traffic REMOVED, not moved
        contract = B_ALG * n * T / per_launch_s / 1e9
        roof["hbm_contract"] = {"algorithmic_bytes_per_env_step": B_ALG, "algorithmic_GBps": contract, "peak": peak,
                                "ratio_to_peak": contract / peak, "peak_source": peak_src,
                                "note": "not a utilisation: bytes the fusion removed from HBM"}
        if issue is not None:
            # the bound that binds: warp-instruction issue, 4 schedulers x 148 SMs x f_SM
            per_launch = issue["warp_instructions_by_launch"]
            timed = per_launch[W:W + K] if len(per_launch) >= W + K else per_launch
            inst = sum(timed) / len(timed)
            achieved = inst / per_launch_s / 1e9
            issue_peak = 4 * 148 * f_sm / 1e9
            roof.update({"bound": "issue", "achieved": achieved, "peak": issue_peak, "unit": "Gwarp-inst/s",
                         "frac": achieved / issue_peak, "warp_instructions_per_launch": inst,
                         "warp_instructions_per_env_step": inst * 32 / (n * T),
                         "peak_source": "4 issue slots x 148 SMs x SM clock sampled under load (%.0f MHz)" % (f_sm / 1e6),
                         "instruction_count_source": issue.get("source")})
            alu = issue.get("alu_pipe_warp_instructions_by_launch")
            if alu:
                # the unit that actually binds: the integer/logic ALU pipe takes one warp-instruction
                # every 2 cycles per scheduler (B300_MICROARCH.md "Pipe rates": rt_SMSP = 2), and most
                # of the lock-step (selects, shifts, compares, Philox xors) runs on it
                a_timed = alu[W:W + K] if len(alu) >= W + K else alu
                a_inst = sum(a_timed) / len(a_timed)
                alu_peak = 0.5 * 4 * 148 * f_sm / 1e9
                roof["alu_pipe"] = {"achieved": a_inst / per_launch_s / 1e9, "peak": alu_peak, "unit": "Gwarp-inst/s",
                                    "frac": a_inst / per_launch_s / 1e9 / alu_peak,
                                    "warp_instructions_per_env_step": a_inst * 32 / (n * T),
                                    "peak_source": "1 ALU-pipe warp-instruction per 2 cycles x 4 schedulers x 148 SMs x SM clock"}
        else:
            roof.update({"bound": "hbm", "achieved": contract, "peak": peak, "unit": "GB/s", "frac": contract / peak,
                         "peak_source": peak_src})
        line = {
            "metric": "env-steps/sec (step + tabular-Q update)", "value": value, "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": step_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "locksteps_per_call": T,
                       "q_mode": "private", "rng": "philox4x32-10", "l2": "flushed between timed iterations (256 MiB memset)",
                       "episodes_finished": episodes_finished,
                       "multi_gpu_timing": "per rank: sum of its K device-timed steps + the closing statistics collective, "
                                           "entered after an untimed barrier; value uses the max over ranks"},
            "roofline": roof,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "l2": "flushed before every step", "timed": "host wall clock around each sgk_rollout_tabq_host call"},
            "gpu_launches": 4 * K, "stats_allreduce_ms": drain_ms,
            "clocks": clk,
        }
        line.update(extra)
        if world == 1:
            line["hbm_point"] = hbm_point(torch, gridfast, peak)
            t = time.perf_counter()
            rate = cpu_port_rate(1, 150000)
            line["cpu_baseline"] = {
                "value": rate, "unit": "env-steps/s", "cores": 1, "kind": "port",
                "sample": "150000 env-steps of boat-race tabular-Q, python oracle port, 1 process (%.1f s)" % (time.perf_counter() - t),
                "note": "the port's closed-form epsilon schedule skips the reference's O(anneal) list.pop(0) per step "
                        "(value.py:57): it is FASTER than the real reference agent, i.e. a conservative baseline",
                "c_port": c_port_rate()}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------- sub-records
def _timed_calls(ctx, fn, reps, warm):
    """`reps` calls of fn, each preceded by an (untimed) L2 flush, CUDA events on
    the launching stream; returns seconds per call, max over ranks."""
    torch = ctx["torch"]
    for _ in range(warm):
        fn()
    ctx["barrier"]()
    total = 0.0
    per_call = []
    for _ in range(reps):
        ctx["flush"].zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        per_call.append(a.elapsed_time(b) / 1e3)
        total += per_call[-1]
    (total,) = ctx["max_over_ranks"](total)
    return total / reps, per_call


def bench_c3(ctx):
    """BASELINE config 3: side-effects sokoban, tabular Q with hidden-reward
    (safety performance) tracking, 131,072 environments per GPU -- 1,048,576 at
    --gpus 8 -- private tables, 10,000 lock-steps per call."""
    gf, world, rank = ctx["gridfast"], ctx["world"], ctx["rank"]
    n, T = 131072, 10000
    env = gf.BatchedEnv("SideEffectsSokoban-v0", n, seed=0, env_id0=rank * n, device=ctx["local"])
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, **HP)
    sec, per_call = _timed_calls(ctx, lambda: agent.rollout(T), reps=5, warm=3)
    agent.check()
    tot = env.totals()
    peak, _ = measured_peak()
    rate = world * n * T / sec
    b_alg = B_ALG_BY_ENV["SideEffectsSokoban-v0"]
    rec = {"workload": "side-effects sokoban tabular-Q, 131072 envs/GPU (%d global), private hashed tables (128 slots), "
                       "10000 lock-steps per call" % (world * n),
           "value": rate, "unit": "env-steps/s", "ms_per_call": 1e3 * sec, "kernel": "k_rollout_private<sokoban,philox>",
           "kernel_ms_by_call": [round(1e3 * s, 3) for s in per_call],
           "roofline": {"bound": "hbm", "achieved": rate / world * b_alg / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": rate / world * b_alg / 1e9 / peak, "algorithmic_bytes_per_env_step": b_alg,
                        "note": "671 MB of private tables per GPU do not fit L2: bound by dependent table loads "
                                "(ncu long_scoreboard), see profiles/"},
           "mean_return": tot["sum_return"] / max(tot["episodes"], 1),
           "mean_safety_performance": tot["sum_performance"] / max(tot["episodes"], 1), "episodes_rank0": tot["episodes"]}
    del agent, env
    ctx["torch"].cuda.empty_cache()
    return rec


def bench_c4(ctx):
    """BASELINE config 4 as SURVEY 8(d) specifies it: tomato watering, 65,536
    environments per GPU, SSRL agent with C_prior .01, budget 1,000 queries per
    environment, warm-up fraction .5 -> 500 random-policy warm-up episodes per
    environment (ssrl/warmup.py), then 10,000 learning lock-steps per call."""
    gf, world, rank, torch = ctx["gridfast"], ctx["world"], ctx["rank"], ctx["torch"]
    n, T, budget, warm = 65536, 10000, 1000, 0.5
    env = gf.BatchedEnv("TomatoWatering-v0", n, seed=0, env_id0=rank * n, device=ctx["local"])
    # 16,384 slots hold the 30,000 learning lock-steps measured here (fullest table: about 10,400 keys) without a
    # rehash inside a timed call; tables otherwise start at 4,096 slots and double on demand
    agent = gf.BatchedTabularQ(env, gf.Q_PRIVATE, capacity=16384, **HP)
    agent.enable_ssrl(c_prior=0.01, budget=budget)
    n_warm = int(budget * warm)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx["barrier"]()
    a.record()
    agent.ssrl_warmup(n_warm)
    b.record()
    torch.cuda.synchronize()
    (warm_s,) = ctx["max_over_ranks"](a.elapsed_time(b) / 1e3)
    env.clear_stats()
    sec, per_call = _timed_calls(ctx, lambda: agent.rollout(T), reps=2, warm=1)
    agent.check()
    tot = env.totals()
    bud, eps, corrupt = [t.double().mean().item() for t in agent.ssrl_counters()]
    peak, _ = measured_peak()
    rate = world * n * T / sec
    b_alg = B_ALG_BY_ENV["TomatoWatering-v0"]
    rec = {"workload": "tomato watering + SSRL (C_prior .01, budget 1000, warm-up .5), 65536 envs/GPU, private hashed "
                       "tables (16384 slots, pre-sized for this run; they grow on demand), 10000 lock-steps per call",
           "value": rate, "unit": "env-steps/s", "ms_per_call": 1e3 * sec, "kernel": "k_rollout_private<tomato,philox,ssrl>",
           "kernel_ms_by_call": [round(1e3 * s, 3) for s in per_call],
           "warmup": {"episodes_per_env": n_warm, "seconds": warm_s, "env_steps_per_s": world * n * n_warm * 100 / warm_s,
                      "kernel": "k_rollout_random<tomato,philox> (episodic)"},
           "roofline": {"bound": "hbm", "achieved": rate / world * b_alg / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": rate / world * b_alg / 1e9 / peak, "algorithmic_bytes_per_env_step": b_alg,
                        "note": "tables of %d slots x 65536 envs = %.1f GB per GPU: bound by dependent probes into HBM"
                                % (agent.capacity, agent.capacity * n * 48 / 1e9)},
           "table_capacity": agent.capacity, "table_max_fill": agent.max_fill(),
           "mean_budget_left": bud, "mean_episodes": eps, "mean_corrupt_episodes": corrupt,
           "mean_return": tot["sum_return"] / max(tot["episodes"], 1),
           "mean_safety_performance": tot["sum_performance"] / max(tot["episodes"], 1)}
    del agent, env
    torch.cuda.empty_cache()
    return rec


def bench_c5(ctx):
    """BASELINE config 5: side-effects sokoban DeepQAgent, 4,096 lock-step
    environments sharing one 36-100-100-4 Q network, replay ring in HBM, the
    reference's ratio of one 64-sample optimiser step per env-step (batch
    262,144 per lock-step), network math on tcgen05."""
    gf, world, rank, torch = ctx["gridfast"], ctx["world"], ctx["rank"], ctx["torch"]
    n, batch, T = 4096, 64 * 4096, 500
    env = gf.BatchedEnv("SideEffectsSokoban-v0", n, seed=0, env_id0=rank * n, device=ctx["local"])
    agent = gf.BatchedDeepQ(env, replay_capacity=100 * n, batch_size=batch, lr=1e-3, epsilon=0.01,
                            epsilon_anneal=100000, sync_every=10000, reference_bxb_loss=True)
    agent.set_tensor_cores(True)
    agent.warmup(100)
    sec, per_call = _timed_calls(ctx, lambda: agent.rollout(T), reps=2, warm=1)
    flop = T * (n * 28000.0 + batch * 4 * 28000.0)          # act forward + 4x forward per learned sample
    tf = flop / sec / 1e12
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            tpeak = float(json.load(f)["bf16_tflops_sustained"])
            tsrc = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        tpeak, tsrc = 1400.0, "fallback (B200_PROFILING.md sustained)"
    rec = {"workload": "side-effects sokoban deep-Q, 4096 envs/GPU, one shared MLP 36-100-100-4, replay ring 409600 "
                       "transitions in HBM, batch 262144 per lock-step (= 64 learned samples per env-step), "
                       "%d lock-steps per call" % T,
           "value": world * n * T / sec, "unit": "env-steps/s", "ms_per_call": 1e3 * sec, "us_per_lockstep": 1e6 * sec / T,
           "learned_samples_per_s": world * batch * T / sec, "precision": agent.precision,
           "roofline": {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                        "peak_source": tsrc, "flops": "model flops: 28.0 kFLOP per forward, x4 per learned sample"},
           "loss_gradnorm_clip": agent.last_scalars()}
    del agent, env
    torch.cuda.empty_cache()
    return rec


def bench_c1(ctx):
    """BASELINE config 1, `python main.py boat tabular-q --lr .5`: ONE
    environment behind the reference's own loop shape, through the fused
    LEARN_MAP function (one launch per episode, Train/epsilon logged per step)."""
    import argparse

    import numpy as np
    gf = ctx["gridfast"]

    class NullWriter:
        def add_scalar(self, *a, **k):
            pass
        add_scalars = add_scalar

    class Meter:
        val = avg = max = 0.0

        def update(self, v, n=1):
            self.val = v

    args = argparse.Namespace(lr=0.5, discount=0.99, epsilon=0.01, epsilon_anneal=100000, cheat=False, eval_every=10 ** 9)
    np.random.seed(0)
    env = gf.make("BoatRace-v0", rng="numpy")
    agent = gf.GpuTabularQAgent(env, args)
    history = {"writer": NullWriter(), "t": 0, "episode": 0, "returns": Meter(), "safeties": Meter(), "margins": Meter(),
               "margins_support": Meter()}

    def episodes(m):
        for _ in range(m):
            state = (env.reset(), 0.0, False, {})
            history["episode"] += 1
            gf.tabq_learn_fused(agent, env, state, history, args)
    episodes(20)
    t0, s0 = time.perf_counter(), history["t"]
    episodes(300)
    dt = time.perf_counter() - t0
    steps = history["t"] - s0
    return {"workload": "boat-race tabular-Q, ONE environment, reference loop shape (train.py:62-70) through "
                        "gridfast.tabq_learn_fused on numpy's global stream, 300 episodes",
            "value": steps / dt, "unit": "env-steps/s", "us_per_env_step": 1e6 * dt / steps,
            "launches_per_episode": 5, "note": "latency-bound by construction (one environment); the reference agent "
            "half alone costs 26-48 us per step on one core (SURVEY 3.5)"}


def bench_shared(ctx):
    """Shared-table mode: every GPU keeps a replica of ONE Q table for all its
    environments; every `sync_interval` lock-steps the replicas are merged with
    one NCCL all-reduce of the dense delta-Q array.  Reports the
    rollout rate and the cost of a sync; checks that all replicas end identical."""
    gf, world, rank, torch, dist = ctx["gridfast"], ctx["world"], ctx["rank"], ctx["torch"], ctx["dist"]
    out = {}
    for name, env_id, n, T in (("boat", "BoatRace-v0", 65536, 2000), ("sokoban", "SideEffectsSokoban-v0", 131072, 1000)):
        for interval in (100, 1000):
            env = gf.BatchedEnv(env_id, n, seed=0, env_id0=rank * n, device=ctx["local"])
            agent = gf.BatchedTabularQ(env, gf.Q_SHARED, **HP)
            agent.rebase()
            sync_ms = []

            def run():
                done = 0
                while done < T:
                    agent.rollout(interval)
                    done += interval
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    ctx["sync_shared_table"](agent)
                    b.record()
                    sync_ms.append((a, b))
            sec, _ = _timed_calls(ctx, run, reps=2, warm=1)
            agent.check()
            ms = sorted(a.elapsed_time(b) for a, b in sync_ms[-2 * (T // interval):])
            keys, rows = agent.export(0)
            order = keys.argsort()
            blob = torch.as_tensor(rows[order].reshape(-1).copy()).to(ctx["dev"])
            same = True
            if world > 1:
                sizes = [torch.zeros(1, dtype=torch.int64, device=ctx["dev"]) for _ in range(world)]
                dist.all_gather(sizes, torch.tensor([blob.numel()], device=ctx["dev"]))
                same = all(int(s) == blob.numel() for s in sizes)
                if same:
                    blobs = [torch.empty_like(blob) for _ in range(world)]
                    dist.all_gather(blobs, blob)
                    same = all(torch.equal(x, blobs[0]) for x in blobs)
            out["%s_sync%d" % (name, interval)] = {
                "envs_per_gpu": n, "sync_interval": interval, "value": world * n * T / sec, "unit": "env-steps/s",
                "sync_ms_median": ms[len(ms) // 2], "sync_ms_max": ms[-1], "syncs_timed": len(ms),
                "collective": "1 x all_reduce(sum) of %d x 5 float64 (dense delta-Q + presence)" % agent.dense_size(),
                "states": int(len(keys)), "replicas_identical": bool(same)}
            del agent, env
            torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--main-only", action="store_true", help="only the headline C2 measurement (skip C1/C3/C4/C5 and the shared-table leg)")
    args = ap.parse_args()
    out = _claim_stdout()
    if args.impl == "reference":
        run_reference(args, out)
    else:
        run_ours(args, out)


if __name__ == "__main__":
    main()
