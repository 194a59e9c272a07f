// sgk_table.cuh -- the tabular-Q store.
//
// Replaces `Q = defaultdict(lambda: np.zeros(A))` keyed by
// tuple(board.flatten()) (safe_grid_agents/common/agents/value.py:31-35) with
// open-addressing tables of float64 rows keyed by the lossless 64-bit
// observation code (sgk_envs.cuh obs_key): exact keys, so lookups can never
// alias two boards -- same semantics as the dict.
//
// Layout in HBM: slot-major across tables,
//     keys[slot][table]        u64   (0 = empty)
//     q   [slot][table][4]     f64
// so that when the 32 environments of a warp (32 private tables) sit in the
// same state, their probes and row loads are one contiguous 256 B / 1 KB
// access; a shared table is the n_tables == 1 case of the same layout.
#pragma once
#include "sgk_common.cuh"

#define SGK_ST_FULL 1
#define SGK_ST_REPLAY_DRY 2
#define SGK_NOSLOT 0xFFFFFFFFu

struct TableView {
    unsigned long long *keys;
    double *q;
    double *c;                    // SSRL corruption estimate per slot (or null)
    unsigned long long *winner;   // shared mode: [cap][4] election words
    long long n_tables;
    uint32_t cap, log_cap;
};

__device__ __forceinline__ uint32_t home_slot(uint64_t key, uint32_t log_cap)
{
    return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> (64 - log_cap));
}

// find-or-insert in a table only this thread touches
__device__ __forceinline__ uint32_t find_private(const TableView &T, long long g, uint64_t key, int *status)
{
    uint32_t s = home_slot(key, T.log_cap);
    for (uint32_t i = 0; i < T.cap; i++) {
        unsigned long long *p = T.keys + (size_t)s * T.n_tables + g;
        const unsigned long long k = *p;
        if (k == key) return s;
        if (k == 0ull) { *p = key; return s; }
        s = (s + 1) & (T.cap - 1);
    }
    *status = SGK_ST_FULL;
    return 0;
}

// find-or-insert in a table many threads probe concurrently
__device__ __forceinline__ uint32_t find_shared(const TableView &T, uint64_t key, int *status)
{
    uint32_t s = home_slot(key, T.log_cap);
    for (uint32_t i = 0; i < T.cap; i++) {
        unsigned long long *p = T.keys + s;
        unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(p);
        if (k == 0ull) k = atomicCAS(p, 0ull, (unsigned long long)key);
        if (k == key || k == 0ull) return s;
        s = (s + 1) & (T.cap - 1);
    }
    *status = SGK_ST_FULL;
    return 0;
}

// lookup without insertion; returns false when absent
__device__ __forceinline__ bool lookup(const TableView &T, long long g, uint64_t key, uint32_t &slot)
{
    uint32_t s = home_slot(key, T.log_cap);
    for (uint32_t i = 0; i < T.cap; i++) {
        const unsigned long long k = T.keys[(size_t)s * T.n_tables + g];
        if (k == key) { slot = s; return true; }
        if (k == 0ull) return false;
        s = (s + 1) & (T.cap - 1);
    }
    return false;
}

struct QRow { double v0, v1, v2, v3; };

__device__ __forceinline__ QRow load_row(const TableView &T, long long g, uint32_t slot)
{
    const double2 *p = reinterpret_cast<const double2 *>(T.q + ((size_t)slot * T.n_tables + g) * SGK_NA);
    const double2 a = p[0], b = p[1];
    QRow r; r.v0 = a.x; r.v1 = a.y; r.v2 = b.x; r.v3 = b.y;
    return r;
}

__device__ __forceinline__ void store_q(const TableView &T, long long g, uint32_t slot, int a, double v)
{
    T.q[((size_t)slot * T.n_tables + g) * SGK_NA + a] = v;
}

// np.argmax: first maximum wins (value.py:35)
__device__ __forceinline__ int argmax_first(const QRow &r)
{
    int b = 0; double m = r.v0;
    if (r.v1 > m) { m = r.v1; b = 1; }
    if (r.v2 > m) { m = r.v2; b = 2; }
    if (r.v3 > m) { m = r.v3; b = 3; }
    return b;
}

__device__ __forceinline__ double row_max(const QRow &r)
{
    double m = r.v0;
    if (r.v1 > m) m = r.v1;
    if (r.v2 > m) m = r.v2;
    if (r.v3 > m) m = r.v3;
    return m;
}

__device__ __forceinline__ double row_get(const QRow &r, int a)
{
    return a == 0 ? r.v0 : a == 1 ? r.v1 : a == 2 ? r.v2 : r.v3;
}

__device__ __forceinline__ void row_set(QRow &r, int a, double v)
{
    if (a == 0) r.v0 = v; else if (a == 1) r.v1 = v; else if (a == 2) r.v2 = v; else r.v3 = v;
}

// One Q-learning update, rounded exactly like the reference's float64 numpy
// arithmetic (value.py:50-52): no FMA contraction.
__device__ __forceinline__ double td_update(double q_sa, double reward, double discount, double lr, double next_best)
{
    const double target = __dadd_rn(reward, __dmul_rn(discount, next_best));
    const double differential = __dsub_rn(target, q_sa);
    return __dadd_rn(q_sa, __dmul_rn(lr, differential));
}
