// sgk_table.cuh -- the tabular-Q store.
//
// Replaces `Q = defaultdict(lambda: np.zeros(A))` keyed by
// tuple(board.flatten()) (safe_grid_agents/common/agents/value.py:31-35) with
// open-addressing tables of float64 rows keyed by the lossless 64-bit
// observation code (sgk_envs.cuh obs_key): exact keys, so lookups can never
// alias two boards -- same semantics as the dict.
//
// Two layouts in HBM, chosen per table set at creation (TableView strides):
//   slot-major   keys[slot][table] u64 (0 = empty), q[slot][table][4] f64
//       the dense boat-race tables: 8 slots, and the 32 environments of a
//       warp mostly sit in neighbouring slots, so a warp's probe / row load is
//       a few contiguous 256 B / 1 KB runs; a shared table is its
//       n_tables == 1 case;
//   table-major  keys[table][slot], q[table][slot][4]
//       hashed private tables: every environment probes a different random
//       slot, so nothing coalesces either way, but each table is one
//       contiguous 40 * cap byte region -- a warp's 32 accesses fall into a
//       handful of 2 MB pages instead of 64 (TLB reach, DRAM page locality).
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
    uint32_t n_tables;            // cap * n_tables < 2^32 (checked at creation)
    uint32_t slot_stride, table_stride;   // entry index = slot * slot_stride + table * table_stride
    uint32_t cap, log_cap;
    uint32_t dense_open;          // != 0: minimal perfect hash (boat race), see dense_slot
    uint32_t perfect_n;           // != 0: perfect index rank(agent) * perfect_n + rank(box) (sokoban level 0)
    const uint8_t *perfect_rank;  // ... the rank of a cell among the non-wall cells, 0xFF for a wall
};

// Multiplicative hash of the two key halves, 32-bit arithmetic only.
__device__ __forceinline__ uint32_t home_slot(uint64_t key, uint32_t log_cap)
{
    const uint32_t h = (uint32_t)key * 0x9E3779B1u + (uint32_t)(key >> 32) * 0x85EBCA77u;
    return h >> (32 - log_cap);
}

// Boat race: the observation is a function of the agent's cell alone, so the
// rank of that cell among the open cells is a minimal perfect hash (8 slots).
__device__ __forceinline__ uint32_t dense_slot(uint32_t open32, uint32_t pos)
{
    return __popc(open32 & ((1u << pos) - 1u));
}

// Sokoban level 0: the observation is (agent cell, box cell), both non-wall: 11 x 11 = 121 of the 128
// slots, no probing and no key compare -- a row that was never written reads as the zero row of an
// unseen state, so lookups need not read the key at all; keys are written on touch, so the key set
// still equals the reference dict's.  A board that puts the agent or the box on a wall has no slot.
__device__ __forceinline__ uint32_t perfect_slot(const TableView &T, uint64_t key)
{
    const uint32_t a = T.perfect_rank[(uint32_t)key & 0xFFu], b = T.perfect_rank[(uint32_t)(key >> 8) & 0xFFu];
    return (a == 0xFFu || b == 0xFFu) ? SGK_NOSLOT : a * T.perfect_n + b;
}

__device__ __forceinline__ size_t entry(const TableView &T, uint32_t slot, uint32_t g)
{
    return (size_t)(slot * T.slot_stride + g * T.table_stride);
}

// find-or-insert in a table only this thread touches
__device__ __forceinline__ uint32_t find_private(const TableView &T, uint32_t g, uint64_t key, int *status)
{
    if (T.dense_open) {
        const uint32_t s = dense_slot(T.dense_open, (uint32_t)key & 0xFFu);
        T.keys[entry(T, s, g)] = key;
        return s;
    }
    if (T.perfect_n) {
        const uint32_t s = perfect_slot(T, key);
        if (s == SGK_NOSLOT) *status = SGK_ST_FULL;
        else T.keys[entry(T, s, g)] = key;
        return s;
    }
    uint32_t s = home_slot(key, T.log_cap);
    for (uint32_t i = 0; i < T.cap; i++) {
        unsigned long long *p = T.keys + entry(T, s, g);
        const unsigned long long k = *p;
        if (k == key) return s;
        if (k == 0ull) { *p = key; return s; }
        s = (s + 1) & (T.cap - 1);
    }
    // Full: the state reads as an unseen one (zero row) and is not written;
    // the sticky status makes the next synchronising call fail with SGK_EFULL.
    // (The host grows hashed private tables ahead of time, sgk_tabq_grow, so
    // this is reached only when device memory is exhausted.)
    *status = SGK_ST_FULL;
    return SGK_NOSLOT;
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
    return SGK_NOSLOT;
}

// lookup without insertion; returns false when absent
__device__ __forceinline__ bool lookup(const TableView &T, uint32_t g, uint64_t key, uint32_t &slot)
{
    if (T.dense_open) {
        slot = dense_slot(T.dense_open, (uint32_t)key & 0xFFu);
        return T.keys[entry(T, slot, g)] == key;
    }
    if (T.perfect_n) {
        slot = perfect_slot(T, key);
        return slot != SGK_NOSLOT && T.keys[entry(T, slot, g)] == key;
    }
    uint32_t s = home_slot(key, T.log_cap);
    for (uint32_t i = 0; i < T.cap; i++) {
        const unsigned long long k = T.keys[entry(T, s, g)];
        if (k == key) { slot = s; return true; }
        if (k == 0ull) return false;
        s = (s + 1) & (T.cap - 1);
    }
    return false;
}

struct QRow { double v0, v1, v2, v3; };

// slot == SGK_NOSLOT (table full): the row of a state the table could not take
__device__ __forceinline__ QRow load_row(const TableView &T, uint32_t g, uint32_t slot)
{
    if (slot == SGK_NOSLOT) return QRow{0.0, 0.0, 0.0, 0.0};
    const double2 *p = reinterpret_cast<const double2 *>(T.q + entry(T, slot, g) * SGK_NA);
    const double2 a = p[0], b = p[1];
    QRow r; r.v0 = a.x; r.v1 = a.y; r.v2 = b.x; r.v3 = b.y;
    return r;
}

__device__ __forceinline__ void store_q(const TableView &T, uint32_t g, uint32_t slot, int a, double v)
{
    if (slot == SGK_NOSLOT) return;
    T.q[entry(T, slot, g) * SGK_NA + a] = v;
}

// Hashed private table, hot path: probe the home slot and fetch its row in
// the same round trip; anything else (collision, first touch) goes the slow way.
static __device__ __noinline__ uint32_t find_private_slow(const TableView &T, uint32_t g, uint64_t key, int *status)
{
    return find_private(T, g, key, status);
}

__device__ __forceinline__ uint32_t find_row_private(const TableView &T, uint32_t g, uint64_t key, QRow &row, int *status)
{
    uint32_t s = home_slot(key, T.log_cap);
    const unsigned long long k = T.keys[entry(T, s, g)];
    row = load_row(T, g, s);
    if (k != key) {
        s = find_private_slow(T, g, key, status);
        row = load_row(T, g, s);
    }
    return s;
}

// np.argmax: first maximum wins (value.py:35)
// ... and the maximum itself (what learn bootstraps from, value.py:48-50).
// Two-level tournament: the halves compare in parallel, so the dependent chain
// is two compare+select levels instead of three; strict > at both levels keeps
// "first maximum wins" (a tie between the halves goes to the lower one).
__device__ __forceinline__ int argmax_first(const QRow &r, double &m)
{
    const bool p01 = r.v1 > r.v0, p23 = r.v3 > r.v2;
    const double m01 = p01 ? r.v1 : r.v0, m23 = p23 ? r.v3 : r.v2;
    const bool hi = m23 > m01;
    m = hi ? m23 : m01;
    return hi ? (p23 ? 3 : 2) : (p01 ? 1 : 0);
}

__device__ __forceinline__ int argmax_first(const QRow &r)
{
    double m;
    return argmax_first(r, m);
}

__device__ __forceinline__ double row_max(const QRow &r)
{
    const double m01 = r.v1 > r.v0 ? r.v1 : r.v0, m23 = r.v3 > r.v2 ? r.v3 : r.v2;
    return m23 > m01 ? m23 : m01;
}

// branch-free element access (selects, not register shuffles behind branches)
__device__ __forceinline__ double row_get(const QRow &r, int a)
{
    const double lo = (a & 1) ? r.v1 : r.v0;
    const double hi = (a & 1) ? r.v3 : r.v2;
    return (a & 2) ? hi : lo;
}

__device__ __forceinline__ void row_set_if(QRow &r, bool cond, int a, double v)
{
    r.v0 = (cond && a == 0) ? v : r.v0;
    r.v1 = (cond && a == 1) ? v : r.v1;
    r.v2 = (cond && a == 2) ? v : r.v2;
    r.v3 = (cond && a == 3) ? v : r.v3;
}

__device__ __forceinline__ void row_set(QRow &r, int a, double v) { row_set_if(r, true, a, v); }

// One Q-learning update, rounded exactly like the reference's float64 numpy
// arithmetic (value.py:50-52): no FMA contraction.
__device__ __forceinline__ double td_update(double q_sa, double reward, double discount, double lr, double next_best)
{
    const double target = __dadd_rn(reward, __dmul_rn(discount, next_best));
    const double differential = __dsub_rn(target, q_sa);
    return __dadd_rn(q_sa, __dmul_rn(lr, differential));
}
