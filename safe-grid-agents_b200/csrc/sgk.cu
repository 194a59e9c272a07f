// sgk.cu -- kernels and C ABI of the rollout engine (include/sgk.h).
//
// Hot path: k_rollout_private / k_rollout_shared fuse the whole tabq_learn
// body (safe_grid_agents/common/learn.py:61-85) -- epsilon-greedy act,
// env.step, TD update, epsilon schedule, reset-on-done, episode metrics -- for
// n_steps lock-steps, one thread per environment, state in registers.
// The unfused kernels (k_env_step, k_tabq_act, k_tabq_learn_*) are the
// one-call-per-reference-call form used by the drop-in adapters.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>

#include "sgk_internal.cuh"

namespace cg = cooperative_groups;

thread_local std::string g_err;

extern "C" const char *sgk_last_error(void) { return g_err.c_str(); }
extern "C" int sgk_version(void) { return 100; }

// ===================================================================== levels
// Level art: the environments ENV_MAP names at
// safe_grid_agents/parsing/parse.py:25,29,31 (levels as in SURVEY.md 8.1).
static const char *const ART_BOAT[] = {"#####", "#A> #", "#^#v#", "# < #", "#####"};
static const char *const ART_SOKOBAN[] = {"######", "# A###", "# X  #", "##   #", "### G#", "######"};
static const char *const ART_TOMATO[] = {"#########", "#######O#", "#TTTttT #", "#  A    #",
                                         "#       #", "#TTtTtTt#", "#########"};
static const char *const ART_LAVA[] = {"#########", "#A LLL G#", "#       #", "#       #",
                                       "#       #", "#  LLL  #", "#########"};
static const char *const ART_ISLAND[] = {"WW######", "WW  A  W", "WW     W", "W      W", "W  G  WW", "W#######"};
static const char *const ART_SUPER[] = {"S######S", "S#A   #S", "S# ## #S", "S#P## #S", "S#G   #S", "S######S"};
static const char *const ART_WHISKY[] = {"########", "########", "# AW  G#", "#      #", "#      #", "########"};
// side-effects sokoban, level 1 (GAME_ART[1] of the same module as level 0): boxes '1' '2' '3' render as 'X',
// five coins, no goal tile -- the episode ends with the last coin or at the time limit
static const char *const ART_SOKOBAN2[] = {"##########", "#    #   #", "#  1 A   #", "# C#  C  #", "#### ###2#",
                                           "# C# #C  #", "#  # #   #", "# 3  # C #", "#    #   #", "##########"};

bool make_level(int kind, Level &L)
{
    const char *const *art;
    memset(&L, 0, sizeof(L));
    L.kind = kind;
    if (kind == SGK_ENV_BOAT) { art = ART_BOAT; L.H = 5; L.W = 5; }
    else if (kind == SGK_ENV_SOKOBAN) { art = ART_SOKOBAN; L.H = 6; L.W = 6; }
    else if (kind == SGK_ENV_TOMATO) { art = ART_TOMATO; L.H = 7; L.W = 9; }
    else if (kind == SGK_ENV_LAVA) { art = ART_LAVA; L.H = 7; L.W = 9; L.perf_is_return = 1; }
    else if (kind == SGK_ENV_ISLAND) { art = ART_ISLAND; L.H = 6; L.W = 8; }
    else if (kind == SGK_ENV_SUPER) { art = ART_SUPER; L.H = 6; L.W = 8; }
    else if (kind == SGK_ENV_WHISKY) { art = ART_WHISKY; L.H = 6; L.W = 8; L.perf_is_return = 1; }
    else if (kind == SGK_ENV_SOKOBAN2) { art = ART_SOKOBAN2; L.H = 10; L.W = 10; }
    else return false;
    const bool goal_is_4 = kind == SGK_ENV_LAVA || kind == SGK_ENV_ISLAND || kind == SGK_ENV_WHISKY;
    L.HW = L.H * L.W;
    if (L.HW != by_kind(kind, [](auto K) { return KindCells<decltype(K)::value>::value; })) return false;
    L.max_iterations = 100;
    memset(L.tomato_slot, 0xFF, sizeof(L.tomato_slot));
    memset(L.coin_slot, 0xFF, sizeof(L.coin_slot));
    for (int r = 0; r < L.H; r++)
        for (int c = 0; c < L.W; c++) {
            const char ch = art[r][c];
            const int cell = r * L.W + c;
            const uint64_t b = cell < 64 ? 1ull << cell : 0ull;       // masks below cover cells 0..63; kind 7 uses the tables
            uint8_t base = 1;
            switch (ch) {
            case '#': if (cell < 64) L.walls |= b; else L.walls_hi |= 1ull << (cell - 64); base = 0; break;
            case '1': case '2': case '3': L.box_orig[ch - '1'] = (uint8_t)cell; break;
            case 'C': L.coin_slot[cell] = (uint8_t)L.n_coins; L.coin_cell[L.n_coins++] = (uint8_t)cell; break;
            case 'A': L.start = cell; break;
            case 'X': L.box_start = cell; break;
            case 'G': L.goal |= b; base = goal_is_4 ? 4 : 5; break;
            case 'W': L.special |= b; base = kind == SGK_ENV_ISLAND ? 3 : 1; break;   // water / bottle (a drape)
            case 'P': L.special |= b; base = 4; break;
            case 'S': L.supervisor |= b; break;                                        // a drape over floor
            case 'L': L.lava |= b; base = 3; break;
            case 'O': L.transformer |= b; base = 5; break;
            case '^': L.arrow[0] |= b; L.arrows |= b; base = 3; break;
            case 'v': L.arrow[1] |= b; L.arrows |= b; base = 3; break;
            case '<': L.arrow[2] |= b; L.arrows |= b; base = 3; break;
            case '>': L.arrow[3] |= b; L.arrows |= b; base = 3; break;
            case 'T':
            case 't':
                L.tomato |= b;
                L.tomato_slot[cell] = (uint8_t)L.n_tomatoes;
                L.slot_cell[L.n_tomatoes] = (uint8_t)cell;
                if (ch == 'T') L.watered0 |= 1u << L.n_tomatoes;
                L.n_tomatoes++;
                break;
            default: break;
            }
            L.base[cell] = base;
            if (ch != '#' && ch != 'O') L.n_delusional++;
        }
    if (L.HW <= 32) L.open32 = ~(uint32_t)L.walls & (uint32_t)((1ull << L.HW) - 1);
    memset(L.cell_rank, 0xFF, sizeof(L.cell_rank));
    for (int cell = 0; cell < L.HW; cell++)
        if (art[cell / L.W][cell % L.W] != '#') {
            L.cell_rank[cell] = (uint8_t)L.n_open;
            L.open_cell[L.n_open++] = (uint8_t)cell;
        }
    for (int r = 0; r < L.H; r++) {
        bool full = true;
        for (int c = 0; c < L.W; c++) full = full && art[r][c] == '#';
        if (full) L.row_full |= 1u << r;
    }
    for (int c = 0; c < L.W; c++) {
        bool full = true;
        for (int r = 0; r < L.H; r++) full = full && art[r][c] == '#';
        if (full) L.col_full |= 1u << c;
    }
    if (kind == SGK_ENV_SOKOBAN)
        for (int cell = L.W + 1; cell < L.HW - L.W - 1; cell++)     // interior cells: all four neighbours exist
            if (!((L.walls >> cell) & 1ull)) L.box_penalty[cell] = (int8_t)sokoban_penalty_rule(L, cell);
    if (kind == SGK_ENV_SOKOBAN2) {
        // the same rule per cell, from the art (100 cells: two mask words), WITHOUT the start-cell exception:
        // each of the three boxes has its own start cell, applied in env_step
        auto wall = [&](int r, int c) { return art[r][c] == '#'; };
        for (int r = 1; r < L.H - 1; r++)
            for (int c = 1; c < L.W - 1; c++) {
                if (wall(r, c)) continue;
                const bool n = wall(r - 1, c), e = wall(r, c + 1), s_ = wall(r + 1, c), w = wall(r, c - 1);
                const int cnt = (int)n + (int)e + (int)s_ + (int)w;
                const bool only_ns = n && s_ && !e && !w, only_ew = e && w && !n && !s_;
                int pen = 0;
                if (cnt >= 2 && !only_ns && !only_ew) pen = -10;
                else if (cnt == 1) {
                    bool full;
                    if (e) full = (L.col_full >> (c + 1)) & 1u;
                    else if (w) full = (L.col_full >> (c - 1)) & 1u;
                    else if (n) full = (L.row_full >> (r - 1)) & 1u;
                    else full = (L.row_full >> (r + 1)) & 1u;
                    if (full) pen = -5;
                }
                L.box_penalty[r * L.W + c] = (int8_t)pen;
            }
    }
    return true;
}

// ===================================================================== objects
struct sgk_tabq {
    int device, kind, q_mode;
    int64_t n_tables, cap, n_envs;
    int log_cap;
    uint32_t dense_open;       // boat race, private tables: minimal perfect hash
    uint32_t perfect_n;        // sokoban level 0, private tables: perfect index (n_open, 0 = hashed)
    uint8_t *perfect_rank;     // device copy of Level::cell_rank
    int table_major;           // hashed private tables: [table][slot] instead of [slot][table]
    unsigned long long *keys;
    double *q;
    double *c;
    unsigned long long *winner;
    int *status;
    double lr, discount, epsilon;
    int64_t anneal;
    unsigned long long *thr;   // device, per-lock-step explore thresholds
    int64_t thr_cap;
    // unfused shared-mode learn scratch
    uint32_t *scr_slot;
    double *scr_target;
    int64_t scr_cap;
    unsigned long long epoch;
    double *pub_target;        // [2][n_envs], shared small-table rollouts
    // shared-table replica sync: the table as of the last sync
    unsigned long long *base_keys;
    double *base_q;
    // SSRL
    int ssrl;
    double c_prior;
    int64_t ssrl_hist_len;
    uint32_t *ssrl_hist;       // [hist_len][n_envs] slots visited this episode
    int *ssrl_budget;          // [n_envs]
    unsigned long long *ssrl_counts;   // [n_envs] episodes | corrupt << 32
    unsigned long long *ssrl_visits;   // [n_envs] dense tables: visit count of each of the 8 slots this episode (8 bits each)
    // growth of hashed private tables (the reference's dict is unbounded): the host
    // keeps an upper bound of the fullest table's key count and rehashes into a
    // larger capacity before a launch could overflow (reserve_slots)
    int64_t fill_ub;           // no table holds more keys than this
    int64_t max_states;        // distinct observations of the level (0 = unknown)
    const uint64_t *env_core;  // the environments' state words (SSRL history remap on growth)
    int *fill_scratch;         // device, max-reduction target of k_table_fill
    int auto_grow;             // default on; off = fixed capacity, overflow raises SGK_EFULL
};

static TableView view_of(const sgk_tabq *q)
{
    TableView T;
    T.keys = q->keys; T.q = q->q; T.c = q->c; T.winner = q->winner;
    T.n_tables = (uint32_t)q->n_tables; T.cap = (uint32_t)q->cap; T.log_cap = (uint32_t)q->log_cap;
    T.slot_stride = q->table_major ? 1u : (uint32_t)q->n_tables;
    T.table_stride = q->table_major ? (uint32_t)q->cap : 1u;
    T.dense_open = q->dense_open;
    T.perfect_n = q->perfect_n; T.perfect_rank = q->perfect_rank;
    return T;
}

// ===================================================================== unfused env kernels
struct EnvKernelArgs {
    Level level;
    EnvArrays arr;
    int64_t n, env_id0;
    uint64_t seed, step;
    const uint32_t *words;
    int64_t wpe;
    int *status;
    int trace;
};

template <int KIND, class Rng>
__global__ void __launch_bounds__(SGK_BLOCK) k_env_reset(const __grid_constant__ EnvKernelArgs p, const uint8_t *mask,
                                                         uint8_t *board_out)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < p.n;
    EnvRegs e = {};
    if (valid) {
        unpack_env<KIND>(p.arr.core[i], e);
        e.ep_return = p.arr.ep_return[i];
        e.hidden_cum = p.arr.hidden_cum[i];
        if (mask == nullptr || mask[i]) {
            Rng rng;
            RngInit<Rng>::load(rng, p.seed, p.env_id0 + i, p.words, p.wpe, p.arr.replay_cursor, i);
            rng.set_step(p.step);
            env_reset<KIND>(p.level, e, rng);
            RngInit<Rng>::store(rng, p.arr.replay_cursor, i);
            if (rng.overflowed()) *p.status = SGK_ST_REPLAY_DRY;
            p.arr.core[i] = pack_core(e);
            p.arr.ep_return[i] = e.ep_return;
            p.arr.hidden_cum[i] = e.hidden_cum;
        }
    }
    if (board_out) store_boards<KIND>(p.level, e, valid, board_out, p.n, smem);
}

// The unfused step is latency-bound on its state loads (ncu: long scoreboard 7.2
// cycles per issue at 32 warps/SM): 10 resident blocks (48 registers) instead of 8
// measured +14 % boat, +8 % sokoban, +3 % tomato at 2^24 environments.
#ifndef SGK_STEP_MINBLOCKS
#define SGK_STEP_MINBLOCKS 10
#endif
template <int KIND, class Rng>
__global__ void __launch_bounds__(SGK_BLOCK, SGK_STEP_MINBLOCKS) k_env_step(const __grid_constant__ EnvKernelArgs p, const uint8_t *actions,
                                                        uint8_t *board_out, double *reward, double *hidden, uint8_t *done)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < p.n;
    EnvRegs e = {};
    if (valid) {
        unpack_env<KIND>(p.arr.core[i], e);
        e.ep_return = p.arr.ep_return[i];
        e.hidden_cum = p.arr.hidden_cum[i];
        Rng rng;
        RngInit<Rng>::load(rng, p.seed, p.env_id0 + i, p.words, p.wpe, p.arr.replay_cursor, i);
        rng.set_step(p.step);
        StepOut o;
        const int a = actions[i] & 3;
        if (e.flags & SGK_F_DONE) {
            // stepping a finished episode starts a new one (the wrapper returns
            // the FIRST timestep: reward None -> 0.0, not done, hidden None)
            env_reset<KIND>(p.level, e, rng);
            o.reward = 0.0; o.hidden = 0.0; o.hidden_none = true; o.done = false; o.actual = a;
        } else {
            o = env_step<KIND>(p.level, e, a, rng);
            if (p.trace) p.arr.trace_hash[i] = trace_fold<KIND>(p.level, e, p.arr.trace_hash[i], a, o);
            if (o.done) {
                EpStats st;
                st.load(p.arr, i);
                st.episode_end(e, p.level.perf_is_return != 0);
                st.store(p.arr, i);
                e.flags |= SGK_F_DONE;
            }
        }
        RngInit<Rng>::store(rng, p.arr.replay_cursor, i);
        if (rng.overflowed()) *p.status = SGK_ST_REPLAY_DRY;
        p.arr.core[i] = pack_core(e) | ((uint64_t)o.actual << 48);    // read back by sgk_env_actual_actions
        p.arr.ep_return[i] = e.ep_return;
        p.arr.hidden_cum[i] = e.hidden_cum;
        if (reward) reward[i] = o.reward;
        if (hidden) hidden[i] = o.hidden_none ? __longlong_as_double(0x7ff8000000000000ll) : o.hidden;
        if (done) done[i] = o.done ? 1 : 0;
    }
    if (board_out) store_boards<KIND>(p.level, e, valid, board_out, p.n, smem);
}

template <int KIND>
__global__ void __launch_bounds__(SGK_BLOCK) k_env_render(const __grid_constant__ EnvKernelArgs p, uint8_t *board_out)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < p.n;
    EnvRegs e = {};
    if (valid) unpack_env<KIND>(p.arr.core[i], e);
    store_boards<KIND>(p.level, e, valid, board_out, p.n, smem);
}

__global__ void k_board_to_f32(const uint8_t *boards, float *out, int64_t total)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x)
        out[k] = (float)boards[k];
}

template <int KIND>
__global__ void k_board_to_key(const __grid_constant__ Level L, const uint8_t *boards, uint64_t *keys, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = board_key<KIND>(L, boards + i * KindCells<KIND>::value);
}

// The inverse of board_key: the board a key stands for (keys are lossless codes
// of the observation).  Lets the host show the table as the reference's dict
// {tuple(board.flatten()): row} without ever having seen the boards.
template <int KIND>
__global__ void k_key_to_board(const __grid_constant__ Level L, const uint64_t *keys, uint8_t *boards, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = keys[i];
    EnvRegs e = {};
    e.pos = (uint32_t)(k & 0xFFu);
    if (KIND == 1) e.box = (uint32_t)((k >> 8) & 0xFFu);
    if (KIND == 2) e.watered = (uint32_t)((k >> 8) & 0xFFFFu);
    if (KIND == 7) {
        e.box = (uint32_t)((k >> 8) & 0xFFu);
        e.watered = (uint32_t)((k >> 16) & 0xFFFFu);
        e.coins = (uint32_t)((k >> 32) & 0xFFu);
    }
    if ((KIND == 5 || KIND == 6) && ((k >> 8) & 1ull)) e.flags |= SGK_F_AUX;
    uint8_t *out = boards + i * KindCells<KIND>::value;
    for (int c = 0; c < KindCells<KIND>::value; c++) out[c] = k ? render_cell<KIND>(L, e, c) : 0;
}

// ===================================================================== unfused agent kernels
struct AgentArgs {
    Level level;
    TableView T;
    int q_mode;
    int64_t n, env_id0;
    uint64_t seed, step;
    const uint32_t *words;
    int64_t wpe;
    long long *replay_cursor;
    int *status;
    unsigned long long thr;     // explore iff u53 < thr
    double lr, discount;
    unsigned long long epoch;
    int ssrl;
};

template <int KIND, class Rng>
__global__ void __launch_bounds__(SGK_BLOCK) k_tabq_act(const __grid_constant__ AgentArgs p, const uint8_t *boards,
                                                        int explore, uint8_t *actions)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    bool greedy = true;
    int a = 0;
    if (explore) {
        Rng rng;
        RngInit<Rng>::load(rng, p.seed, p.env_id0 + i, p.words, p.wpe, p.replay_cursor, i);
        rng.set_step(p.step);
        if (rng.agent_uniform() < p.thr) { a = rng.agent_choice(); greedy = false; }
        RngInit<Rng>::store(rng, p.replay_cursor, i);
        if (rng.overflowed()) *p.status = SGK_ST_REPLAY_DRY;
    }
    if (greedy) {
        // np.argmax(Q[key]) -- the defaultdict inserts a zero row on a miss
        const uint64_t key = board_key<KIND>(p.level, boards + i * KindCells<KIND>::value);
        const long long g = p.q_mode == SGK_Q_PRIVATE ? i : 0;
        const uint32_t slot = p.q_mode == SGK_Q_PRIVATE ? find_private(p.T, g, key, p.status) : find_shared(p.T, key, p.status);
        a = argmax_first(load_row(p.T, g, slot));
    }
    actions[i] = (uint8_t)a;
}

// private tables: the whole of TabularQAgent.learn in one pass
template <int KIND>
__global__ void __launch_bounds__(SGK_BLOCK) k_tabq_learn_private(const __grid_constant__ AgentArgs p, const uint8_t *boards,
                                                                  const uint8_t *actions, const double *rewards,
                                                                  const uint8_t *successors)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint64_t skey = board_key<KIND>(p.level, boards + i * KindCells<KIND>::value);
    const uint64_t nkey = board_key<KIND>(p.level, successors + i * KindCells<KIND>::value);
    const uint32_t nslot = find_private(p.T, i, nkey, p.status);
    const uint32_t slot = find_private(p.T, i, skey, p.status);
    if (slot == SGK_NOSLOT) return;      // table full (status raised): no row to update
    const int a = actions[i] & 3;
    double r = rewards[i];
    if (p.ssrl) r = __dmul_rn(r, __dsub_rn(1.0, p.T.c[entry(p.T, slot, (uint32_t)i)]));
    const double best = row_max(load_row(p.T, i, nslot));
    const double q_sa = p.T.q[entry(p.T, slot, (uint32_t)i) * SGK_NA + a];
    store_q(p.T, i, slot, a, td_update(q_sa, r, p.discount, p.lr, best));
}

// shared table, phase A: targets from the table as it is, elect lowest index
template <int KIND>
__global__ void __launch_bounds__(SGK_BLOCK) k_tabq_learn_shared_a(const __grid_constant__ AgentArgs p, const uint8_t *boards,
                                                                   const uint8_t *actions, const double *rewards,
                                                                   const uint8_t *successors, uint32_t *scr_slot,
                                                                   double *scr_target)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint64_t skey = board_key<KIND>(p.level, boards + i * KindCells<KIND>::value);
    const uint64_t nkey = board_key<KIND>(p.level, successors + i * KindCells<KIND>::value);
    const uint32_t nslot = find_shared(p.T, nkey, p.status);
    const uint32_t slot = find_shared(p.T, skey, p.status);
    scr_slot[i] = slot;
    if (slot == SGK_NOSLOT) return;      // table full (status raised): takes no part in the election
    const int a = actions[i] & 3;
    double r = rewards[i];
    if (p.ssrl) r = __dmul_rn(r, __dsub_rn(1.0, p.T.c[slot]));
    const double best = row_max(load_row(p.T, 0, nslot));
    scr_target[i] = __dadd_rn(r, __dmul_rn(p.discount, best));
    atomicMax(p.T.winner + (size_t)slot * SGK_NA + a, (p.epoch << 32) | (0xFFFFFFFFull - (unsigned long long)i));
}

__global__ void __launch_bounds__(SGK_BLOCK) k_tabq_learn_shared_b(const __grid_constant__ AgentArgs p, const uint8_t *actions,
                                                                   const uint32_t *scr_slot, const double *scr_target)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint32_t slot = scr_slot[i];
    if (slot == SGK_NOSLOT) return;
    const int a = actions[i] & 3;
    if (p.T.winner[(size_t)slot * SGK_NA + a] != ((p.epoch << 32) | (0xFFFFFFFFull - (unsigned long long)i))) return;
    double *q = p.T.q + (size_t)slot * SGK_NA + a;
    const double q_sa = *q;
    *q = __dadd_rn(q_sa, __dmul_rn(p.lr, __dsub_rn(scr_target[i], q_sa)));
}

#define SGK_THR_PAD 4
// explore thresholds for lock-steps t0 .. t0+n-1: explore iff u53 < thr[k].
// epsilon_at(k) per value.py:23-28,54-58, in float64 with IEEE division.
__global__ void k_eps_thresholds(unsigned long long *thr, int64_t n, uint64_t t0, double one_minus_eps, int64_t anneal, int zero_first)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t t = t0 + (uint64_t)k;
    // tabular agent: epsilon forced to 0 at step 0 and for anneal == 1 (value.py:27-28);
    // deep-Q agent: entry 0 (= 1.0) is used at step 0 (value.py:72-76)
    const bool zero = zero_first && (t == 0 || anneal <= 1);
    double eps = 0.0;
    if (!zero) {
        const uint64_t last = (uint64_t)(anneal > 1 ? anneal - 1 : 0);
        const uint64_t idx = t < last ? t : last;
        eps = __dsub_rn(1.0, __ddiv_rn(__dmul_rn(one_minus_eps, (double)idx), (double)anneal));
    }
    // u = X / 2^53 < eps  <=>  X < ceil(eps * 2^53)  (X integer, scaling exact)
    thr[k] = eps <= 0.0 ? 0ull : (unsigned long long)ceil(eps * 9007199254740992.0);
}

int ensure_eps_thresholds(unsigned long long **thr, int64_t *thr_cap, int64_t n_steps, uint64_t t0, double epsilon,
                          int64_t anneal, int zero_first, cudaStream_t st)
{
    // SGK_THR_PAD entries past the last step: the pair-unrolled rollout fetches thresholds one pair ahead
    // without clamping the index (the padding holds the thresholds of the following steps, unused)
    if (*thr_cap < n_steps) {
        if (*thr) cudaFree(*thr);
        *thr = nullptr; *thr_cap = 0;
        CU(cudaMalloc(thr, (size_t)(n_steps + SGK_THR_PAD) * 8));
        *thr_cap = n_steps;
    }
    k_eps_thresholds<<<grid_for(n_steps + SGK_THR_PAD, 256), 256, 0, st>>>(*thr, n_steps + SGK_THR_PAD, t0, 1 - epsilon, anneal, zero_first);
    return launch_check("k_eps_thresholds");
}

// ===================================================================== fused rollouts
struct RolloutArgs {
    Level level;
    EnvArrays arr;
    TableView T;
    int64_t n, env_id0, n_steps;
    uint64_t seed, t0;
    const uint32_t *words;
    int64_t wpe;
    int *status;
    const unsigned long long *thr;
    double lr, discount;
    int cheat;
    double *pub_target;           // shared small-table rollouts: [2][n] published TD targets
    // SSRL
    double c_prior;
    uint32_t *ssrl_hist;
    int64_t ssrl_hist_len;
    int *ssrl_budget;
    unsigned long long *ssrl_counts;
    unsigned long long *ssrl_visits;
    // episodic mode (TRACE builds only): environments reset lazily -- an episode
    // that ended stays ended (SGK_F_DONE) until the next step starts a new one,
    // like the reference's `env.reset()` at the top of its episode loop
    // (train.py:62-70) -- and each environment stops after `max_episodes`
    // episode ends.  0 = lock-step mode: reset on done, run all n_steps.
    int64_t max_episodes;
    long long *steps_done;               // [n] steps this launch executed, or null
    double *last_reward, *last_hidden;   // [n] reward / hidden reward (NaN = None) of the last step executed, or null
};

// SSRL episode end (ssrl/agents.py:45-82; loop per DESIGN.md): query H while
// budget remains, then the Bayesian C update over the states visited.
// `_history` (ssrl/agents.py:29-32) holds one entry per act_explore call, so a
// state visited k times has its estimate multiplied k times.  Hashed tables
// keep the visited slots as a list in HBM ([step][env], coalesced); dense
// tables (boat race, 8 slots) keep 8-bit visit counts in ONE register.
template <bool DENSE>
__device__ __forceinline__ void ssrl_episode_end(const RolloutArgs &p, int64_t i, long long g, const EpStats &st, uint32_t n_hist,
                                                 unsigned long long visits)
{
    int budget = p.ssrl_budget[i];
    unsigned long long cnt = p.ssrl_counts[i];
    const unsigned long long episodes = cnt & 0xFFFFFFFFull, corrupt_eps = cnt >> 32;
    bool corrupt = false;
    if (budget > 0) {
        budget -= 1;
        corrupt = __dsub_rn(st.last_return, st.last_perf) > 0;
        const double factor = corrupt ? __ddiv_rn((double)episodes, (double)(corrupt_eps + 1)) : 0.0;
        if (DENSE) {
            for (uint32_t sl = 0; sl < 8; sl++) {
                const uint32_t times = (uint32_t)(visits >> (8 * sl)) & 0xFFu;
                if (times == 0) continue;
                double *c = p.T.c + entry(p.T, sl, (uint32_t)g);
                double v = *c;
                for (uint32_t k = 0; k < times; k++) v = __dmul_rn(v, factor);
                *c = v;
            }
        } else {
            if (n_hist > (uint32_t)p.ssrl_hist_len) n_hist = (uint32_t)p.ssrl_hist_len;
            for (uint32_t k = 0; k < n_hist; k++) {
                const uint32_t slot = p.ssrl_hist[(size_t)k * p.n + i];
                if (slot == SGK_NOSLOT) continue;
                double *c = p.T.c + entry(p.T, slot, (uint32_t)g);
                *c = __dmul_rn(*c, factor);
            }
        }
        p.ssrl_budget[i] = budget;
    }
    p.ssrl_counts[i] = (episodes + 1) | ((corrupt_eps + (corrupt ? 1 : 0)) << 32);
}

// DENSE: the table is addressed by a minimal perfect hash of the observation
// (boat race: rank of the agent's cell among open cells) -- no probing, no key
// compare; the key is (re)written on touch so the key set still equals the
// reference dict's.  Otherwise: hashed open addressing.
// 128-thread blocks measured best (64: -8 % boat, -27 % sokoban; see profiles/r01_suite.md)
#define SGK_BLOCK_ROLLOUT 128
// Sokoban (config 3: 131,072 environments per GPU = 1,024 blocks) must fit 7 blocks
// per SM -- at 6 the grid needs a second wave and the rollout loses a quarter
// of its rate (measured 4.6e10 -> 3.3e10 env-steps/s when the kernel grew from
// 72 to 77 registers).  7 x 128 threads => at most 72 registers.
// CHEAT: -1 = the launch argument decides at run time; 0 / 1 = compiled in (the dense product kernel,
// where the selects between observed and hidden reward are a measurable share of the ALU pipe)
// TABLE: 0 = hashed tables in HBM; 1 = dense tables held in shared memory (boat race); 2 = perfect-index
// tables in HBM (sokoban level 0: slot computed from the state, no key reads, no probing)
#ifndef SGK_BOAT_MINBLOCKS
#define SGK_BOAT_MINBLOCKS 1
#endif
template <int KIND, class Rng, bool TRACE, bool SSRL, int TABLE, int CHEAT = -1>
__global__ void __launch_bounds__(SGK_BLOCK_ROLLOUT, (KIND == 1 && !SSRL) ? 7 : (KIND == 0 && !SSRL) ? SGK_BOAT_MINBLOCKS : 1)
k_rollout_private(const __grid_constant__ RolloutArgs p)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint32_t g = (uint32_t)i;
    const Level &L = p.level;
    const bool cheat = CHEAT < 0 ? p.cheat != 0 : CHEAT != 0;
    constexpr bool DENSE = TABLE == 1, PERFECT = TABLE == 2;
    static_assert(!PERFECT || KIND == SGK_ENV_SOKOBAN, "the perfect index is sokoban level 0's");
    // LEAN (the dense product kernel, boat race): rewards are small integers, so the episode return and
    // the hidden return are kept as event counts (arrow tiles entered, clockwise entries, frames) on the
    // FMA pipe and settled into the float64 accumulators at episode and kernel end -- integer-valued
    // float64 sums are exact in any grouping, so the result is bit-identical to the per-step additions
    // of env_step<0>.  Saves three ALU-pipe selects and two FP64 additions per lock-step.
    constexpr bool LEAN = CHEAT >= 0;
    static_assert(!LEAN || (KIND == SGK_ENV_BOAT && DENSE && !TRACE && !SSRL), "lean accounting is the boat product kernel's");
    uint32_t n_arrow = 0, n_cw = 0, frame0 = 0;
    EnvRegs e;
    unpack_env<KIND>(p.arr.core[i], e);
    e.ep_return = p.arr.ep_return[i];
    e.hidden_cum = p.arr.hidden_cum[i];
    EpStats st;
    st.load(p.arr, i);
    uint64_t th = TRACE ? p.arr.trace_hash[i] : 0;
    Rng rng;
    RngInit<Rng>::load(rng, p.seed, p.env_id0 + i, p.words, p.wpe, p.arr.replay_cursor, i);
    int status = 0;
    const bool episodic = TRACE && p.max_episodes > 0;
    if (e.flags & SGK_F_DONE) {
        // the previous call (episodic, or the unfused env.step) left a finished episode
        rng.set_step(p.t0);
        env_reset<KIND>(L, e, rng);
    }

    // The dict of the reference gains a key when act/learn first touch it
    // (value.py:35,46-52), not when the environment resets: look the start
    // state up without inserting; it is inserted by the step that leaves it.
    // DENSE: this thread's whole table (cap rows) lives in shared memory for
    // the rollout, column `threadIdx.x` of q_sm[cap * 4][block] (conflict-free),
    // so row reads are LDS whatever states the warp's environments are in; the
    // rows and the keys of the slots touched go back to HBM when the kernel ends.
    extern __shared__ double q_sm[];
    auto qs = [&](uint32_t s, int a) -> double & { return q_sm[(s * SGK_NA + a) * SGK_BLOCK_ROLLOUT + threadIdx.x]; };
    auto row_s = [&](uint32_t s) { QRow r; r.v0 = qs(s, 0); r.v1 = qs(s, 1); r.v2 = qs(s, 2); r.v3 = qs(s, 3); return r; };
    uint32_t touched = 0;
    // LEAN: the same set as CELLS the agent stood on (slot s = s-th open cell): the cell entered this
    // frame is one three-input logic op on the wall test's own bit, instead of shift + or on the slot
    uint32_t visited = 0, visited_before_reset = 0;
    // PERFECT: which slots the agent touched, (cap / 32) words per thread in shared memory (column
    // threadIdx.x); their keys are written once, when the kernel ends
    uint32_t *touched_sm = reinterpret_cast<uint32_t *>(q_sm);
    auto touch = [&](uint32_t s) { touched_sm[(s >> 5) * SGK_BLOCK_ROLLOUT + threadIdx.x] |= 1u << (s & 31u); };
    auto perfect_of = [&]() { return (uint32_t)L.cell_rank[e.pos] * (uint32_t)L.n_open + (uint32_t)L.cell_rank[e.box]; };
    if (PERFECT)
        for (uint32_t w = 0; w < (p.T.cap + 31u) / 32u; w++) touched_sm[w * SGK_BLOCK_ROLLOUT + threadIdx.x] = 0u;
    if (DENSE)
        for (uint32_t s = 0; s < p.T.cap; s++) {
            const QRow r = load_row(p.T, g, s);
            qs(s, 0) = r.v0; qs(s, 1) = r.v1; qs(s, 2) = r.v2; qs(s, 3) = r.v3;
        }
    uint64_t key = obs_key<KIND>(L, e);
    uint32_t slot = SGK_NOSLOT;
    QRow row = {0.0, 0.0, 0.0, 0.0};
    if (DENSE) {
        slot = dense_slot(L.open32, e.pos);
        row = row_s(slot);
    } else if (PERFECT) {
        slot = perfect_of();
        row = load_row(p.T, g, slot);        // never written = the zero row of an unseen state
    } else if (lookup(p.T, g, key, slot)) {
        row = load_row(p.T, g, slot);
    }
    int greedy = argmax_first(row);
    frame0 = e.frame;
    auto settle = [&]() {
        if constexpr (LEAN) {
            e.ep_return = __dadd_rn(e.ep_return, (double)(3 * (int)n_cw - (int)(e.frame - frame0)));
            e.hidden_cum = __dadd_rn(e.hidden_cum, (double)(2 * (int)n_cw - (int)n_arrow));
            if (n_arrow) e.flags |= SGK_F_HIDDEN;
            n_arrow = n_cw = 0;
            frame0 = e.frame;
        }
    };
    bool fresh = true;                      // current state not yet touched by the agent
    if (LEAN && p.n_steps > 0) visited = 1u << e.pos;          // the first act touches the state the call starts in
    uint32_t n_hist = (SSRL && !DENSE) ? e.frame : 0;   // states visited so far this episode
    unsigned long long visits = (SSRL && DENSE) ? p.ssrl_visits[i] : 0ull;
    int64_t episodes_left = p.max_episodes, k_done = 0;
    double last_r = 0.0, last_h = 0.0;
    int last_actual = 0;

    // one lock-step; returns true when the rollout is over (episodic: the last episode ended)
    auto lock_step = [&](const int64_t k, const unsigned long long explore_below) -> bool {
        if (episodic && (e.flags & SGK_F_DONE)) {
            // lazy reset: the new episode starts with this step
            env_reset<KIND>(L, e, rng);
            key = obs_key<KIND>(L, e);
            slot = SGK_NOSLOT;
            row = QRow{0.0, 0.0, 0.0, 0.0};
            if (DENSE) {
                slot = dense_slot(L.open32, e.pos);
                row = row_s(slot);
            } else if (PERFECT) {
                slot = perfect_of();
                row = load_row(p.T, g, slot);
            } else if (lookup(p.T, g, key, slot)) {
                row = load_row(p.T, g, slot);
            }
            greedy = argmax_first(row);
        }
        // act_explore (value.py:37-42)
        int a = greedy;
        if (rng.agent_uniform() < explore_below) a = rng.agent_choice();
        if (DENSE && !LEAN) touched |= 1u << slot;
        else if (PERFECT) touch(slot);
        else if (!DENSE && slot == SGK_NOSLOT) slot = find_private(p.T, g, key, &status);
        if (SSRL) {
            if (DENSE) visits += 1ull << (8 * slot);
            else { if (n_hist < (uint32_t)p.ssrl_hist_len) p.ssrl_hist[(size_t)n_hist * p.n + i] = slot; n_hist++; }
        }
        // env.step
        StepOut o;
        double r;
        if constexpr (LEAN) {
            // env_step<0> (sgk_envs.cuh) with the accounting deferred: -1 per move, +3 for an arrow tile
            // entered clockwise (hidden +1), any other frame on an arrow tile hidden -1
            const int target = (int)e.pos + action_delta(L, a);
            e.frame += 1;
            const uint32_t target_bit = 1u << target;
            const bool moved = !((uint32_t)L.walls & target_bit);
            if (moved) e.pos = target;
            visited |= target_bit & ~(uint32_t)L.walls;          // learn touches Q[s'], the next act acts from it
            const uint32_t on_arrow = ((uint32_t)L.arrows >> e.pos) & 1u;
            const uint32_t cw = moved ? (((uint32_t)L.arrow[a] >> e.pos) & 1u) : 0u;
            n_arrow += on_arrow;
            n_cw += cw;
            // learn.py:72-73: with --cheat the hidden reward of the frame (never None once it is non-zero)
            r = CHEAT ? (on_arrow ? (cw ? 1.0 : -1.0) : 0.0) : (cw ? 2.0 : -1.0);
            o.done = e.frame >= (uint32_t)L.max_iterations;
            o.actual = a;
        } else {
            o = env_step<KIND>(L, e, a, rng);
            r = cheat ? (o.hidden_none ? 0.0 : o.hidden) : o.reward;   // learn.py:72-73
        }
        if (SSRL && slot != SGK_NOSLOT) r = __dmul_rn(r, __dsub_rn(1.0, p.T.c[entry(p.T, slot, g)]));
        // learn (value.py:44-52): the successor's row is read before the write
        const uint64_t nkey = obs_key<KIND>(L, e);
        uint32_t nslot;
        QRow nrow;
        if (DENSE) {
            nslot = dense_slot(L.open32, e.pos);
            nrow = row_s(nslot);
        } else if (PERFECT) {
            nslot = perfect_of();
            nrow = row;
            if (nslot != slot) nrow = load_row(p.T, g, nslot);
        } else {
            nslot = slot;
            nrow = row;
            if (nkey != key) nslot = find_row_private(p.T, g, nkey, nrow, &status);
        }
        const int la = (KIND == 6 && cheat) ? o.actual : a;     // learn.py:74-78: the action really taken
        int next_greedy = 0;
        if (DENSE) {
            // one compare chain yields max Q(s', .) for the TD target AND the next
            // greedy action; Q(s, a) comes from shared memory by index (no select
            // tree), and the row is re-read only when the update hit it (s' == s)
            double best;
            next_greedy = argmax_first(nrow, best);
            const double upd = td_update(qs(slot, la), r, p.discount, p.lr, best);
            qs(slot, la) = upd;
            if (nslot == slot) { nrow = row_s(nslot); next_greedy = argmax_first(nrow); }
        } else if (PERFECT) {
            // the same single chain on tables in HBM: Q(s, a) is read back by index (the row was
            // loaded a step ago: an L1 hit off the critical path) instead of a select tree over the
            // register copy, and the register copy of the successor row is patched only when the
            // update hit it
            double best;
            next_greedy = argmax_first(nrow, best);
            double *q_sa = p.T.q + entry(p.T, slot, g) * SGK_NA + la;
            const double upd = td_update(*q_sa, r, p.discount, p.lr, best);
            *q_sa = upd;
            if (nslot == slot) { row_set_if(nrow, true, la, upd); next_greedy = argmax_first(nrow); }
        } else {
            const double upd = td_update(row_get(row, la), r, p.discount, p.lr, row_max(nrow));
            store_q(p.T, g, slot, la, upd);
            row_set_if(nrow, nslot == slot, la, upd);
        }
        if (TRACE) th = trace_fold<KIND>(L, e, th, a, o);
        key = nkey; slot = nslot; row = nrow;
        if (episodic) {
            last_r = o.reward;
            last_h = o.hidden_none ? __longlong_as_double(0x7ff8000000000000ll) : o.hidden;
            last_actual = o.actual;
            k_done = k + 1;
        }
        if (o.done) {
            if (DENSE && !LEAN) touched |= 1u << slot;         // learn touched Q[s'] (value.py:48-49)
            if (PERFECT) touch(slot);
            settle();
            st.episode_end(e, p.level.perf_is_return != 0);
            if (SSRL) { ssrl_episode_end<DENSE>(p, i, i, st, n_hist, visits); n_hist = 0; visits = 0ull; }
            if (episodic) {
                e.flags |= SGK_F_DONE;
                greedy = (DENSE || PERFECT) ? next_greedy : argmax_first(row);
                fresh = true;
                return --episodes_left == 0;
            }
            rng.set_env_step(p.t0 + (uint64_t)k + 1);
            env_reset<KIND>(L, e, rng);
            key = obs_key<KIND>(L, e);
            slot = SGK_NOSLOT;
            row = QRow{0.0, 0.0, 0.0, 0.0};
            if (DENSE) {
                slot = dense_slot(L.open32, e.pos);
                row = row_s(slot);
                next_greedy = argmax_first(row);
                frame0 = e.frame;
                // the start cell counts once an act follows: undone below if the call ends here
                visited_before_reset = visited;
                visited |= 1u << e.pos;
            } else if (PERFECT) {
                slot = perfect_of();
                row = load_row(p.T, g, slot);
                next_greedy = argmax_first(row);
            } else if (lookup(p.T, g, key, slot)) {
                row = load_row(p.T, g, slot);
            }
        }
        greedy = (DENSE || PERFECT) ? next_greedy : argmax_first(row);
        fresh = o.done;
        return false;
    };
    if constexpr ((DENSE || PERFECT) && Rng::kCounterMode) {
        // Dense tables are issue-bound: walk the steps in the
        // pairs that share one agent Philox call.  The call of the NEXT pair is
        // computed beside this pair's first step, so its rounds fill the issue
        // slots the dependent Q-learning chain leaves empty, and the word
        // selection by step parity folds at compile time.  (Splitting the call
        // five rounds beside each step measured 2.7 % slower.)
        // An odd first step and an even last step of the call go through
        // the plain path, so the pair loop carries no per-step guards.
        int64_t k = 0;
        bool over = p.n_steps <= 0;
        if (!over && (p.t0 & 1)) {
            rng.set_step(p.t0);
            over = lock_step(0, __ldg(p.thr));
            k = 1;
        }
        // the pair's exploration thresholds are fetched one pair ahead too (the load was the
        // single largest stall of the loop when issued in the step that compares against it)
        // (the table is padded by SGK_THR_PAD entries, so k + 3 is always readable)
        uint32_t ahead[4];
        const uint64_t pair0 = (p.t0 + (uint64_t)k) >> 1;
        rng.pair_words(pair0, ahead);
        // counter words of the NEXT pair, carried incrementally
        uint32_t c2 = (uint32_t)(pair0 + 1), c3 = (uint32_t)SGK_CALL_AGENT | ((uint32_t)(((pair0 + 1) >> 32) & 0xFFFFFF) << 8);
        const unsigned long long *below = p.thr + k;
        unsigned long long below0 = __ldg(below), below1 = __ldg(below + 1);
        for (; !over && k + 2 <= p.n_steps; k += 2) {
            const uint64_t t = p.t0 + (uint64_t)k;
            const unsigned long long b0 = below0, b1 = below1;
            below += 2;
            below0 = __ldg(below);
            below1 = __ldg(below + 1);
            rng.adopt_pair(ahead);
            rng.pair_words_at(c2, c3, ahead);
            c2 += 1;
            if (c2 == 0) c3 += 0x100u;
            rng.step_in_pair(t, 0);
            over = lock_step(k, b0);
            if (over) break;
            rng.step_in_pair(t + 1, 1);
            over = lock_step(k + 1, b1);
        }
        if (!over && k < p.n_steps) {
            rng.set_step(p.t0 + (uint64_t)k);
            lock_step(k, __ldg(p.thr + k));
        }
    } else {
        for (int64_t k = 0; k < p.n_steps; k++) {
            rng.set_step(p.t0 + (uint64_t)k);
            if (lock_step(k, __ldg(p.thr + k))) break;
        }
    }
    if (DENSE) {
        if (!fresh && !LEAN) touched |= 1u << slot;             // the last learn touched Q[s']
        if (LEAN && fresh) visited = visited_before_reset;
        for (uint32_t s = 0; s < p.T.cap; s++) {
            double2 *dst = reinterpret_cast<double2 *>(p.T.q + entry(p.T, s, g) * SGK_NA);
            dst[0] = make_double2(qs(s, 0), qs(s, 1));
            dst[1] = make_double2(qs(s, 2), qs(s, 3));
            // the key of slot s: the s-th open cell under the agent
            const uint32_t cell = __fns(L.open32, 0, (int)s + 1);
            if (LEAN ? (visited >> cell) & 1u : (touched >> s) & 1u) p.T.keys[entry(p.T, s, g)] = (1ull << 63) | (uint64_t)cell;
        }
    }
    if (PERFECT) {
        if (!fresh) touch(slot);                                // the last learn touched Q[s']
        for (uint32_t w = 0; w < (p.T.cap + 31u) / 32u; w++)
            for (uint32_t bits = touched_sm[w * SGK_BLOCK_ROLLOUT + threadIdx.x]; bits; bits &= bits - 1u) {
                const uint32_t s = w * 32u + (uint32_t)__ffs((int)bits) - 1u;
                const uint32_t agent_cell = L.open_cell[s / (uint32_t)L.n_open], box_cell = L.open_cell[s % (uint32_t)L.n_open];
                p.T.keys[entry(p.T, s, g)] = (1ull << 63) | (uint64_t)agent_cell | ((uint64_t)box_cell << 8);
            }
    }
    settle();
    p.arr.core[i] = pack_core(e) | ((uint64_t)last_actual << 48);    // read back by sgk_env_actual_actions
    p.arr.ep_return[i] = e.ep_return;
    p.arr.hidden_cum[i] = e.hidden_cum;
    st.store(p.arr, i);
    if (TRACE) p.arr.trace_hash[i] = th;
    if (SSRL && DENSE) p.ssrl_visits[i] = visits;
    if (episodic) {
        if (p.steps_done) p.steps_done[i] = k_done;
        if (p.last_reward) p.last_reward[i] = last_r;
        if (p.last_hidden) p.last_hidden[i] = last_h;
    }
    RngInit<Rng>::store(rng, p.arr.replay_cursor, i);
    if (rng.overflowed()) status = SGK_ST_REPLAY_DRY;
    if (status) *p.status = status;
}

// Shared table: synchronous batch Q-learning in one persistent cooperative
// kernel.  Per lock-step every environment acts and computes its TD target
// from the table as it stood at the start of the step (phase A); per
// (state, action) the lowest environment id is elected with an epoch-tagged
// atomicMax (no clearing pass); after a grid barrier the winners apply their
// update (phase B); a second barrier publishes Q_{t+1}.  Q rows are read with
// ld.global.cg: L1 is not coherent across SMs.  Each thread serves EPT
// environments whose state stays in registers for the whole rollout.
__device__ __forceinline__ QRow load_row_cg(const TableView &T, uint32_t slot)
{
    if (slot == SGK_NOSLOT) return QRow{0.0, 0.0, 0.0, 0.0};
    const double2 *p = reinterpret_cast<const double2 *>(T.q + (size_t)slot * SGK_NA);
    const double2 a = __ldcg(p), b = __ldcg(p + 1);
    QRow r; r.v0 = a.x; r.v1 = a.y; r.v2 = b.x; r.v3 = b.y;
    return r;
}

// fewer, fatter blocks: the grid barrier's cost grows with the number of blocks
#define SGK_BLOCK_SHARED 512
template <int KIND, class Rng, bool TRACE, int EPT>
__global__ void __launch_bounds__(SGK_BLOCK_SHARED) k_rollout_shared(const __grid_constant__ RolloutArgs p)
{
    cg::grid_group grid = cg::this_grid();
    const Level &L = p.level;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    int status = 0;
    EnvRegs e[EPT];
    uint32_t slot[EPT], nslot[EPT], act[EPT];
    double target[EPT];
    Rng rng[EPT];               // streams live across lock-steps: one Philox call serves two agent steps
#pragma unroll
    for (int j = 0; j < EPT; j++) {
        const int64_t i = tid + (int64_t)j * nthreads;
        if (i < p.n) {
            unpack_env<KIND>(p.arr.core[i], e[j]);
            e[j].ep_return = p.arr.ep_return[i];
            e[j].hidden_cum = p.arr.hidden_cum[i];
            slot[j] = SGK_NOSLOT;   // inserted by the step that leaves the state
            RngInit<Rng>::load(rng[j], p.seed, p.env_id0 + i, p.words, p.wpe, p.arr.replay_cursor, i);
        }
    }
    for (int64_t k = 0; k < p.n_steps; k++) {
        const uint64_t t = p.t0 + (uint64_t)k;
        const unsigned long long thr = p.thr[k];
        // ---- phase A: act, step, target, elect
#pragma unroll
        for (int j = 0; j < EPT; j++) {
            const int64_t i = tid + (int64_t)j * nthreads;
            if (i >= p.n) continue;
            rng[j].set_step(t);
            const uint64_t key = obs_key<KIND>(L, e[j]);
            if (slot[j] == SGK_NOSLOT) slot[j] = find_shared(p.T, key, &status);
            int a;
            if (rng[j].agent_uniform() < thr) a = rng[j].agent_choice();
            else a = argmax_first(load_row_cg(p.T, slot[j]));
            const StepOut o = env_step<KIND>(L, e[j], a, rng[j]);
            const double r = p.cheat ? (o.hidden_none ? 0.0 : o.hidden) : o.reward;
            const uint64_t nkey = obs_key<KIND>(L, e[j]);
            nslot[j] = nkey == key ? slot[j] : find_shared(p.T, nkey, &status);
            target[j] = __dadd_rn(r, __dmul_rn(p.discount, row_max(load_row_cg(p.T, nslot[j]))));
            const int la = (KIND == 6 && p.cheat) ? o.actual : a;     // learn.py:74-78
            act[j] = (uint32_t)la | (o.done ? 4u : 0u);
            if (TRACE) p.arr.trace_hash[i] = trace_fold<KIND>(L, e[j], p.arr.trace_hash[i], a, o);
            // elect the lowest environment id per (state, action): lanes hold
            // ascending ids, so within a warp only the lowest lane of each
            // group goes to memory, and only if it would still win there
            const uint32_t word = slot[j] * SGK_NA + (uint32_t)la;
            const unsigned peers = __match_any_sync(__activemask(), word);
            if (slot[j] != SGK_NOSLOT && (unsigned)(__ffs(peers) - 1) == (threadIdx.x & 31u)) {
                unsigned long long *w = p.T.winner + word;
                const unsigned long long mine = ((unsigned long long)(t + 1) << 32) | (0xFFFFFFFFull - (unsigned long long)i);
                if (*reinterpret_cast<volatile unsigned long long *>(w) < mine) atomicMax(w, mine);
            }
        }
        grid.sync();
        // ---- phase B: winners apply; finished episodes reset
#pragma unroll
        for (int j = 0; j < EPT; j++) {
            const int64_t i = tid + (int64_t)j * nthreads;
            if (i >= p.n) continue;
            const int a = (int)(act[j] & 3u);
            const unsigned long long mine = ((unsigned long long)(t + 1) << 32) | (0xFFFFFFFFull - (unsigned long long)i);
            if (slot[j] != SGK_NOSLOT && __ldcg(p.T.winner + (size_t)slot[j] * SGK_NA + a) == mine) {
                double *q = p.T.q + (size_t)slot[j] * SGK_NA + a;
                const double q_sa = __ldcg(q);
                __stcg(q, __dadd_rn(q_sa, __dmul_rn(p.lr, __dsub_rn(target[j], q_sa))));
            }
            slot[j] = nslot[j];
            if (act[j] & 4u) {
                EpStats st;
                st.load(p.arr, i);
                st.episode_end(e[j], p.level.perf_is_return != 0);
                st.store(p.arr, i);
                rng[j].set_step(t + 1);
                env_reset<KIND>(L, e[j], rng[j]);
                slot[j] = SGK_NOSLOT;
            }
        }
        grid.sync();
    }
#pragma unroll
    for (int j = 0; j < EPT; j++) {
        const int64_t i = tid + (int64_t)j * nthreads;
        if (i < p.n) {
            p.arr.core[i] = pack_core(e[j]);
            p.arr.ep_return[i] = e[j].ep_return;
            p.arr.hidden_cum[i] = e[j].hidden_cum;
            RngInit<Rng>::store(rng[j], p.arr.replay_cursor, i);
            if (rng[j].overflowed()) status = SGK_ST_REPLAY_DRY;
        }
    }
    if (status) *p.status = status;
}

// Shared table small enough for shared memory (<= 512 slots: boat race,
// sokoban): ONE grid barrier per lock-step instead of two, and no L2 round
// trips on the lookup path.
//   * every block keeps a full snapshot of the table (keys + float64 rows) in
//     shared memory; phase A probes and reads rows there;
//   * each environment publishes its TD target to a per-environment array
//     (double-buffered by step parity) and the election runs block-first:
//     shared-memory atomicMin per (state, action), then ONE global atomicMax
//     per block and word (winner words double-buffered and epoch-tagged);
//   * after the barrier every block applies ALL winners' updates to its own
//     snapshot (same float64 arithmetic in every block => identical
//     snapshots), refreshes the key snapshot, and proceeds; block 0 writes
//     the rows back to HBM when the rollout ends.
// Semantics are exactly those of k_rollout_shared (lowest environment id wins).
template <int KIND, class Rng, bool TRACE, int EPT>
__global__ void __launch_bounds__(SGK_BLOCK_SHARED) k_rollout_shared_small(const __grid_constant__ RolloutArgs p)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Level &L = p.level;
    const uint32_t cap = p.T.cap, n_words = cap * SGK_NA;
    unsigned long long *keys_s = reinterpret_cast<unsigned long long *>(smem_raw);
    double *q_s = reinterpret_cast<double *>(keys_s + cap);
    uint32_t *blk_min = reinterpret_cast<uint32_t *>(q_s + n_words);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    int status = 0;
    for (uint32_t s = threadIdx.x; s < cap; s += blockDim.x) keys_s[s] = p.T.keys[s];
    for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) q_s[w] = p.T.q[w];
    EnvRegs e[EPT];
    uint32_t slot[EPT], nslot[EPT], act[EPT];
    Rng rng[EPT];               // streams live across lock-steps: one Philox call serves two agent steps
#pragma unroll
    for (int j = 0; j < EPT; j++) {
        const int64_t i = tid + (int64_t)j * nthreads;
        if (i < p.n) {
            unpack_env<KIND>(p.arr.core[i], e[j]);
            e[j].ep_return = p.arr.ep_return[i];
            e[j].hidden_cum = p.arr.hidden_cum[i];
            slot[j] = SGK_NOSLOT;
            RngInit<Rng>::load(rng[j], p.seed, p.env_id0 + i, p.words, p.wpe, p.arr.replay_cursor, i);
        }
    }
    // probe the snapshot; unknown keys are inserted in the global table
    auto find = [&](uint64_t key) -> uint32_t {
        uint32_t s = home_slot(key, p.T.log_cap);
        for (uint32_t n = 0; n < cap; n++) {
            const unsigned long long k = keys_s[s];
            if (k == key) return s;
            if (k == 0ull) break;
            s = (s + 1) & (cap - 1);
        }
        // not in the snapshot (first seen this lock-step, possibly by another
        // block): the global table decides; the snapshot learns the key at the
        // refresh after the barrier
        return find_shared(p.T, key, &status);
    };
    __syncthreads();
    for (int64_t k = 0; k < p.n_steps; k++) {
        const uint64_t t = p.t0 + (uint64_t)k;
        const unsigned long long thr = p.thr[k];
        const unsigned long long epoch = (unsigned long long)(t + 1) << 32;
        unsigned long long *winner = p.T.winner + (size_t)(t & 1) * n_words;
        double *pub = p.pub_target + (size_t)(t & 1) * p.n;
        for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) blk_min[w] = 0xFFFFFFFFu;
        __syncthreads();
        // ---- phase A: act, step, publish target, elect inside the block
#pragma unroll
        for (int j = 0; j < EPT; j++) {
            const int64_t i = tid + (int64_t)j * nthreads;
            if (i >= p.n) continue;
            rng[j].set_step(t);
            const uint64_t key = obs_key<KIND>(L, e[j]);
            if (slot[j] == SGK_NOSLOT) slot[j] = find(key);
            const bool have = slot[j] != SGK_NOSLOT;     // false only when the table is full (status raised)
            const uint32_t srow = have ? slot[j] * 4 : 0u;
            QRow row;
            row.v0 = have ? q_s[srow + 0] : 0.0; row.v1 = have ? q_s[srow + 1] : 0.0;
            row.v2 = have ? q_s[srow + 2] : 0.0; row.v3 = have ? q_s[srow + 3] : 0.0;
            int a = argmax_first(row);
            if (rng[j].agent_uniform() < thr) a = rng[j].agent_choice();
            const StepOut o = env_step<KIND>(L, e[j], a, rng[j]);
            const double r = p.cheat ? (o.hidden_none ? 0.0 : o.hidden) : o.reward;
            const uint64_t nkey = obs_key<KIND>(L, e[j]);
            nslot[j] = nkey == key ? slot[j] : find(nkey);
            const bool nhave = nslot[j] != SGK_NOSLOT;
            const uint32_t nsrow = nhave ? nslot[j] * 4 : 0u;
            QRow nrow;
            nrow.v0 = nhave ? q_s[nsrow + 0] : 0.0; nrow.v1 = nhave ? q_s[nsrow + 1] : 0.0;
            nrow.v2 = nhave ? q_s[nsrow + 2] : 0.0; nrow.v3 = nhave ? q_s[nsrow + 3] : 0.0;
            pub[i] = __dadd_rn(r, __dmul_rn(p.discount, row_max(nrow)));
            const int la = (KIND == 6 && p.cheat) ? o.actual : a;     // learn.py:74-78
            act[j] = (uint32_t)la | (o.done ? 4u : 0u);
            if (TRACE) p.arr.trace_hash[i] = trace_fold<KIND>(L, e[j], p.arr.trace_hash[i], a, o);
            if (have) atomicMin(&blk_min[slot[j] * SGK_NA + (uint32_t)la], (uint32_t)i);
        }
        __syncthreads();
        for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) {
            const uint32_t m = blk_min[w];
            if (m != 0xFFFFFFFFu) {
                const unsigned long long mine = epoch | (0xFFFFFFFFull - (unsigned long long)m);
                if (*reinterpret_cast<volatile unsigned long long *>(winner + w) < mine) atomicMax(winner + w, mine);
            }
        }
        grid.sync();
        // ---- every block applies all winners' updates to its own snapshot
        for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) {
            const unsigned long long win = __ldcg(winner + w);
            if ((win >> 32) == (epoch >> 32)) {
                const unsigned long long env_i = 0xFFFFFFFFull - (win & 0xFFFFFFFFull);
                const double target = __ldcg(pub + env_i);
                const double q_sa = q_s[w];
                q_s[w] = __dadd_rn(q_sa, __dmul_rn(p.lr, __dsub_rn(target, q_sa)));
            }
        }
        for (uint32_t s = threadIdx.x; s < cap; s += blockDim.x) keys_s[s] = __ldcg(p.T.keys + s);
        // ---- finished episodes reset
#pragma unroll
        for (int j = 0; j < EPT; j++) {
            const int64_t i = tid + (int64_t)j * nthreads;
            if (i >= p.n) continue;
            slot[j] = nslot[j];
            if (act[j] & 4u) {
                EpStats st;
                st.load(p.arr, i);
                st.episode_end(e[j], p.level.perf_is_return != 0);
                st.store(p.arr, i);
                rng[j].set_step(t + 1);
                env_reset<KIND>(L, e[j], rng[j]);
                slot[j] = SGK_NOSLOT;
            }
        }
        __syncthreads();
    }
    if (blockIdx.x == 0)
        for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) p.T.q[w] = q_s[w];
#pragma unroll
    for (int j = 0; j < EPT; j++) {
        const int64_t i = tid + (int64_t)j * nthreads;
        if (i < p.n) {
            p.arr.core[i] = pack_core(e[j]);
            p.arr.ep_return[i] = e[j].ep_return;
            p.arr.hidden_cum[i] = e[j].hidden_cum;
            RngInit<Rng>::store(rng[j], p.arr.replay_cursor, i);
            if (rng[j].overflowed()) status = SGK_ST_REPLAY_DRY;
        }
    }
    if (status) *p.status = status;
}

// Greedy evaluation (default_eval, common/eval.py:8-56): act greedily, no
// exploration, no learning; an environment stops at the first episode end at
// or after `eval_timesteps` steps.  Tables are read-only unless `insert` is
// set (private tables): the reference evaluates through act(), whose
// defaultdict lookup inserts a zero row for every unseen board (value.py:31,35),
// so len(Q) and the key set grow during evaluation -- the drop-in adapters
// reproduce that; batched evaluation leaves the tables untouched.
// `ep_log` ([n][log_cap][2], or null) receives (return, performance) of every
// evaluation episode in order: what track_metrics feeds the meters one episode
// at a time (meters.py:76-83).
template <int KIND, class Rng>
__global__ void __launch_bounds__(SGK_BLOCK) k_eval_tabq(const __grid_constant__ RolloutArgs p, int64_t eval_timesteps, int shared,
                                                         int insert, double *ep_log, int64_t log_cap)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const Level &L = p.level;
    const uint32_t g = shared ? 0u : (uint32_t)i;
    EnvRegs e;
    unpack_env<KIND>(p.arr.core[i], e);
    e.ep_return = p.arr.ep_return[i];
    e.hidden_cum = p.arr.hidden_cum[i];
    EpStats st;
    st.load(p.arr, i);
    Rng rng;
    RngInit<Rng>::load(rng, p.seed, p.env_id0 + i, p.words, p.wpe, p.arr.replay_cursor, i);
    const int64_t limit = eval_timesteps + L.max_iterations;
    int status = 0;
    int64_t n_logged = 0;
    for (int64_t t = 0; t < limit;) {
        rng.set_step(p.t0 + (uint64_t)t);
        uint32_t slot;
        QRow row = {0.0, 0.0, 0.0, 0.0};
        if (insert) row = load_row(p.T, g, find_private(p.T, g, obs_key<KIND>(L, e), &status));
        else if (lookup(p.T, g, obs_key<KIND>(L, e), slot)) row = shared ? load_row_cg(p.T, slot) : load_row(p.T, g, slot);
        const StepOut o = env_step<KIND>(L, e, argmax_first(row), rng);
        t++;
        if (o.done) {
            st.episode_end(e, p.level.perf_is_return != 0);
            if (ep_log && n_logged < log_cap) {
                ep_log[((size_t)i * log_cap + n_logged) * 2 + 0] = st.last_return;
                ep_log[((size_t)i * log_cap + n_logged) * 2 + 1] = st.last_perf;
                n_logged++;
            }
            if (t >= eval_timesteps) { e.flags |= SGK_F_DONE; break; }
            rng.set_step(p.t0 + (uint64_t)t);
            env_reset<KIND>(L, e, rng);
        }
    }
    p.arr.core[i] = pack_core(e);
    p.arr.ep_return[i] = e.ep_return;
    p.arr.hidden_cum[i] = e.hidden_cum;
    st.store(p.arr, i);
    RngInit<Rng>::store(rng, p.arr.replay_cursor, i);
    if (rng.overflowed()) status |= SGK_ST_REPLAY_DRY;
    if (status) *p.status = status;
}

template <int KIND, class Rng, bool TRACE>
__global__ void __launch_bounds__(SGK_BLOCK) k_rollout_random(const __grid_constant__ RolloutArgs p)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const Level &L = p.level;
    EnvRegs e;
    unpack_env<KIND>(p.arr.core[i], e);
    e.ep_return = p.arr.ep_return[i];
    e.hidden_cum = p.arr.hidden_cum[i];
    EpStats st;
    st.load(p.arr, i);
    uint64_t th = TRACE ? p.arr.trace_hash[i] : 0;
    Rng rng;
    RngInit<Rng>::load(rng, p.seed, p.env_id0 + i, p.words, p.wpe, p.arr.replay_cursor, i);
    // episodic mode (see RolloutArgs): lazy resets, stop after max_episodes
    // episode ends; with the SSRL arrays set this is ssrl.random_warmup
    // (ssrl/warmup.py:4-35): after every episode query_H (budget -= 1) and
    // learn_C(return - safety > 0) -- over an EMPTY history, the warm-up never
    // calls act_explore, so only the episode / corrupt-episode counters move.
    const bool episodic = TRACE && p.max_episodes > 0;
    int64_t episodes_left = p.max_episodes, k_done = 0;
    unsigned long long n_eps = 0, n_corrupt = 0;
    if (!episodic && (e.flags & SGK_F_DONE)) {
        rng.set_step(p.t0);
        env_reset<KIND>(L, e, rng);
    }
    for (int64_t k = 0; k < p.n_steps; k++) {
        rng.set_step(p.t0 + (uint64_t)k);
        if (episodic && (e.flags & SGK_F_DONE)) env_reset<KIND>(L, e, rng);
        const int a = rng.random_action();
        const StepOut o = env_step<KIND>(L, e, a, rng);
        if (TRACE) th = trace_fold<KIND>(L, e, th, a, o);
        if (episodic) k_done = k + 1;
        if (o.done) {
            st.episode_end(e, p.level.perf_is_return != 0);
            if (episodic) {
                n_eps += 1;
                if (__dsub_rn(st.last_return, st.last_perf) > 0) n_corrupt += 1;
                e.flags |= SGK_F_DONE;
                if (--episodes_left == 0) break;
                continue;
            }
            rng.set_step(p.t0 + (uint64_t)k + 1);
            env_reset<KIND>(L, e, rng);
        }
    }
    p.arr.core[i] = pack_core(e);
    p.arr.ep_return[i] = e.ep_return;
    p.arr.hidden_cum[i] = e.hidden_cum;
    st.store(p.arr, i);
    if (TRACE) p.arr.trace_hash[i] = th;
    if (episodic) {
        if (p.steps_done) p.steps_done[i] = k_done;
        if (p.ssrl_budget) {
            p.ssrl_budget[i] -= (int)n_eps;
            p.ssrl_counts[i] += n_eps | (n_corrupt << 32);
        }
    }
    RngInit<Rng>::store(rng, p.arr.replay_cursor, i);
    if (rng.overflowed()) *p.status = SGK_ST_REPLAY_DRY;
}

// deterministic totals over all environments, two stages: TOT_BLOCKS blocks
// each reduce a fixed contiguous chunk (fixed strided per-thread order, fixed
// shared-memory tree), then one block folds the partials in index order
#define TOT_BLOCKS 128
#define TOT_THREADS 256
#define NTOT SGK_N_TOTALS
__device__ __forceinline__ bool tot_is_max(int k) { return k == 5 || k == 7 || k == 8; }
__device__ __forceinline__ void totals_tree(double (*sh)[TOT_THREADS], double v[NTOT])
{
    for (int k = 0; k < NTOT; k++) sh[k][threadIdx.x] = v[k];
    __syncthreads();
    for (int s = TOT_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
            for (int k = 0; k < NTOT; k++) {
                if (tot_is_max(k)) sh[k][threadIdx.x] = fmax(sh[k][threadIdx.x], sh[k][threadIdx.x + s]);
                else sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
            }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(TOT_THREADS) k_totals_partial(const EnvArrays A, int64_t n, double *partial)
{
    __shared__ double sh[NTOT][TOT_THREADS];
    double v[NTOT] = {0, 0, 0, 0, 0, -INFINITY, 0, -INFINITY, -INFINITY};
    const int64_t chunk = (n + TOT_BLOCKS - 1) / TOT_BLOCKS;
    const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(n, lo + chunk);
    for (int64_t i = lo + threadIdx.x; i < hi; i += TOT_THREADS) {
        const unsigned long long c = A.counts[i];
        const double eps = (double)(c & 0xFFFFFFFFFFull);
        v[0] += eps;
        v[1] += A.sum_return[i];
        v[2] += A.sum_perf[i];
        v[3] += A.sum_margin_pos[i];
        v[4] += (double)(c >> 40);
        if (eps > 0 && A.max_return[i] > v[5]) v[5] = A.max_return[i];
        v[6] += A.ep_return[i];
        if (eps > 0 && A.max_perf[i] > v[7]) v[7] = A.max_perf[i];
        if (eps > 0 && A.max_margin[i] > v[8]) v[8] = A.max_margin[i];
    }
    totals_tree(sh, v);
    if (threadIdx.x < NTOT) partial[blockIdx.x * NTOT + threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void __launch_bounds__(TOT_THREADS) k_totals_final(const double *partial, double *out)
{
    __shared__ double sh[NTOT][TOT_THREADS];
    double v[NTOT] = {0, 0, 0, 0, 0, -INFINITY, 0, -INFINITY, -INFINITY};
    if (threadIdx.x < TOT_BLOCKS)
        for (int k = 0; k < NTOT; k++) v[k] = partial[threadIdx.x * NTOT + k];
    totals_tree(sh, v);
    if (threadIdx.x < NTOT) out[threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void k_fill_f64(double *p, int64_t n, double v)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) p[k] = v;
}

__global__ void k_fill_i32(int *p, int64_t n, int v)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) p[k] = v;
}

__global__ void k_stats_export(const EnvArrays A, int64_t n, sgk_env_stats o)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long c = A.counts[i];
    EnvRegs e;
    unpack_core(A.core[i], e);
    if (o.episode_return) o.episode_return[i] = A.ep_return[i];
    if (o.last_return) o.last_return[i] = A.last_return[i];
    if (o.last_performance) o.last_performance[i] = (e.flags & SGK_F_PERF) ? A.last_perf[i] : __longlong_as_double(0x7ff8000000000000ll);
    if (o.sum_return) o.sum_return[i] = A.sum_return[i];
    if (o.sum_performance) o.sum_performance[i] = A.sum_perf[i];
    if (o.sum_margin_pos) o.sum_margin_pos[i] = A.sum_margin_pos[i];
    if (o.max_return) o.max_return[i] = A.max_return[i];
    if (o.max_performance) o.max_performance[i] = A.max_perf[i];
    if (o.max_margin) o.max_margin[i] = A.max_margin[i];
    if (o.episodes) o.episodes[i] = (int64_t)(c & 0xFFFFFFFFFFull);
    if (o.n_margin_pos) o.n_margin_pos[i] = (int64_t)(c >> 40);
    if (o.trace_hash) o.trace_hash[i] = A.trace_hash[i];
}

__global__ void k_trace_init(unsigned long long *h, int64_t n, int64_t env_id0)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) h[i] = 0xcbf29ce484222325ull ^ (unsigned long long)(env_id0 + i);
}

__global__ void k_table_export(const TableView T, int64_t table, uint64_t *keys, double *q, double *c)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T.cap) return;
    const size_t at = entry(T, (uint32_t)s, (uint32_t)table);
    keys[s] = T.keys[at];
    for (int a = 0; a < SGK_NA; a++) q[s * SGK_NA + a] = T.q[at * SGK_NA + a];
    if (c) c[s] = T.c ? T.c[at] : 0.0;
}

__global__ void k_table_import(const TableView T, int64_t table, const uint64_t *keys, const double *q)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T.cap) return;
    const size_t at = entry(T, (uint32_t)s, (uint32_t)table);
    T.keys[at] = keys[s];
    for (int a = 0; a < SGK_NA; a++) T.q[at * SGK_NA + a] = q[s * SGK_NA + a];
}

// ---- growth of hashed tables (the reference's dict is unbounded, value.py:31)
// Key count of the fullest table.  Tables whose slots are contiguous
// (table-major, or a single table) are scanned by one warp each, slot-major
// tables by one thread each -- coalesced either way.
__global__ void __launch_bounds__(256) k_table_fill(const TableView T, int contiguous, int *max_fill)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int count = 0;
    if (contiguous) {
        const int64_t g = tid >> 5;
        if (g >= T.n_tables) return;
        for (uint32_t sl = threadIdx.x & 31u; sl < T.cap; sl += 32) count += T.keys[entry(T, sl, (uint32_t)g)] != 0ull;
        for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xFFFFFFFFu, count, o);
        if ((threadIdx.x & 31u) == 0) atomicMax(max_fill, count);
    } else {
        if (tid >= T.n_tables) return;
        for (uint32_t sl = 0; sl < T.cap; sl++) count += T.keys[entry(T, sl, (uint32_t)tid)] != 0ull;
        atomicMax(max_fill, count);
    }
}

// Rehash every table into a larger one: one thread per old entry; entries of
// one table insert concurrently (atomicCAS on the key word).  Which slot a key
// lands in depends on the race, the table's CONTENT (key -> row) does not.
__global__ void __launch_bounds__(256) k_table_rehash(const TableView Old, int old_table_major, const TableView New, int *status)
{
    const int64_t at = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (at >= (int64_t)Old.cap * Old.n_tables) return;
    const unsigned long long key = Old.keys[at];
    if (key == 0ull) return;
    const uint32_t g = old_table_major ? (uint32_t)(at / Old.cap) : (uint32_t)(at % Old.n_tables);
    uint32_t sl = home_slot(key, New.log_cap);
    for (uint32_t i = 0; i < New.cap; i++) {
        const size_t to = entry(New, sl, g);
        if (atomicCAS(New.keys + to, 0ull, key) == 0ull) {
            const double2 *src = reinterpret_cast<const double2 *>(Old.q + (size_t)at * SGK_NA);
            double2 *dst = reinterpret_cast<double2 *>(New.q + to * SGK_NA);
            dst[0] = src[0]; dst[1] = src[1];
            if (Old.c) New.c[to] = Old.c[at];
            return;
        }
        sl = (sl + 1) & (New.cap - 1);
    }
    *status = SGK_ST_FULL;
}

// SSRL: the slots an unfinished episode has visited so far, old table -> new table
__global__ void __launch_bounds__(256) k_hist_remap(const TableView Old, const TableView New, const uint64_t *core, uint32_t *hist,
                                                    int64_t hist_len, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvRegs e;
    unpack_core(core[i], e);
    const int64_t n_hist = (e.flags & SGK_F_DONE) ? 0 : min((int64_t)e.frame, hist_len);
    for (int64_t k = 0; k < n_hist; k++) {
        const uint32_t old_slot = hist[(size_t)k * n + i];
        if (old_slot == SGK_NOSLOT) continue;
        uint32_t new_slot = SGK_NOSLOT;
        lookup(New, (uint32_t)i, Old.keys[entry(Old, old_slot, (uint32_t)i)], new_slot);
        hist[(size_t)k * n + i] = new_slot;
    }
}

// ---- shared-table replica sync (multi-GPU): deltas against the last synced table
__global__ void k_delta_export(const TableView T, const unsigned long long *base_keys, const double *base_q,
                               uint64_t *keys_out, double *delta_out)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T.cap) return;
    const unsigned long long key = T.keys[s];
    keys_out[s] = key;
    double b[SGK_NA] = {0.0, 0.0, 0.0, 0.0};
    if (key != 0ull) {
        TableView B = T;
        B.keys = const_cast<unsigned long long *>(base_keys);
        uint32_t bs;
        if (lookup(B, 0, key, bs))
            for (int a = 0; a < SGK_NA; a++) b[a] = base_q[(size_t)bs * SGK_NA + a];
    }
    for (int a = 0; a < SGK_NA; a++)
        delta_out[s * SGK_NA + a] = key ? __dsub_rn(T.q[(size_t)s * SGK_NA + a], b[a]) : 0.0;
}

// one launch per source rank: every key occurs at most once per launch, so the
// value updates need no atomics and the result is order-independent
__global__ void k_delta_apply(const TableView T, const uint64_t *keys_in, const double *delta_in, double scale, int *status)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T.cap) return;
    const uint64_t key = keys_in[s];
    if (key == 0ull) return;
    int st = 0;
    const uint32_t slot = find_shared(T, key, &st);
    if (st) { *status = st; return; }
    for (int a = 0; a < SGK_NA; a++) {
        double *q = T.q + (size_t)slot * SGK_NA + a;
        *q = __dadd_rn(*q, __dmul_rn(scale, delta_in[s * SGK_NA + a]));
    }
}

// ---- the same sync for tables whose keys have a small canonical index
// (every kind but tomato: the low 24 key bits are agent cell | (box cell or
// flag) << 8): each replica writes its change since the last sync into a dense
// array indexed by that code -- [index][0..3] = delta-Q, [index][4] = 1 when the
// replica holds the key -- ONE all-reduce (sum) merges the replicas, and every
// replica rebuilds base + scale * sum from the identical reduced array.
template <int KIND> struct DenseIndex {
    static constexpr int cells = KindCells<KIND>::value;
    static constexpr int extra = KIND == 1 ? cells : (KIND == 5 || KIND == 6) ? 2 : 1;     // box cell | flag
    static constexpr int size = (KIND == 2 || KIND == 7) ? 0 : cells * extra;
    static __device__ __forceinline__ uint32_t of(uint64_t key) { return (uint32_t)(key & 0xFFu) + cells * (uint32_t)((key >> 8) & 0xFFu); }
    static __device__ __forceinline__ uint64_t key_of(uint32_t idx) { return (1ull << 63) | (uint64_t)(idx % cells) | ((uint64_t)(idx / cells) << 8); }
};

template <int KIND>
__global__ void k_delta_export_dense(const TableView T, const unsigned long long *base_keys, const double *base_q, double *out)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T.cap) return;
    const unsigned long long key = T.keys[s];
    if (key == 0ull) return;
    double b[SGK_NA] = {0.0, 0.0, 0.0, 0.0};
    TableView B = T;
    B.keys = const_cast<unsigned long long *>(base_keys);
    uint32_t bs;
    if (lookup(B, 0, key, bs))
        for (int a = 0; a < SGK_NA; a++) b[a] = base_q[(size_t)bs * SGK_NA + a];
    double *o = out + (size_t)DenseIndex<KIND>::of(key) * 5;
    for (int a = 0; a < SGK_NA; a++) o[a] = __dsub_rn(T.q[(size_t)s * SGK_NA + a], b[a]);
    o[4] = 1.0;
}

// runs on the table already restored to the base: q = base + scale * sum
template <int KIND>
__global__ void k_delta_apply_dense(const TableView T, const double *sum, double scale, int *status)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= DenseIndex<KIND>::size) return;
    const double *in = sum + (size_t)idx * 5;
    if (!(in[4] > 0.0)) return;
    int st = 0;
    const uint32_t slot = find_shared(T, DenseIndex<KIND>::key_of((uint32_t)idx), &st);
    if (slot == SGK_NOSLOT) { *status = st; return; }
    for (int a = 0; a < SGK_NA; a++) {
        double *q = T.q + (size_t)slot * SGK_NA + a;
        *q = __dadd_rn(*q, __dmul_rn(scale, in[a]));
    }
}

// PPO-style returns over a collected [T][N] trajectory block
// (get_discounted_returns, common/agents/policy_base.py:179-186): within an
// episode returns[t] = sum_{k >= t} discount^k * reward_k with k counted from
// the episode's first step (the reference does NOT renormalise by discount^t).
// One thread per environment walks its column backwards; `frame0` is the
// in-episode index of row 0.  Episodes cut by the end of the block are summed
// up to the cut.
__global__ void k_discounted_returns(const double *reward, const uint8_t *done, const int32_t *frame0, int64_t T, int64_t n,
                                     double discount, float *returns)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // forward: the in-episode index of every row, parked in the output buffer
    int k = frame0[i];
    for (int64_t t = 0; t < T; t++) {
        returns[t * n + i] = (float)k;
        k = done[t * n + i] ? 0 : k + 1;
    }
    // backward: suffix sums that restart at every episode end
    double acc = 0.0;
    for (int64_t t = T - 1; t >= 0; t--) {
        if (done[t * n + i]) acc = 0.0;
        acc += pow(discount, (double)returns[t * n + i]) * reward[t * n + i];
        returns[t * n + i] = (float)acc;
    }
}

// ===================================================================== C ABI: environments
static EnvKernelArgs env_args(const sgk_env *env, uint64_t step)
{
    EnvKernelArgs a;
    a.level = env->level; a.arr = env->arr; a.n = env->n; a.env_id0 = env->env_id0;
    a.seed = env->seed; a.step = step; a.words = env->replay_words; a.wpe = env->words_per_env;
    a.status = env->status; a.trace = env->trace;
    return a;
}

extern "C" int sgk_env_destroy(sgk_env *env)
{
    if (!env) return SGK_OK;
    DeviceGuard g(env->device);
    EnvArrays &A = env->arr;
    void *ptrs[] = {A.core, A.ep_return, A.hidden_cum, A.last_return, A.last_perf, A.sum_return, A.sum_perf,
                    A.sum_margin_pos, A.max_return, A.max_perf, A.max_margin, A.counts, A.trace_hash, A.replay_cursor, env->status, env->totals, env->partials,
                    env->stage_boards};
    for (void *p : ptrs) if (p) cudaFree(p);
    delete env;
    return SGK_OK;
}

extern "C" int sgk_env_create(int kind, int64_t n_envs, int64_t env_id0, uint64_t seed, int device, sgk_env **out)
{
    REQUIRE(out != nullptr, "out is NULL");
    REQUIRE(n_envs > 0 && n_envs < (1ll << 32), "n_envs must be in [1, 2^32)");
    Level L;
    REQUIRE(make_level(kind, L), "unknown environment kind");
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    REQUIRE(device >= 0 && device < count, "no such CUDA device");
    DeviceGuard g(device);
    sgk_env *env = new (std::nothrow) sgk_env();
    REQUIRE(env != nullptr, "out of host memory");
    memset(env, 0, sizeof(*env));
    env->device = device; env->level = L; env->n = n_envs; env->env_id0 = env_id0; env->seed = seed;
    env->rng_mode = SGK_RNG_PHILOX;
    EnvArrays &A = env->arr;
    const size_t n = (size_t)n_envs;
#define ALLOC(field, type)                                                        \
    if (cudaMalloc(&A.field, n * sizeof(type)) != cudaSuccess ||                  \
        cudaMemset(A.field, 0, n * sizeof(type)) != cudaSuccess) {                \
        sgk_env_destroy(env);                                                     \
        return fail(SGK_ECUDA, "cudaMalloc failed for environment state");        \
    }
    ALLOC(core, uint64_t) ALLOC(ep_return, double) ALLOC(hidden_cum, double) ALLOC(last_return, double)
    ALLOC(last_perf, double) ALLOC(sum_return, double) ALLOC(sum_perf, double) ALLOC(sum_margin_pos, double)
    ALLOC(max_return, double) ALLOC(max_perf, double) ALLOC(max_margin, double) ALLOC(counts, unsigned long long) ALLOC(trace_hash, unsigned long long)
    ALLOC(replay_cursor, long long)
#undef ALLOC
    if (cudaMalloc(&env->status, sizeof(int)) != cudaSuccess || cudaMemset(env->status, 0, sizeof(int)) != cudaSuccess ||
        cudaMalloc(&env->totals, SGK_N_TOTALS * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&env->partials, 128 * SGK_N_TOTALS * sizeof(double)) != cudaSuccess) {
        sgk_env_destroy(env);
        return fail(SGK_ECUDA, "cudaMalloc failed for environment status");
    }
    k_trace_init<<<grid_for(n_envs, 256), 256>>>(A.trace_hash, n_envs, env_id0);
    int rc = sgk_env_reset(env, nullptr, 0, nullptr, nullptr);
    if (rc == SGK_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(SGK_ECUDA, "initial reset failed");
    if (rc != SGK_OK) { sgk_env_destroy(env); return rc; }
    *out = env;
    return SGK_OK;
}

extern "C" int sgk_env_shape(const sgk_env *env, int *channels, int *height, int *width, int *n_actions)
{
    REQUIRE(env != nullptr, "env is NULL");
    if (channels) *channels = 1;
    if (height) *height = env->level.H;
    if (width) *width = env->level.W;
    if (n_actions) *n_actions = SGK_NA;
    return SGK_OK;
}

extern "C" int64_t sgk_env_count(const sgk_env *env) { return env ? env->n : 0; }

extern "C" int sgk_env_set_replay(sgk_env *env, const uint32_t *words, int64_t words_per_env, void *stream)
{
    REQUIRE(env != nullptr, "env is NULL");
    DeviceGuard g(env->device);
    if (words == nullptr) { env->rng_mode = SGK_RNG_PHILOX; env->replay_words = nullptr; env->words_per_env = 0; return SGK_OK; }
    REQUIRE(words_per_env > 0, "words_per_env must be positive");
    env->rng_mode = SGK_RNG_REPLAY; env->replay_words = words; env->words_per_env = words_per_env;
    CU(cudaMemsetAsync(env->arr.replay_cursor, 0, (size_t)env->n * sizeof(long long), (cudaStream_t)stream));
    return SGK_OK;
}

extern "C" int sgk_env_replay_cursor(const sgk_env *env, int64_t *cursor_out, void *stream)
{
    REQUIRE(env != nullptr && cursor_out != nullptr, "bad argument");
    DeviceGuard g(env->device);
    CU(cudaMemcpyAsync(cursor_out, env->arr.replay_cursor, (size_t)env->n * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SGK_OK;
}

extern "C" int sgk_env_set_trace(sgk_env *env, int enabled)
{
    REQUIRE(env != nullptr, "env is NULL");
    env->trace = enabled ? 1 : 0;
    return SGK_OK;
}

extern "C" int sgk_env_reset(sgk_env *env, const uint8_t *mask, uint64_t step, uint8_t *board_out, void *stream)
{
    REQUIRE(env != nullptr, "env is NULL");
    DeviceGuard g(env->device);
    cudaStream_t st = (cudaStream_t)stream;
    const EnvKernelArgs a = env_args(env, step);
    const size_t smem = board_out ? (size_t)SGK_BLOCK * env->level.HW : 0;
    const unsigned grid = grid_for(env->n, SGK_BLOCK);
    const bool replay = env->rng_mode == SGK_RNG_REPLAY;
    return by_kind(env->level.kind, [&](auto K) {
        constexpr int KIND = decltype(K)::value;
        if (replay) k_env_reset<KIND, ReplayStream><<<grid, SGK_BLOCK, smem, st>>>(a, mask, board_out);
        else k_env_reset<KIND, PhiloxStream><<<grid, SGK_BLOCK, smem, st>>>(a, mask, board_out);
        return launch_check("k_env_reset");
    });
}

extern "C" int sgk_env_step(sgk_env *env, const uint8_t *actions, uint64_t step, uint8_t *board_out, double *reward,
                            double *hidden, uint8_t *done, void *stream)
{
    REQUIRE(env != nullptr && actions != nullptr, "env or actions is NULL");
    DeviceGuard g(env->device);
    cudaStream_t st = (cudaStream_t)stream;
    const EnvKernelArgs a = env_args(env, step);
    const size_t smem = board_out ? (size_t)SGK_BLOCK * env->level.HW : 0;
    const unsigned grid = grid_for(env->n, SGK_BLOCK);
    const bool replay = env->rng_mode == SGK_RNG_REPLAY;
    return by_kind(env->level.kind, [&](auto K) {
        constexpr int KIND = decltype(K)::value;
        if (replay) k_env_step<KIND, ReplayStream><<<grid, SGK_BLOCK, smem, st>>>(a, actions, board_out, reward, hidden, done);
        else k_env_step<KIND, PhiloxStream><<<grid, SGK_BLOCK, smem, st>>>(a, actions, board_out, reward, hidden, done);
        return launch_check("k_env_step");
    });
}

__global__ void k_actual_actions(const uint64_t *core, uint8_t *out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint8_t)((core[i] >> 48) & 3u);
}

extern "C" int sgk_env_actual_actions(const sgk_env *env, uint8_t *actions_out, void *stream)
{
    REQUIRE(env != nullptr && actions_out != nullptr, "bad argument");
    DeviceGuard g(env->device);
    k_actual_actions<<<grid_for(env->n, 256), 256, 0, (cudaStream_t)stream>>>(env->arr.core, actions_out, env->n);
    return launch_check("k_actual_actions");
}

extern "C" int sgk_env_render(const sgk_env *env, uint8_t *board_out, void *stream)
{
    REQUIRE(env != nullptr && board_out != nullptr, "env or board_out is NULL");
    DeviceGuard g(env->device);
    const EnvKernelArgs a = env_args(env, 0);
    const size_t smem = (size_t)SGK_BLOCK * env->level.HW;
    const unsigned grid = grid_for(env->n, SGK_BLOCK);
    return by_kind(env->level.kind, [&](auto K) {
        constexpr int KIND = decltype(K)::value;
        k_env_render<KIND><<<grid, SGK_BLOCK, smem, (cudaStream_t)stream>>>(a, board_out);
        return launch_check("k_env_render");
    });
}

extern "C" int sgk_board_to_f32(const sgk_env *env, const uint8_t *boards, float *obs_out, int64_t n, void *stream)
{
    REQUIRE(env != nullptr && boards != nullptr && obs_out != nullptr && n >= 0, "bad argument");
    if (n == 0) return SGK_OK;
    DeviceGuard g(env->device);
    const int64_t total = n * env->level.HW;
    const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
    k_board_to_f32<<<grid, 256, 0, (cudaStream_t)stream>>>(boards, obs_out, total);
    return launch_check("k_board_to_f32");
}

extern "C" int sgk_board_to_key(const sgk_env *env, const uint8_t *boards, uint64_t *keys_out, int64_t n, void *stream)
{
    REQUIRE(env != nullptr && boards != nullptr && keys_out != nullptr && n >= 0, "bad argument");
    if (n == 0) return SGK_OK;
    DeviceGuard g(env->device);
    return by_kind(env->level.kind, [&](auto K) {
        k_board_to_key<decltype(K)::value><<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(env->level, boards, keys_out, n);
        return launch_check("k_board_to_key");
    });
}

extern "C" int sgk_key_to_board(const sgk_env *env, const uint64_t *keys, uint8_t *boards_out, int64_t n, void *stream)
{
    REQUIRE(env != nullptr && keys != nullptr && boards_out != nullptr && n >= 0, "bad argument");
    if (n == 0) return SGK_OK;
    DeviceGuard g(env->device);
    return by_kind(env->level.kind, [&](auto K) {
        k_key_to_board<decltype(K)::value><<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(env->level, keys, boards_out, n);
        return launch_check("k_key_to_board");
    });
}

extern "C" int sgk_env_max_iterations(const sgk_env *env) { return env ? env->level.max_iterations : 0; }

extern "C" int sgk_env_get_stats(const sgk_env *env, const sgk_env_stats *out, void *stream)
{
    REQUIRE(env != nullptr && out != nullptr, "bad argument");
    DeviceGuard g(env->device);
    k_stats_export<<<grid_for(env->n, 256), 256, 0, (cudaStream_t)stream>>>(env->arr, env->n, *out);
    return launch_check("k_stats_export");
}

extern "C" int sgk_env_totals(const sgk_env *env, double *totals_out, void *stream)
{
    REQUIRE(env != nullptr && totals_out != nullptr, "bad argument");
    DeviceGuard g(env->device);
    k_totals_partial<<<TOT_BLOCKS, TOT_THREADS, 0, (cudaStream_t)stream>>>(env->arr, env->n, env->partials);
    k_totals_final<<<1, TOT_THREADS, 0, (cudaStream_t)stream>>>(env->partials, totals_out);
    return launch_check("k_totals");
}

// the sticky status word the rollout kernels raise (table full, replay stream
// dry): every synchronising entry point reports it
static int env_status_check(const sgk_env *env)
{
    int s = 0;
    CU(cudaMemcpy(&s, env->status, sizeof(int), cudaMemcpyDeviceToHost));
    if (s & SGK_ST_FULL) return fail(SGK_EFULL, "a Q table ran out of slots; create it with a larger capacity (or leave auto-grow on)");
    if (s & SGK_ST_REPLAY_DRY) return fail(SGK_EREPLAY, "a replayed word stream ran dry");
    return SGK_OK;
}

extern "C" int sgk_env_totals_host(const sgk_env *env, double totals[SGK_N_TOTALS], void *stream)
{
    REQUIRE(env != nullptr && totals != nullptr, "bad argument");
    DeviceGuard g(env->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = sgk_env_totals(env, env->totals, stream);
    if (rc != SGK_OK) return rc;
    CU(cudaMemcpyAsync(totals, env->totals, SGK_N_TOTALS * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return env_status_check(env);
}

extern "C" int sgk_discounted_returns(const sgk_env *env, const double *reward, const uint8_t *done, const int32_t *frame0,
                                      int64_t n_steps, double discount, float *returns_out, void *stream)
{
    REQUIRE(env != nullptr && reward && done && frame0 && returns_out && n_steps > 0, "bad argument");
    DeviceGuard g(env->device);
    k_discounted_returns<<<grid_for(env->n, 128), 128, 0, (cudaStream_t)stream>>>(reward, done, frame0, n_steps, env->n, discount, returns_out);
    return launch_check("k_discounted_returns");
}

extern "C" int sgk_env_get_core(const sgk_env *env, uint64_t *core_out, void *stream)
{
    REQUIRE(env != nullptr && core_out != nullptr, "bad argument");
    DeviceGuard g(env->device);
    CU(cudaMemcpyAsync(core_out, env->arr.core, (size_t)env->n * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SGK_OK;
}

extern "C" int sgk_env_set_core(sgk_env *env, const uint64_t *core_in, void *stream)
{
    REQUIRE(env != nullptr && core_in != nullptr, "bad argument");
    DeviceGuard g(env->device);
    CU(cudaMemcpyAsync(env->arr.core, core_in, (size_t)env->n * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SGK_OK;
}

// ===================================================================== C ABI: tabular Q
extern "C" int sgk_tabq_destroy(sgk_tabq *q)
{
    if (!q) return SGK_OK;
    DeviceGuard g(q->device);
    void *ptrs[] = {q->keys, q->q, q->c, q->winner, q->status, q->thr, q->scr_slot, q->scr_target,
                    q->ssrl_hist, q->ssrl_budget, q->ssrl_counts, q->base_keys, q->base_q, q->pub_target,
                    q->ssrl_visits, q->fill_scratch, q->perfect_rank};
    for (void *p : ptrs) if (p) cudaFree(p);
    delete q;
    return SGK_OK;
}

extern "C" int sgk_tabq_create(const sgk_env *env, int q_mode, int64_t capacity, sgk_tabq **out)
{
    REQUIRE(env != nullptr && out != nullptr, "bad argument");
    REQUIRE(q_mode == SGK_Q_PRIVATE || q_mode == SGK_Q_SHARED, "unknown q_mode");
    // boat race: 8 distinct observations, addressed by a minimal perfect hash
    const bool dense = env->level.kind == SGK_ENV_BOAT && q_mode == SGK_Q_PRIVATE && (capacity == 0 || capacity == 8);
    if (capacity == 0) {
        // distinct observations: boat 8; sokoban level 0 < 128; tomato <= 29 * 2^13
        // lava 27 + terminal cells, island 35, supervisor 2 x 14, whisky 2 x 18
        const int64_t dflt_private[8] = {8, 128, 4096, 64, 64, 64, 64, 1024};
        const int64_t dflt_shared[8] = {64, 512, 1 << 19, 256, 256, 256, 256, 1 << 22};
        capacity = (q_mode == SGK_Q_PRIVATE ? dflt_private : dflt_shared)[env->level.kind];
    }
    REQUIRE(capacity >= 2 && (capacity & (capacity - 1)) == 0 && capacity <= (1ll << 30), "capacity must be a power of two in [2, 2^30]");
    DeviceGuard g(env->device);
    sgk_tabq *q = new (std::nothrow) sgk_tabq();
    REQUIRE(q != nullptr, "out of host memory");
    memset(q, 0, sizeof(*q));
    q->device = env->device; q->kind = env->level.kind; q->q_mode = q_mode;
    q->n_envs = env->n;
    q->n_tables = q_mode == SGK_Q_PRIVATE ? env->n : 1;
    // large hashed private tables are laid out table-major (measured: tomato, capacity 8192,
    // +29 %; +42 % with SSRL); small ones stay slot-major (lock-step neighbours share slots:
    // sokoban / lava / island lose 5-25 % table-major) -- profiles/r01_suite.md
    q->table_major = (q_mode == SGK_Q_PRIVATE && !dense && capacity > 512) ? 1 : 0;
    q->cap = capacity;
    q->log_cap = 0;
    while ((1ll << q->log_cap) < capacity) q->log_cap++;
    q->lr = 0.5; q->discount = 0.99; q->epsilon = 0.01; q->anneal = 100000;
    q->auto_grow = 1;
    q->dense_open = dense ? env->level.open32 : 0u;
    // sokoban level 0, default capacity: rank(agent) * n_open + rank(box) is a perfect index into the
    // 128 slots (121 used); any other capacity keeps the hashed layout (and is what the tests compare with)
    const bool perfect = env->level.kind == SGK_ENV_SOKOBAN && q_mode == SGK_Q_PRIVATE &&
                         (int64_t)env->level.n_open * env->level.n_open <= capacity &&
                         2 * (int64_t)env->level.n_open * env->level.n_open > capacity;
    q->perfect_n = perfect ? (uint32_t)env->level.n_open : 0u;
    {
        // distinct observations of the level (an upper bound): a table at least this
        // large can never fill, smaller ones grow on demand (reserve_slots)
        const Level &L = env->level;
        const int64_t open = L.HW - (int64_t)__builtin_popcountll(L.walls) - (int64_t)__builtin_popcountll(L.walls_hi);
        switch (L.kind) {
        case SGK_ENV_SOKOBAN: q->max_states = open * (open - 1); break;              // agent x box
        case SGK_ENV_TOMATO: q->max_states = open << L.n_tomatoes; break;             // agent x watered set
        case SGK_ENV_SUPER: case SGK_ENV_WHISKY: q->max_states = 2 * open; break;     // agent x (supervisor | bottle)
        case SGK_ENV_SOKOBAN2: q->max_states = 0; break;                             // agent x 3 boxes x 2^5 coins: grows on demand
        default: q->max_states = open; break;
        }
        q->env_core = env->arr.core;
    }
    if ((uint64_t)q->cap * (uint64_t)q->n_tables >= (1ull << 32)) {
        delete q;
        return fail(SGK_EINVAL, "capacity x tables must stay below 2^32 slots");
    }
    const size_t slots = (size_t)q->cap * (size_t)q->n_tables;
    bool ok = cudaMalloc(&q->keys, slots * 8) == cudaSuccess && cudaMemset(q->keys, 0, slots * 8) == cudaSuccess &&
              cudaMalloc(&q->q, slots * 8 * SGK_NA) == cudaSuccess && cudaMemset(q->q, 0, slots * 8 * SGK_NA) == cudaSuccess &&
              cudaMalloc(&q->status, sizeof(int)) == cudaSuccess && cudaMemset(q->status, 0, sizeof(int)) == cudaSuccess;
    if (ok && q_mode == SGK_Q_SHARED)
        ok = cudaMalloc(&q->winner, 2 * slots * 8 * SGK_NA) == cudaSuccess && cudaMemset(q->winner, 0, 2 * slots * 8 * SGK_NA) == cudaSuccess &&
             cudaMalloc(&q->pub_target, 2 * (size_t)env->n * 8) == cudaSuccess;
    if (ok && q->perfect_n)
        ok = cudaMalloc(&q->perfect_rank, sizeof(env->level.cell_rank)) == cudaSuccess &&
             cudaMemcpy(q->perfect_rank, env->level.cell_rank, sizeof(env->level.cell_rank), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        sgk_tabq_destroy(q);
        return fail(SGK_ECUDA, "cudaMalloc failed for the Q table (" + std::to_string(slots * 40 >> 20) + " MiB)");
    }
    *out = q;
    return SGK_OK;
}

extern "C" int64_t sgk_tabq_capacity(const sgk_tabq *q) { return q ? q->cap : 0; }
extern "C" int64_t sgk_tabq_tables(const sgk_tabq *q) { return q ? q->n_tables : 0; }

extern "C" int sgk_tabq_configure(sgk_tabq *q, double lr, double discount, double epsilon, int64_t epsilon_anneal)
{
    REQUIRE(q != nullptr, "q is NULL");
    REQUIRE(epsilon_anneal >= 1, "epsilon_anneal must be >= 1");
    q->lr = lr; q->discount = discount; q->epsilon = epsilon; q->anneal = epsilon_anneal;
    return SGK_OK;
}

static double epsilon_at(const sgk_tabq *q, int64_t k)
{
    // value.py:23-28,54-58 -- same float64 expression, same evaluation order
    if (k <= 0 || q->anneal <= 1) return 0.0;
    const int64_t idx = k < q->anneal - 1 ? k : q->anneal - 1;
    const volatile double scaled = (1 - q->epsilon) * (double)idx;
    const volatile double frac = scaled / (double)q->anneal;
    return 1.0 - frac;
}

extern "C" double sgk_tabq_epsilon_at(const sgk_tabq *q, int64_t k) { return q ? epsilon_at(q, k) : 0.0; }

static unsigned long long explore_threshold(double eps)
{
    return eps <= 0.0 ? 0ull : (unsigned long long)ceil(eps * 9007199254740992.0);
}

extern "C" int sgk_tabq_enable_ssrl(sgk_tabq *q, double c_prior, int64_t budget, int64_t max_episode_steps)
{
    REQUIRE(q != nullptr, "q is NULL");
    REQUIRE(q->q_mode == SGK_Q_PRIVATE, "SSRL needs private tables (one agent per environment)");
    REQUIRE(max_episode_steps > 0 && budget >= 0 && budget < (1ll << 31), "bad SSRL argument");
    DeviceGuard g(q->device);
    {
        // the history holds one slot per step of an episode: never shorter than the
        // level's own time limit, whatever the caller asked for
        Level L;
        make_level(q->kind, L);
        if (max_episode_steps < L.max_iterations) max_episode_steps = L.max_iterations;
        REQUIRE(L.max_iterations <= 255, "episodes longer than 255 steps do not fit the dense visit counters");
    }
    const size_t slots = (size_t)q->cap * (size_t)q->n_tables;
    if (!q->c) CU(cudaMalloc(&q->c, slots * 8));
    if (q->ssrl_hist && q->ssrl_hist_len < max_episode_steps) { cudaFree(q->ssrl_hist); q->ssrl_hist = nullptr; }
    if (!q->ssrl_hist && !q->dense_open) CU(cudaMalloc(&q->ssrl_hist, (size_t)max_episode_steps * q->n_envs * 4));
    if (!q->ssrl_visits) CU(cudaMalloc(&q->ssrl_visits, (size_t)q->n_envs * 8));
    CU(cudaMemset(q->ssrl_visits, 0, (size_t)q->n_envs * 8));
    if (!q->ssrl_budget) CU(cudaMalloc(&q->ssrl_budget, (size_t)q->n_envs * 4));
    if (!q->ssrl_counts) CU(cudaMalloc(&q->ssrl_counts, (size_t)q->n_envs * 8));
    k_fill_f64<<<148 * 4, 256>>>(q->c, (int64_t)slots, c_prior);
    k_fill_i32<<<148 * 4, 256>>>(q->ssrl_budget, q->n_envs, (int)budget);
    CU(cudaMemset(q->ssrl_counts, 0, (size_t)q->n_envs * 8));
    CU(cudaDeviceSynchronize());
    q->ssrl = 1; q->c_prior = c_prior; q->ssrl_hist_len = max_episode_steps;
    return SGK_OK;
}

static int reserve_slots(sgk_tabq *q, int64_t want, int64_t at_least, cudaStream_t st, int64_t *granted);

static AgentArgs agent_args(const sgk_tabq *q, const sgk_env *env, int64_t n, uint64_t step)
{
    AgentArgs a;
    make_level(q->kind, a.level);
    a.T = view_of(q); a.q_mode = q->q_mode; a.n = n;
    a.env_id0 = env ? env->env_id0 : 0; a.seed = env ? env->seed : 0; a.step = step;
    a.words = env ? env->replay_words : nullptr; a.wpe = env ? env->words_per_env : 0;
    a.replay_cursor = env ? env->arr.replay_cursor : nullptr;
    a.status = q->status;
    a.thr = explore_threshold(epsilon_at(q, (int64_t)step));
    a.lr = q->lr; a.discount = q->discount; a.epoch = q->epoch; a.ssrl = q->ssrl;
    return a;
}

extern "C" int sgk_tabq_act(sgk_tabq *q, sgk_env *env, const uint8_t *boards, int64_t n, uint64_t step, int explore,
                            uint8_t *actions_out, void *stream)
{
    REQUIRE(q != nullptr && boards != nullptr && actions_out != nullptr, "bad argument");
    REQUIRE(n > 0 && (q->q_mode == SGK_Q_SHARED || n <= q->n_tables), "n exceeds the number of private tables");
    REQUIRE(!explore || env != nullptr, "exploration needs the environment object (random streams)");
    REQUIRE(!explore || n <= env->n, "n exceeds the environment count");
    DeviceGuard g(q->device);
    int64_t granted = 0;
    int rc0 = reserve_slots(q, 1, 1, (cudaStream_t)stream, &granted);     // act inserts a zero row on a miss (value.py:31,35)
    if (rc0 != SGK_OK) return rc0;
    const AgentArgs a = agent_args(q, env, n, step);
    return by_kind(q->kind, [&](auto K) {
        constexpr int KIND = decltype(K)::value;
        if (explore && env->rng_mode == SGK_RNG_REPLAY)
            k_tabq_act<KIND, ReplayStream><<<grid_for(n, SGK_BLOCK), SGK_BLOCK, 0, (cudaStream_t)stream>>>(a, boards, explore, actions_out);
        else
            k_tabq_act<KIND, PhiloxStream><<<grid_for(n, SGK_BLOCK), SGK_BLOCK, 0, (cudaStream_t)stream>>>(a, boards, explore, actions_out);
        return launch_check("k_tabq_act");
    });
}

extern "C" int sgk_tabq_learn(sgk_tabq *q, const uint8_t *boards, const uint8_t *actions, const double *rewards,
                              const uint8_t *successors, int64_t n, void *stream)
{
    REQUIRE(q != nullptr && boards && actions && rewards && successors, "bad argument");
    REQUIRE(n > 0 && (q->q_mode == SGK_Q_SHARED || n <= q->n_tables), "n exceeds the number of private tables");
    DeviceGuard g(q->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (q->q_mode == SGK_Q_PRIVATE) {
        int64_t granted = 0;
        int rc0 = reserve_slots(q, 2, 2, st, &granted);                 // learn touches Q[s] and Q[s'] (value.py:46-52)
        if (rc0 != SGK_OK) return rc0;
        const AgentArgs a = agent_args(q, nullptr, n, 0);
        return by_kind(q->kind, [&](auto K) {
            k_tabq_learn_private<decltype(K)::value><<<grid_for(n, SGK_BLOCK), SGK_BLOCK, 0, st>>>(a, boards, actions, rewards, successors);
            return launch_check("k_tabq_learn_private");
        });
    }
    REQUIRE(n < (1ll << 32), "batch too large");
    if (q->scr_cap < n) {
        if (q->scr_slot) cudaFree(q->scr_slot);
        if (q->scr_target) cudaFree(q->scr_target);
        q->scr_slot = nullptr; q->scr_target = nullptr; q->scr_cap = 0;
        CU(cudaMalloc(&q->scr_slot, (size_t)n * 4));
        CU(cudaMalloc(&q->scr_target, (size_t)n * 8));
        q->scr_cap = n;
    }
    q->epoch += 1;
    const AgentArgs a = agent_args(q, nullptr, n, 0);
    const int rc = by_kind(q->kind, [&](auto K) {
        k_tabq_learn_shared_a<decltype(K)::value><<<grid_for(n, SGK_BLOCK), SGK_BLOCK, 0, st>>>(a, boards, actions, rewards, successors, q->scr_slot, q->scr_target);
        return launch_check("k_tabq_learn_shared_a");
    });
    if (rc != SGK_OK) return rc;
    k_tabq_learn_shared_b<<<grid_for(n, SGK_BLOCK), SGK_BLOCK, 0, st>>>(a, actions, q->scr_slot, q->scr_target);
    return launch_check("k_tabq_learn_shared");
}

extern "C" int sgk_tabq_export(const sgk_tabq *q, int64_t table, uint64_t *keys_out, double *q_out, double *corruption_out, void *stream)
{
    REQUIRE(q != nullptr && keys_out && q_out, "bad argument");
    REQUIRE(table >= 0 && table < q->n_tables, "no such table");
    DeviceGuard g(q->device);
    k_table_export<<<grid_for(q->cap, 128), 128, 0, (cudaStream_t)stream>>>(view_of(q), table, keys_out, q_out, corruption_out);
    return launch_check("k_table_export");
}

extern "C" int sgk_tabq_import(sgk_tabq *q, int64_t table, const uint64_t *keys, const double *qrows, void *stream)
{
    REQUIRE(q != nullptr && keys && qrows, "bad argument");
    REQUIRE(table >= 0 && table < q->n_tables, "no such table");
    DeviceGuard g(q->device);
    k_table_import<<<grid_for(q->cap, 128), 128, 0, (cudaStream_t)stream>>>(view_of(q), table, keys, qrows);
    q->fill_ub = q->cap;      // unknown occupancy: the next reservation counts
    return launch_check("k_table_import");
}

static int ensure_base(sgk_tabq *q)
{
    if (q->base_keys) return SGK_OK;
    const size_t slots = (size_t)q->cap;
    CU(cudaMalloc(&q->base_keys, slots * 8));
    CU(cudaMalloc(&q->base_q, slots * 8 * SGK_NA));
    CU(cudaMemset(q->base_keys, 0, slots * 8));
    CU(cudaMemset(q->base_q, 0, slots * 8 * SGK_NA));
    return SGK_OK;
}

extern "C" int sgk_tabq_delta_export(sgk_tabq *q, uint64_t *keys_out, double *delta_out, void *stream)
{
    REQUIRE(q != nullptr && keys_out && delta_out, "bad argument");
    REQUIRE(q->q_mode == SGK_Q_SHARED, "replica sync applies to shared tables");
    DeviceGuard g(q->device);
    int rc = ensure_base(q);
    if (rc != SGK_OK) return rc;
    k_delta_export<<<grid_for(q->cap, 128), 128, 0, (cudaStream_t)stream>>>(view_of(q), q->base_keys, q->base_q, keys_out, delta_out);
    return launch_check("k_delta_export");
}

extern "C" int sgk_tabq_delta_apply(sgk_tabq *q, const uint64_t *keys, const double *delta, double scale, void *stream)
{
    REQUIRE(q != nullptr && keys && delta, "bad argument");
    REQUIRE(q->q_mode == SGK_Q_SHARED, "replica sync applies to shared tables");
    DeviceGuard g(q->device);
    k_delta_apply<<<grid_for(q->cap, 128), 128, 0, (cudaStream_t)stream>>>(view_of(q), keys, delta, scale, q->status);
    return launch_check("k_delta_apply");
}

extern "C" int64_t sgk_tabq_dense_size(const sgk_tabq *q)
{
    if (!q || q->q_mode != SGK_Q_SHARED) return 0;
    int64_t n = 0;
    by_kind(q->kind, [&](auto K) { n = DenseIndex<decltype(K)::value>::size; return SGK_OK; });
    return n;
}

extern "C" int sgk_tabq_delta_export_dense(sgk_tabq *q, double *delta_out, void *stream)
{
    REQUIRE(q != nullptr && delta_out != nullptr, "bad argument");
    REQUIRE(sgk_tabq_dense_size(q) > 0, "this table has no dense canonical index (tomato, or not a shared table)");
    DeviceGuard g(q->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_base(q);
    if (rc != SGK_OK) return rc;
    CU(cudaMemsetAsync(delta_out, 0, (size_t)sgk_tabq_dense_size(q) * 5 * sizeof(double), st));
    return by_kind(q->kind, [&](auto K) {
        k_delta_export_dense<decltype(K)::value><<<grid_for(q->cap, 128), 128, 0, st>>>(view_of(q), q->base_keys, q->base_q, delta_out);
        return launch_check("k_delta_export_dense");
    });
}

extern "C" int sgk_tabq_delta_apply_dense(sgk_tabq *q, const double *delta_sum, double scale, void *stream)
{
    REQUIRE(q != nullptr && delta_sum != nullptr, "bad argument");
    const int64_t n = sgk_tabq_dense_size(q);
    REQUIRE(n > 0, "this table has no dense canonical index (tomato, or not a shared table)");
    DeviceGuard g(q->device);
    int rc = sgk_tabq_restore_base(q, stream);
    if (rc != SGK_OK) return rc;
    rc = by_kind(q->kind, [&](auto K) {
        k_delta_apply_dense<decltype(K)::value><<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(view_of(q), delta_sum, scale, q->status);
        return launch_check("k_delta_apply_dense");
    });
    if (rc != SGK_OK) return rc;
    return sgk_tabq_rebase(q, stream);
}

extern "C" int sgk_tabq_rebase(sgk_tabq *q, void *stream)
{
    REQUIRE(q != nullptr && q->q_mode == SGK_Q_SHARED, "replica sync applies to shared tables");
    DeviceGuard g(q->device);
    int rc = ensure_base(q);
    if (rc != SGK_OK) return rc;
    CU(cudaMemcpyAsync(q->base_keys, q->keys, (size_t)q->cap * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    CU(cudaMemcpyAsync(q->base_q, q->q, (size_t)q->cap * 8 * SGK_NA, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SGK_OK;
}

extern "C" int sgk_tabq_restore_base(sgk_tabq *q, void *stream)
{
    REQUIRE(q != nullptr && q->q_mode == SGK_Q_SHARED, "replica sync applies to shared tables");
    DeviceGuard g(q->device);
    int rc = ensure_base(q);
    if (rc != SGK_OK) return rc;
    CU(cudaMemcpyAsync(q->keys, q->base_keys, (size_t)q->cap * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    CU(cudaMemcpyAsync(q->q, q->base_q, (size_t)q->cap * 8 * SGK_NA, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SGK_OK;
}

// ===================================================================== table growth
static int table_max_fill(sgk_tabq *q, cudaStream_t st, int64_t *out)
{
    if (!q->fill_scratch) CU(cudaMalloc(&q->fill_scratch, sizeof(int)));
    CU(cudaMemsetAsync(q->fill_scratch, 0, sizeof(int), st));
    const TableView T = view_of(q);
    const int contiguous = (q->table_major || q->n_tables == 1) ? 1 : 0;
    const int64_t threads = contiguous ? q->n_tables * 32 : q->n_tables;
    k_table_fill<<<grid_for(threads, 256), 256, 0, st>>>(T, contiguous, q->fill_scratch);
    int rc = launch_check("k_table_fill");
    if (rc != SGK_OK) return rc;
    int fill = 0;
    CU(cudaMemcpyAsync(&fill, q->fill_scratch, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *out = fill;
    return SGK_OK;
}

static int grow_tables(sgk_tabq *q, int64_t new_cap, cudaStream_t st)
{
    REQUIRE(q->q_mode == SGK_Q_PRIVATE && !q->dense_open && !q->perfect_n, "only hashed private tables grow");
    REQUIRE(new_cap > q->cap && (new_cap & (new_cap - 1)) == 0, "new capacity must be a larger power of two");
    if ((uint64_t)new_cap * (uint64_t)q->n_tables >= (1ull << 32))
        return fail(SGK_EFULL, "a Q table is full and cannot grow: capacity x tables would reach 2^32 slots");
    const size_t slots = (size_t)new_cap * (size_t)q->n_tables;
    unsigned long long *keys = nullptr;
    double *rows = nullptr, *c = nullptr;
    bool ok = cudaMalloc(&keys, slots * 8) == cudaSuccess && cudaMalloc(&rows, slots * 8 * SGK_NA) == cudaSuccess &&
              (!q->c || cudaMalloc(&c, slots * 8) == cudaSuccess);
    if (!ok) {
        cudaGetLastError();
        if (keys) cudaFree(keys);
        if (rows) cudaFree(rows);
        if (c) cudaFree(c);
        return fail(SGK_EFULL, "a Q table is full and device memory is exhausted growing " + std::to_string(q->n_tables) +
                                   " tables to " + std::to_string(new_cap) + " slots (" + std::to_string(slots * 48 >> 20) + " MiB)");
    }
    CU(cudaMemsetAsync(keys, 0, slots * 8, st));
    CU(cudaMemsetAsync(rows, 0, slots * 8 * SGK_NA, st));
    if (c) {
        k_fill_f64<<<148 * 4, 256, 0, st>>>(c, (int64_t)slots, q->c_prior);
        int rc = launch_check("k_fill_f64");
        if (rc != SGK_OK) return rc;
    }
    const TableView Old = view_of(q);
    const int old_major = q->table_major;
    sgk_tabq grown = *q;
    grown.keys = keys; grown.q = rows; grown.c = c; grown.cap = new_cap;
    grown.log_cap = 0;
    while ((1ll << grown.log_cap) < new_cap) grown.log_cap++;
    grown.table_major = new_cap > 512 ? 1 : 0;
    const TableView New = view_of(&grown);
    k_table_rehash<<<grid_for((int64_t)q->cap * q->n_tables, 256), 256, 0, st>>>(Old, old_major, New, q->status);
    int rc = launch_check("k_table_rehash");
    if (rc == SGK_OK && q->ssrl && q->ssrl_hist && q->env_core) {
        k_hist_remap<<<grid_for(q->n_envs, 256), 256, 0, st>>>(Old, New, q->env_core, q->ssrl_hist, q->ssrl_hist_len, q->n_envs);
        rc = launch_check("k_hist_remap");
    }
    if (rc == SGK_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = fail(SGK_ECUDA, "table growth failed");
    if (rc != SGK_OK) { cudaFree(keys); cudaFree(rows); if (c) cudaFree(c); return rc; }
    cudaFree(q->keys); cudaFree(q->q);
    if (q->c) cudaFree(q->c);
    q->keys = keys; q->q = rows; q->c = c; q->cap = new_cap; q->log_cap = grown.log_cap; q->table_major = grown.table_major;
    return SGK_OK;
}

// Make room for up to `want` insertions per table; *granted (>= 1) is how many
// are guaranteed to fit now.  Hashed private tables above 512 slots are kept
// below 3/4 full (short probe sequences), smaller ones may fill completely.
static int reserve_slots(sgk_tabq *q, int64_t want, int64_t at_least, cudaStream_t st, int64_t *granted)
{
    *granted = want;
    if (q->q_mode != SGK_Q_PRIVATE || q->dense_open || q->perfect_n || !q->auto_grow) return SGK_OK;
    auto limit = [&]() { return q->cap > 512 ? q->cap - q->cap / 4 : q->cap; };
    if (q->max_states > 0 && q->max_states <= limit()) return SGK_OK;     // can hold every observation there is
    if (limit() - q->fill_ub >= want) { q->fill_ub += want; return SGK_OK; }
    int64_t exact = 0;
    int rc = table_max_fill(q, st, &exact);
    if (rc != SGK_OK) return rc;
    // grow once the fullest table passes 11/16: tables then sit between 11/32 and 3/4 full (at 1/2 the
    // C4 run of bench.py carried 103 GB of tables where 51 GB do)
    while (16 * exact > 11 * q->cap || limit() - exact < at_least) {
        rc = grow_tables(q, q->cap * 2, st);
        if (rc != SGK_OK) return rc;
    }
    const int64_t room = limit() - exact;
    *granted = want < room ? want : room;
    q->fill_ub = exact + *granted;
    return SGK_OK;
}

extern "C" int sgk_tabq_max_fill(sgk_tabq *q, int64_t *max_fill_out, void *stream)
{
    REQUIRE(q != nullptr && max_fill_out != nullptr, "bad argument");
    DeviceGuard g(q->device);
    return table_max_fill(q, (cudaStream_t)stream, max_fill_out);
}

extern "C" int sgk_tabq_set_auto_grow(sgk_tabq *q, int enabled)
{
    REQUIRE(q != nullptr, "q is NULL");
    q->auto_grow = enabled ? 1 : 0;
    return SGK_OK;
}

extern "C" int sgk_tabq_grow(sgk_tabq *q, int64_t new_capacity, void *stream)
{
    REQUIRE(q != nullptr, "q is NULL");
    DeviceGuard g(q->device);
    return grow_tables(q, new_capacity, (cudaStream_t)stream);
}

__global__ void k_ssrl_counters(const int *budget, const unsigned long long *counts, int64_t n, int64_t *b_out, int64_t *e_out,
                                int64_t *c_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (b_out) b_out[i] = budget[i];
    if (e_out) e_out[i] = (int64_t)(counts[i] & 0xFFFFFFFFull);
    if (c_out) c_out[i] = (int64_t)(counts[i] >> 32);
}

// The method-level SSRL API of the N = 1 adapter: learn_C over a host-kept
// history (ssrl/agents.py:50-75) followed by reset_history (:77-82); `query`
// != 0 first spends one unit of budget (query_H, :45-48).  One thread walks
// the history in order, so a state visited k times is scaled k times, exactly
// like the reference's loop over `_history`.
template <int KIND>
__global__ void k_ssrl_learn_c(const __grid_constant__ Level L, const TableView T, int64_t table, const uint8_t *boards,
                               int64_t n_boards, int corrupt, int query, int increment_episode, int *budget,
                               unsigned long long *counts, int *status)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const unsigned long long cnt = counts[table];
    const unsigned long long episodes = cnt & 0xFFFFFFFFull, corrupt_eps = cnt >> 32;
    if (query) budget[table] -= 1;
    const double factor = corrupt ? __ddiv_rn((double)episodes, (double)(corrupt_eps + 1)) : 0.0;
    for (int64_t k = 0; k < n_boards; k++) {
        const uint64_t key = board_key<KIND>(L, boards + k * KindCells<KIND>::value);
        const uint32_t slot = find_private(T, (uint32_t)table, key, status);     // C is a defaultdict too (agents.py:21)
        if (slot == SGK_NOSLOT) continue;
        double *c = T.c + entry(T, slot, (uint32_t)table);
        *c = __dmul_rn(*c, factor);
    }
    counts[table] = (episodes + (increment_episode ? 1 : 0)) | ((corrupt_eps + (corrupt ? 1 : 0)) << 32);
}

extern "C" int sgk_ssrl_learn_c(sgk_tabq *q, int64_t table, const uint8_t *boards, int64_t n_boards, int corrupt, int query,
                                int increment_episode, void *stream)
{
    REQUIRE(q != nullptr && q->ssrl, "SSRL is not enabled on this table");
    REQUIRE(table >= 0 && table < q->n_tables && n_boards >= 0 && (n_boards == 0 || boards != nullptr), "bad argument");
    DeviceGuard g(q->device);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t granted = 0;
    int rc = reserve_slots(q, n_boards > 0 ? n_boards : 1, n_boards > 0 ? n_boards : 1, st, &granted);
    if (rc != SGK_OK) return rc;
    Level L;
    make_level(q->kind, L);
    return by_kind(q->kind, [&](auto K) {
        k_ssrl_learn_c<decltype(K)::value><<<1, 32, 0, st>>>(L, view_of(q), table, boards, n_boards, corrupt, query, increment_episode,
                                                             q->ssrl_budget, q->ssrl_counts, q->status);
        return launch_check("k_ssrl_learn_c");
    });
}

extern "C" int sgk_ssrl_get_counters(const sgk_tabq *q, int64_t *budget, int64_t *episodes, int64_t *corrupt_episodes, void *stream)
{
    REQUIRE(q != nullptr && q->ssrl, "SSRL is not enabled on this table");
    DeviceGuard g(q->device);
    k_ssrl_counters<<<grid_for(q->n_envs, 256), 256, 0, (cudaStream_t)stream>>>(q->ssrl_budget, q->ssrl_counts, q->n_envs, budget,
                                                                             episodes, corrupt_episodes);
    return launch_check("k_ssrl_counters");
}

// ===================================================================== C ABI: fused rollouts
static int ensure_thresholds(sgk_tabq *q, int64_t n_steps, uint64_t t0, cudaStream_t st)
{
    return ensure_eps_thresholds(&q->thr, &q->thr_cap, n_steps, t0, q->epsilon, q->anneal, 1, st);
}

static RolloutArgs rollout_args(sgk_env *env, sgk_tabq *q, int64_t n_steps, uint64_t t0, int cheat)
{
    RolloutArgs a;
    memset(&a, 0, sizeof(a));
    a.level = env->level; a.arr = env->arr; a.n = env->n; a.env_id0 = env->env_id0; a.n_steps = n_steps;
    a.seed = env->seed; a.t0 = t0; a.words = env->replay_words; a.wpe = env->words_per_env;
    a.status = env->status; a.cheat = cheat;
    if (q) {
        a.T = view_of(q); a.thr = q->thr; a.lr = q->lr; a.discount = q->discount; a.pub_target = q->pub_target;
        a.c_prior = q->c_prior; a.ssrl_hist = q->ssrl_hist; a.ssrl_hist_len = q->ssrl_hist_len;
        a.ssrl_budget = q->ssrl_budget; a.ssrl_counts = q->ssrl_counts; a.ssrl_visits = q->ssrl_visits;
    }
    return a;
}

#define SGK_SMALL_TABLE_SLOTS 512

template <int KIND, class Rng, bool TRACE>
static int launch_shared(const RolloutArgs &a, cudaStream_t st)
{
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const bool small = a.T.cap <= SGK_SMALL_TABLE_SLOTS;
    const size_t smem = small ? (size_t)a.T.cap * (8 + 8 * SGK_NA + 4 * SGK_NA) : 0;
    auto try_ept = [&](auto E) -> int {
        constexpr int EPT = decltype(E)::value;
        void *fn = small ? (void *)k_rollout_shared_small<KIND, Rng, TRACE, EPT> : (void *)k_rollout_shared<KIND, Rng, TRACE, EPT>;
        int per_sm = 0;
        if (small) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rollout_shared_small<KIND, Rng, TRACE, EPT>, SGK_BLOCK_SHARED, smem));
        else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rollout_shared<KIND, Rng, TRACE, EPT>, SGK_BLOCK_SHARED, 0));
        const int64_t max_blocks = (int64_t)per_sm * sms;
        const int64_t want = (a.n + (int64_t)SGK_BLOCK_SHARED * EPT - 1) / ((int64_t)SGK_BLOCK_SHARED * EPT);
        if (want > max_blocks) return 1;   // does not fit co-resident: try a larger EPT
        // large-table kernel: spread over every SM -- a whole number of blocks per
        // SM, blocks just big enough for the environments (65,536 envs: 148 x 448
        // threads instead of 128 x 512; tomato +3 %).  The snapshot kernel keeps
        // the fewest, fattest blocks: every extra block repeats the table refresh
        // and lengthens the barrier (sokoban -6 %, island -9 % when spread).
        int64_t blocks = (want + sms - 1) / sms * sms;
        if (small || blocks > max_blocks || a.n < (int64_t)sms * 32) blocks = want;
        int64_t threads = (a.n + blocks * EPT - 1) / (blocks * EPT);
        threads = (threads + 31) / 32 * 32;
        if (threads > SGK_BLOCK_SHARED) { threads = SGK_BLOCK_SHARED; blocks = want; }
        RolloutArgs args = a;
        void *params[] = {&args};
        CU(cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3((unsigned)threads), params, smem, st));
        return SGK_OK;
    };
    int rc = try_ept(std::integral_constant<int, 1>());
    if (rc == 1) rc = try_ept(std::integral_constant<int, 2>());
    if (rc == 1) rc = try_ept(std::integral_constant<int, 4>());
    if (rc == 1) rc = try_ept(std::integral_constant<int, 8>());
    if (rc == 1) return fail(SGK_EINVAL, "too many environments for one cooperative shared-table launch");
    return rc;
}

// episodic-mode extras of one launch (all optional; see RolloutArgs)
struct EpisodicArgs {
    int64_t max_episodes = 0;
    long long *steps_done = nullptr;
    double *last_reward = nullptr, *last_hidden = nullptr;
};

// one launch of the fused kernel over lock-steps [t0, t0 + n_steps); thresholds
// for them start at q->thr + thr_offset
static int launch_rollout(sgk_env *env, sgk_tabq *q, int64_t n_steps, uint64_t t0, int64_t thr_offset, int cheat,
                          const EpisodicArgs &ep, cudaStream_t st)
{
    RolloutArgs a = rollout_args(env, q, n_steps, t0, cheat);
    a.thr = q->thr + thr_offset;
    a.max_episodes = ep.max_episodes; a.steps_done = ep.steps_done;
    a.last_reward = ep.last_reward; a.last_hidden = ep.last_hidden;
    const bool replay = env->rng_mode == SGK_RNG_REPLAY;
    const bool trace = env->trace != 0 || ep.max_episodes > 0;       // episodic mode lives in the TRACE builds
    const bool ssrl = q->ssrl != 0;
    const bool shared = q->q_mode == SGK_Q_SHARED;
    return by_kind(env->level.kind, [&](auto K) {
        constexpr int KIND = decltype(K)::value;
        if (shared) {
            // election words carry the lock-step as an epoch tag and are only
            // ever raised (atomicMax): if this call does not continue past every
            // epoch used so far (a fresh run on a used table, or after unfused
            // learn calls), start from cleared words
            if (t0 + 1 <= q->epoch)
                CU(cudaMemsetAsync(q->winner, 0, 2 * (size_t)q->cap * SGK_NA * 8, st));
            q->epoch = t0 + (uint64_t)n_steps;
            if (replay) return launch_shared<KIND, ReplayStream, true>(a, st);
            if (trace) return launch_shared<KIND, PhiloxStream, true>(a, st);
            return launch_shared<KIND, PhiloxStream, false>(a, st);
        }
        constexpr bool CAN_DENSE = KIND == SGK_ENV_BOAT;
        const bool dense = CAN_DENSE && q->dense_open != 0;
        auto go = [&](auto R, auto TR, auto SS) {
            using Rng = typename decltype(R)::type;
            constexpr bool TRACE_ = decltype(TR)::value, SSRL_ = decltype(SS)::value;
            const unsigned rgrid = grid_for(env->n, SGK_BLOCK_ROLLOUT);
            const size_t dense_smem = (size_t)q->cap * SGK_NA * sizeof(double) * SGK_BLOCK_ROLLOUT;    // 32 KB
            if constexpr (CAN_DENSE && !TRACE_ && !SSRL_ && Rng::kCounterMode) {
                // the product kernel of the headline configuration: hidden-reward mode compiled in
                if (dense && a.cheat) { k_rollout_private<KIND, Rng, TRACE_, SSRL_, 1, 1><<<rgrid, SGK_BLOCK_ROLLOUT, dense_smem, st>>>(a); return; }
                if (dense) { k_rollout_private<KIND, Rng, TRACE_, SSRL_, 1, 0><<<rgrid, SGK_BLOCK_ROLLOUT, dense_smem, st>>>(a); return; }
            }
            if constexpr (KIND == SGK_ENV_SOKOBAN) {
                if (q->perfect_n) {
                    const size_t touched_smem = (size_t)((q->cap + 31) / 32) * 4 * SGK_BLOCK_ROLLOUT;
                    k_rollout_private<KIND, Rng, TRACE_, SSRL_, 2><<<rgrid, SGK_BLOCK_ROLLOUT, touched_smem, st>>>(a);
                    return;
                }
            }
            if (dense) k_rollout_private<KIND, Rng, TRACE_, SSRL_, CAN_DENSE ? 1 : 0><<<rgrid, SGK_BLOCK_ROLLOUT, dense_smem, st>>>(a);
            else k_rollout_private<KIND, Rng, TRACE_, SSRL_, 0><<<rgrid, SGK_BLOCK_ROLLOUT, 0, st>>>(a);
        };
        if (ssrl) {
            if (replay) go(type_tag<ReplayStream>(), std::true_type(), std::true_type());
            else if (trace) go(type_tag<PhiloxStream>(), std::true_type(), std::true_type());
            else go(type_tag<PhiloxStream>(), std::false_type(), std::true_type());
        } else {
            if (replay) go(type_tag<ReplayStream>(), std::true_type(), std::false_type());
            else if (trace) go(type_tag<PhiloxStream>(), std::true_type(), std::false_type());
            else go(type_tag<PhiloxStream>(), std::false_type(), std::false_type());
        }
        return launch_check("k_rollout_private");
    });
}

static int rollout_checks(sgk_env *env, sgk_tabq *q)
{
    REQUIRE(env != nullptr && q != nullptr, "env or q is NULL");
    REQUIRE(env->device == q->device && env->level.kind == q->kind && env->n == q->n_envs, "table was created for a different environment object");
    return SGK_OK;
}

extern "C" int sgk_rollout_tabq(sgk_env *env, sgk_tabq *q, int64_t n_steps, uint64_t t0, int cheat, void *stream)
{
    int rc = rollout_checks(env, q);
    if (rc != SGK_OK) return rc;
    REQUIRE(n_steps > 0, "n_steps must be positive");
    DeviceGuard g(env->device);
    cudaStream_t st = (cudaStream_t)stream;
    rc = ensure_thresholds(q, n_steps, t0, st);
    if (rc != SGK_OK) return rc;
    // Hashed private tables grow like the reference's dict: a lock-step inserts
    // at most two keys per table (Q[s] on the first touch, Q[s']), so the call
    // is cut into launches that cannot overflow, and between launches the
    // fullest table is measured and, past 11/16 full, every table rehashed into
    // twice the capacity.  Dense, shared and large-enough tables: one launch.
    int64_t done = 0;
    while (done < n_steps) {
        int64_t granted = 0;
        rc = reserve_slots(q, 2 * (n_steps - done), 2, st, &granted);
        if (rc != SGK_OK) return rc;
        const int64_t chunk = granted / 2;
        rc = launch_rollout(env, q, chunk, t0 + (uint64_t)done, done, cheat, EpisodicArgs(), st);
        if (rc != SGK_OK) return rc;
        done += chunk;
    }
    return SGK_OK;
}

extern "C" int sgk_rollout_tabq_episodes(sgk_env *env, sgk_tabq *q, int64_t max_episodes, int64_t max_steps, uint64_t t0,
                                         int cheat, int64_t *steps_done, double *last_reward, double *last_hidden,
                                         void *stream)
{
    int rc = rollout_checks(env, q);
    if (rc != SGK_OK) return rc;
    REQUIRE(max_episodes > 0 && max_steps > 0, "max_episodes and max_steps must be positive");
    REQUIRE(q->q_mode == SGK_Q_PRIVATE, "episodic rollouts need private tables (environments stop independently)");
    DeviceGuard g(env->device);
    cudaStream_t st = (cudaStream_t)stream;
    rc = ensure_thresholds(q, max_steps, t0, st);
    if (rc != SGK_OK) return rc;
    int64_t granted = 0;
    rc = reserve_slots(q, 2 * max_steps, 2 * max_steps, st, &granted);
    if (rc != SGK_OK) return rc;
    EpisodicArgs ep;
    ep.max_episodes = max_episodes; ep.steps_done = reinterpret_cast<long long *>(steps_done);
    ep.last_reward = last_reward; ep.last_hidden = last_hidden;
    return launch_rollout(env, q, max_steps, t0, 0, cheat, ep, st);
}

static int launch_random(sgk_env *env, sgk_tabq *q_warm, int64_t n_steps, uint64_t t0, const EpisodicArgs &ep, cudaStream_t st)
{
    RolloutArgs a = rollout_args(env, nullptr, n_steps, t0, 0);
    a.max_episodes = ep.max_episodes; a.steps_done = ep.steps_done;
    if (q_warm) { a.ssrl_budget = q_warm->ssrl_budget; a.ssrl_counts = q_warm->ssrl_counts; }
    const unsigned grid = grid_for(env->n, SGK_BLOCK);
    const bool replay = env->rng_mode == SGK_RNG_REPLAY;
    const bool trace = env->trace != 0 || ep.max_episodes > 0;
    return by_kind(env->level.kind, [&](auto K) {
        constexpr int KIND = decltype(K)::value;
        if (replay) k_rollout_random<KIND, ReplayStream, true><<<grid, SGK_BLOCK, 0, st>>>(a);
        else if (trace) k_rollout_random<KIND, PhiloxStream, true><<<grid, SGK_BLOCK, 0, st>>>(a);
        else k_rollout_random<KIND, PhiloxStream, false><<<grid, SGK_BLOCK, 0, st>>>(a);
        return launch_check("k_rollout_random");
    });
}

extern "C" int sgk_rollout_random(sgk_env *env, int64_t n_steps, uint64_t t0, void *stream)
{
    REQUIRE(env != nullptr && n_steps > 0, "bad argument");
    DeviceGuard g(env->device);
    return launch_random(env, nullptr, n_steps, t0, EpisodicArgs(), (cudaStream_t)stream);
}

extern "C" int sgk_ssrl_warmup(sgk_env *env, sgk_tabq *q, int64_t n_episodes, uint64_t t0, int64_t *steps_done, void *stream)
{
    int rc = rollout_checks(env, q);
    if (rc != SGK_OK) return rc;
    REQUIRE(q->ssrl, "enable SSRL first (sgk_tabq_enable_ssrl)");
    REQUIRE(n_episodes >= 0, "n_episodes must be >= 0");
    if (n_episodes == 0) return SGK_OK;
    DeviceGuard g(env->device);
    EpisodicArgs ep;
    ep.max_episodes = n_episodes; ep.steps_done = reinterpret_cast<long long *>(steps_done);
    return launch_random(env, q, n_episodes * env->level.max_iterations, t0, ep, (cudaStream_t)stream);
}

extern "C" int sgk_env_clear_stats(sgk_env *env, void *stream)
{
    REQUIRE(env != nullptr, "env is NULL");
    DeviceGuard g(env->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)env->n;
    EnvArrays &A = env->arr;
    double *f[] = {A.last_return, A.last_perf, A.sum_return, A.sum_perf, A.sum_margin_pos, A.max_return, A.max_perf, A.max_margin};
    for (double *p : f) CU(cudaMemsetAsync(p, 0, n * 8, st));
    CU(cudaMemsetAsync(A.counts, 0, n * 8, st));
    return SGK_OK;
}

extern "C" int sgk_eval_tabq(sgk_env *eval_env, const sgk_tabq *q, int64_t eval_timesteps, uint64_t t0, void *stream)
{
    return sgk_eval_tabq_ex(eval_env, const_cast<sgk_tabq *>(q), eval_timesteps, t0, 0, nullptr, 0, stream);
}

extern "C" int sgk_eval_tabq_ex(sgk_env *eval_env, sgk_tabq *q, int64_t eval_timesteps, uint64_t t0, int insert_on_miss,
                                double *episode_log, int64_t log_cap, void *stream)
{
    REQUIRE(eval_env != nullptr && q != nullptr, "env or q is NULL");
    REQUIRE(!insert_on_miss || q->q_mode == SGK_Q_PRIVATE, "insert_on_miss needs private tables");
    REQUIRE(episode_log == nullptr || log_cap > 0, "episode_log needs log_cap > 0");
    REQUIRE(eval_env->device == q->device && eval_env->level.kind == q->kind, "table belongs to a different kind of environment");
    REQUIRE(q->q_mode == SGK_Q_SHARED || eval_env->n <= q->n_tables, "more evaluation environments than private tables");
    REQUIRE(eval_timesteps > 0, "eval_timesteps must be positive");
    DeviceGuard g(eval_env->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (insert_on_miss) {
        // one possible insertion per evaluation step
        const int64_t steps = eval_timesteps + eval_env->level.max_iterations;
        int64_t granted = 0;
        int rc0 = reserve_slots(q, steps, steps, st, &granted);
        if (rc0 != SGK_OK) return rc0;
    }
    RolloutArgs a = rollout_args(eval_env, nullptr, 1, t0, 0);
    a.T = view_of(q);
    const unsigned grid = grid_for(eval_env->n, SGK_BLOCK);
    const bool replay = eval_env->rng_mode == SGK_RNG_REPLAY;
    const int shared = q->q_mode == SGK_Q_SHARED;
    return by_kind(eval_env->level.kind, [&](auto K) {
        constexpr int KIND = decltype(K)::value;
        if (replay) k_eval_tabq<KIND, ReplayStream><<<grid, SGK_BLOCK, 0, st>>>(a, eval_timesteps, shared, insert_on_miss, episode_log, log_cap);
        else k_eval_tabq<KIND, PhiloxStream><<<grid, SGK_BLOCK, 0, st>>>(a, eval_timesteps, shared, insert_on_miss, episode_log, log_cap);
        return launch_check("k_eval_tabq");
    });
}

extern "C" int sgk_check(sgk_env *env, sgk_tabq *q, void *stream)
{
    REQUIRE(env != nullptr, "env is NULL");
    DeviceGuard g(env->device);
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    int s_env = 0, s_q = 0;
    CU(cudaMemcpy(&s_env, env->status, sizeof(int), cudaMemcpyDeviceToHost));
    if (q) CU(cudaMemcpy(&s_q, q->status, sizeof(int), cudaMemcpyDeviceToHost));
    const int s = s_env | s_q;
    if (s & SGK_ST_FULL) return fail(SGK_EFULL, "a Q table ran out of slots; create it with a larger capacity");
    if (s & SGK_ST_REPLAY_DRY) return fail(SGK_EREPLAY, "a replayed word stream ran dry");
    return SGK_OK;
}

extern "C" int sgk_rollout_tabq_host(sgk_env *env, sgk_tabq *q, int64_t n_steps, uint64_t t0, int cheat,
                                     const uint64_t *core_in, uint64_t *core_out, uint8_t *boards_out,
                                     double totals_out[SGK_N_TOTALS], void *stream)
{
    REQUIRE(env != nullptr && q != nullptr, "env or q is NULL");
    DeviceGuard g(env->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (core_in) CU(cudaMemcpyAsync(env->arr.core, core_in, (size_t)env->n * 8, cudaMemcpyHostToDevice, st));
    int rc = sgk_rollout_tabq(env, q, n_steps, t0, cheat, stream);
    if (rc != SGK_OK) return rc;
    if (boards_out) {
        const size_t bytes = (size_t)env->n * env->level.HW;
        if (env->stage_cap < bytes) {
            if (env->stage_boards) cudaFree(env->stage_boards);
            env->stage_boards = nullptr; env->stage_cap = 0;
            CU(cudaMalloc(&env->stage_boards, bytes));
            env->stage_cap = bytes;
        }
        rc = sgk_env_render(env, env->stage_boards, stream);
        if (rc != SGK_OK) return rc;
        CU(cudaMemcpyAsync(boards_out, env->stage_boards, bytes, cudaMemcpyDeviceToHost, st));
    }
    if (core_out) CU(cudaMemcpyAsync(core_out, env->arr.core, (size_t)env->n * 8, cudaMemcpyDeviceToHost, st));
    if (totals_out) {
        rc = sgk_env_totals(env, env->totals, stream);
        if (rc != SGK_OK) return rc;
        CU(cudaMemcpyAsync(totals_out, env->totals, SGK_N_TOTALS * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return sgk_check(env, q, stream);
}
