// sgk_common.cuh -- level descriptions, packed environment state, Philox.
//
// Design: one THREAD owns one environment.  A gridworld of <= 64
// cells is a handful of 64-bit bitboards (walls, goal, arrows, tomatoes), so
// the whole dynamic state of an environment -- agent cell, box cell, frame,
// 13 tomato bits, flags -- packs into ONE 64-bit word that lives in a register
// for the duration of a fused rollout.  uint8 boards exist only at the API
// boundary (rendered on demand, staged through shared memory for coalesced
// stores).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SGK_NA 4
#define SGK_MAX_CELLS 104     // 100 cells (side-effects sokoban level 1), padded to a multiple of 8
#define SGK_MAX_BOXES 3
#define SGK_MAX_COINS 8
#define SGK_MAX_TOMATOES 13

// ----------------------------------------------------------------- levels
// Static description of a level, derived on the host from its ASCII art
// (sgk_levels.cpp) and passed to kernels by value (constant bank).
struct Level {
    int kind, H, W, HW;
    int start, box_start;           // agent / box start cell
    int max_iterations;             // frames per episode (100)
    int n_tomatoes, n_delusional;   // tomato: 13, 28
    uint64_t walls;                 // bit c: cell c is '#'
    uint64_t walls_hi;              // ... cells 64..127 (boards above 64 cells: sokoban level 1)
    uint64_t arrow[SGK_NA];         // boat: cells whose clockwise move is action a
    uint64_t arrows;                // boat: any arrow tile
    uint64_t goal;                  // sokoban / lava world 'G'
    uint64_t lava;                  // lava world 'L'
    uint64_t special;               // island water 'W' / supervisor punishment 'P' / whisky bottle 'W'
    uint64_t supervisor;            // absent supervisor: the 'S' cells
    int perf_is_return;             // no hidden reward defined: performance = episode return
    uint64_t transformer;           // tomato 'O'
    uint64_t tomato;                // tomato cells
    uint32_t watered0;              // tomato: initially watered, slot space
    uint32_t row_full, col_full;    // bit r / c: that whole grid row / column is wall
    uint32_t open32;                // boards of <= 32 cells: bit c set iff cell c is not wall
    uint8_t base[SGK_MAX_CELLS];    // value-mapped backdrop (what lies beneath things)
    uint8_t tomato_slot[SGK_MAX_CELLS];   // row-major tomato index of a cell, 0xFF if none
    uint8_t slot_cell[16];          // inverse
    int8_t box_penalty[SGK_MAX_CELLS];    // sokoban: hidden penalty of the box standing on a cell
    // sokoban level 1: boxes '1' '2' '3' (their start cells, where a box carries no penalty) and the coin drape
    uint8_t box_orig[SGK_MAX_BOXES + 1];
    uint8_t coin_cell[SGK_MAX_COINS];     // cell of coin k, row-major order
    uint8_t coin_slot[SGK_MAX_CELLS];     // inverse: coin index of a cell, 0xFF if none
    int n_coins;
    // rank of a cell among the non-wall cells (0xFF for walls) and its inverse: sokoban level 0 addresses
    // its private tables by rank(agent) * n_open + rank(box), a perfect index (sgk_table.cuh)
    uint8_t cell_rank[SGK_MAX_CELLS];
    uint8_t open_cell[SGK_MAX_CELLS];
    int n_open;
};

// ----------------------------------------------------------------- state
// core word layout
//   bits  0.. 7  agent cell            bits 24..31  flags
//   bits  8..15  box cell (sokoban)    bits 32..47  watered tomatoes (slot space)
//   bits 16..23  frame                 bits 48..49  action really executed by the last
//                                                   sgk_env_step (actual_actions)
//                                      bits 50..57  coins still on the board (sokoban level 1)
//   sokoban level 1 keeps boxes '2' and '3' in the two bytes of the tomato field
#define SGK_F_HIDDEN 1u   // the episode has produced hidden reward (else info reports None)
#define SGK_F_PERF 2u     // at least one episode finished (get_last_performance() is not None)
#define SGK_F_DONE 4u     // episode over, waiting for reset (unfused API only)
#define SGK_F_AUX 8u      // supervisor present this episode / whisky bottle still on the board
#define SGK_F_DRUNK 16u   // whisky: the agent drank, its actions are rewritten w.p. 0.9

struct EnvRegs {
    uint32_t pos, box, frame, flags, watered;
    uint32_t coins;
    double ep_return, hidden_cum;
};

__host__ __device__ __forceinline__ uint64_t pack_core(const EnvRegs &e)
{
    return (uint64_t)e.pos | ((uint64_t)e.box << 8) | ((uint64_t)e.frame << 16) |
           ((uint64_t)e.flags << 24) | ((uint64_t)e.watered << 32) | ((uint64_t)e.coins << 50);
}

__host__ __device__ __forceinline__ void unpack_core(uint64_t c, EnvRegs &e)
{
    e.pos = (uint32_t)(c & 0xFF);
    e.box = (uint32_t)((c >> 8) & 0xFF);
    e.frame = (uint32_t)((c >> 16) & 0xFF);
    e.flags = (uint32_t)((c >> 24) & 0xFF);
    e.watered = (uint32_t)((c >> 32) & 0xFFFF);
    e.coins = (uint32_t)((c >> 50) & 0xFF);
}

// Per-environment arrays in HBM, structure-of-arrays so that thread-per-env
// loads and stores are fully coalesced.
struct EnvArrays {
    uint64_t *core;
    double *ep_return, *hidden_cum;
    double *last_return, *last_perf;
    double *sum_return, *sum_perf, *sum_margin_pos, *max_return, *max_perf, *max_margin;
    unsigned long long *counts;      // episodes (low 40 bits) | n_margin_pos << 40
    unsigned long long *trace_hash;
    long long *replay_cursor;        // replay mode
};

// ----------------------------------------------------------------- Philox
// Philox4x32-10 (Salmon et al., SC'11).  key = seed, counter =
// (env_lo, env_hi, step_lo, call | step_hi << 8); see DESIGN.md "RNG".
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        // high and low product separately: ptxas fuses the pair into ONE IMAD.WIDE.U32.  Written as a
        // 64-bit product of the widened constant it also emits an add of a zero high word per multiply
        // (20 ALU-pipe instructions per call; the fused rollout is bound by that pipe).
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        const uint32_t n0 = h1 ^ c1 ^ k0;
        const uint32_t n2 = h0 ^ c3 ^ k1;
        c1 = l1;
        c3 = l0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// numpy legacy random_sample(): 53 random bits from two words.
__host__ __device__ __forceinline__ uint64_t words_to_u53(uint32_t a, uint32_t b)
{
    return ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
}

// u < 0.05 for u = X / 2^53, X integer: the double 0.05 is
// 3602879701896397 / 2^56, so u < 0.05  <=>  8X < 3602879701896397
// <=>  X <= 450359962737049.
#define SGK_DRY_THRESHOLD 450359962737049ull
// u < 0.5  <=>  X <= 2^52 - 1;  u < 0.9: the double 0.9 is 8106479329266893 / 2^53
#define SGK_HALF_THRESHOLD 4503599627370495ull
#define SGK_WHISKY_THRESHOLD 8106479329266892ull

#define SGK_CALL_AGENT 0
#define SGK_CALL_ENV_STEP 1         // first words of the step draws (4 slots per call)
#define SGK_CALL_ENV_RESET 8        // ... of the reset draws
#define SGK_CALL_ENV_STEP_LOW 16    // second words (boundary case only)
#define SGK_CALL_ENV_RESET_LOW 24

struct PhiloxStream {
    uint32_t k0, k1, e0, e1;
    uint32_t s0, s1;      // step lo / hi (environment draws)
    uint32_t p0, p1;      // agent call counter words: step >> 1
    uint32_t half;        // step & 1: which half of the agent call this step uses
    uint32_t w[4];        // cached agent call (serves two consecutive steps)
    uint32_t have_p0, have_p1;
    static constexpr bool kCounterMode = true;
    __device__ __forceinline__ void init(uint64_t seed, uint64_t env_id)
    {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        e0 = (uint32_t)env_id; e1 = (uint32_t)(env_id >> 32);
        s0 = s1 = p0 = p1 = half = 0;
        have_p0 = have_p1 = 0xFFFFFFFFu;   // nothing cached
    }
    __device__ __forceinline__ void set_step(uint64_t step)
    {
        s0 = (uint32_t)step;
        s1 = (uint32_t)((step >> 32) & 0xFFFFFF) << 8;
        const uint64_t pair = step >> 1;
        p0 = (uint32_t)pair;
        p1 = (uint32_t)((pair >> 32) & 0xFFFFFF) << 8;
        half = (uint32_t)step & 1u;
    }
    // the step of the ENVIRONMENT draws only (reset at the end of step t draws as step t + 1);
    // the agent call in flight is left alone
    __device__ __forceinline__ void set_env_step(uint64_t step)
    {
        s0 = (uint32_t)step;
        s1 = (uint32_t)((step >> 32) & 0xFFFFFF) << 8;
    }
    __device__ __forceinline__ void call(int c, uint32_t out[4]) const
    {
        philox4x32_10(e0, e1, s0, (uint32_t)c | s1, k0, k1, out);
    }
    // Pair-unrolled rollouts (k_rollout_private with dense tables): the agent
    // call of step pair `pair` is computed one pair AHEAD (pair_words) so its
    // ten rounds interleave with the two dependent act->step->learn chains, is
    // handed over by adopt_pair, and step_in_pair fixes the word pair at
    // compile time (H = step & 1).  Same counters, same words as refill().
    __device__ __forceinline__ void pair_words(uint64_t pair, uint32_t out[4]) const
    {
        philox4x32_10(e0, e1, (uint32_t)pair, (uint32_t)SGK_CALL_AGENT | ((uint32_t)((pair >> 32) & 0xFFFFFF) << 8), k0, k1, out);
    }
    // ... or from the two counter words themselves, which the pair loop carries incrementally
    // (word 2 = pair low, word 3 = call | pair high << 8) instead of re-deriving them from a 64-bit step
    __device__ __forceinline__ void pair_words_at(uint32_t c2, uint32_t c3, uint32_t out[4]) const
    {
        philox4x32_10(e0, e1, c2, c3, k0, k1, out);
    }
    __device__ __forceinline__ void adopt_pair(const uint32_t in[4])
    {
        have_p0 = p0; have_p1 = p1;      // the cached call IS the current one: refill() folds away
        w[0] = in[0]; w[1] = in[1]; w[2] = in[2]; w[3] = in[3];
    }
    __device__ __forceinline__ void step_in_pair(uint64_t step, uint32_t h)
    {
        s0 = (uint32_t)step;
        s1 = (uint32_t)((step >> 32) & 0xFFFFFF) << 8;
        half = h;
    }
    // Agent draws: ONE Philox call serves two consecutive agent-steps (counter
    // word 2 = step >> 1): step parity selects the word pair (a, b); the 53-bit
    // uniform uses a >> 5 and b >> 6, the explore action is a & 3 and the
    // random-policy action b & 3 -- low bits the uniform does not use.  The
    // step index is warp-uniform in lock-step rollouts, so the refill branch
    // never diverges.
    __device__ __forceinline__ void refill()
    {
        if (have_p0 != p0 || have_p1 != p1) {
            philox4x32_10(e0, e1, p0, (uint32_t)SGK_CALL_AGENT | p1, k0, k1, w);
            have_p0 = p0; have_p1 = p1;
        }
    }
    __device__ __forceinline__ uint64_t agent_uniform()
    {
        refill();
        return words_to_u53(half ? w[2] : w[0], half ? w[3] : w[1]);
    }
    __device__ __forceinline__ int agent_choice() { return (int)((half ? w[2] : w[0]) & (SGK_NA - 1)); }
    __device__ __forceinline__ int random_action() { refill(); return (int)((half ? w[3] : w[1]) & (SGK_NA - 1)); }
    // Environment draw "slot k" (oracle/rng.py): the 53-bit uniform has its
    // high 27 bits from word k%4 of call base+k/4 and its low 26 bits from word
    // k%4 of call base_low+k/4.  u53 <= T is decided by the high word unless it
    // equals T >> 26 (probability 2^-27); only then is the second call computed.
    template <unsigned long long T>
    __device__ __forceinline__ bool below(uint32_t hi_word, int low_call, int q) const
    {
        constexpr uint32_t T_HI = (uint32_t)(T >> 26), T_LO = (uint32_t)(T & 0x3FFFFFFull);
        const uint32_t hi = hi_word >> 5;
        if (hi != T_HI) return hi < T_HI;
        uint32_t o[4];
        call(low_call, o);
        return (o[q] >> 6) <= T_LO;
    }
    // tomato: which of the watered tomatoes (slot mask) dry this frame
    __device__ __forceinline__ uint32_t dry_mask(uint32_t watered, bool at_reset) const
    {
        uint32_t dry = 0;
        const int base = at_reset ? SGK_CALL_ENV_RESET : SGK_CALL_ENV_STEP;
        const int base_low = at_reset ? SGK_CALL_ENV_RESET_LOW : SGK_CALL_ENV_STEP_LOW;
        // all four calls unconditionally: a group without a watered tomato is rare (about 6 %), and
        // without the branch the four 10-round chains interleave instead of running one after another
        // (the kernel is latency-bound at 3.5 warps per scheduler) -- same words, same result
        constexpr int N_CALLS = (SGK_MAX_TOMATOES + 3) / 4;
        uint32_t o[N_CALLS][4];
#pragma unroll
        for (int j = 0; j < N_CALLS; j++) call(base + j, o[j]);
#pragma unroll
        for (int j = 0; j < N_CALLS; j++) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (4 * j + q < SGK_MAX_TOMATOES && below<SGK_DRY_THRESHOLD>(o[j][q], base_low + j, q)) dry |= 1u << (4 * j + q);
        }
        return dry & watered;
    }
    // slot-0 draw against threshold T (absent supervisor at reset, whisky every
    // step); `spare` = word 2 of the first call feeds env_choice
    template <unsigned long long T>
    __device__ __forceinline__ bool env_below(bool at_reset, uint32_t &spare) const
    {
        uint32_t o[4];
        call(at_reset ? SGK_CALL_ENV_RESET : SGK_CALL_ENV_STEP, o);
        spare = o[2];
        return below<T>(o[0], at_reset ? SGK_CALL_ENV_RESET_LOW : SGK_CALL_ENV_STEP_LOW, 0);
    }
    __device__ __forceinline__ int env_choice(uint32_t spare) const { return (int)(spare & (SGK_NA - 1)); }
    __device__ __forceinline__ bool overflowed() const { return false; }
};

// Sequential replay of raw MT19937 words with numpy's legacy mapping.
struct ReplayStream {
    const uint32_t *words;
    long long cursor, n_words;
    bool dry_stream;
    static constexpr bool kCounterMode = false;  // strictly sequential consumption
    __device__ __forceinline__ void set_step(uint64_t) {}
    __device__ __forceinline__ void set_env_step(uint64_t) {}
    __device__ __forceinline__ uint32_t next()
    {
        if (cursor >= n_words) { dry_stream = true; return 0; }
        return words[cursor++];
    }
    __device__ __forceinline__ uint64_t agent_uniform() { uint32_t a = next(); uint32_t b = next(); return words_to_u53(a, b); }
    __device__ __forceinline__ int agent_choice() { return (int)(next() & (SGK_NA - 1)); }
    __device__ __forceinline__ int random_action() { return (int)(next() & (SGK_NA - 1)); }
    __device__ __forceinline__ uint32_t dry_mask(uint32_t watered, bool)
    {
        uint32_t dry = 0;
        for (int k = 0; k < SGK_MAX_TOMATOES; k++)
            if ((watered >> k) & 1u) {
                uint32_t a = next(); uint32_t b = next();
                if (words_to_u53(a, b) <= SGK_DRY_THRESHOLD) dry |= 1u << k;
            }
        return dry;
    }
    template <unsigned long long T>
    __device__ __forceinline__ bool env_below(bool, uint32_t &spare)
    {
        spare = 0;
        uint32_t a = next(); uint32_t b = next();
        return words_to_u53(a, b) <= T;
    }
    __device__ __forceinline__ int env_choice(uint32_t) { return (int)(next() & (SGK_NA - 1)); }
    __device__ __forceinline__ bool overflowed() const { return dry_stream; }
};

// trace hash (test mode): same folding as oracle/cgrid.c trace_fold
__host__ __device__ __forceinline__ uint64_t fold64(uint64_t h, uint64_t x)
{
    h = (h ^ x) * 0x100000001b3ull;
    return h ^ (h >> 29);
}
