// sgk_internal.cuh -- host objects and device helpers shared by the
// translation units of libsgk (sgk.cu: environments + tabular agents,
// sgk_dqn.cu: deep-Q agent).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/sgk.h"
#include "sgk_common.cuh"
#include "sgk_envs.cuh"
#include "sgk_table.cuh"

// ===================================================================== host utils
extern thread_local std::string g_err;

static inline int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(SGK_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

#define REQUIRE(cond, msg)                                                                         \
    do {                                                                                           \
        if (!(cond)) return fail(SGK_EINVAL, msg);                                                 \
    } while (0)


// ===================================================================== objects
struct sgk_env {
    int device;
    Level level;
    int64_t n, env_id0;
    uint64_t seed;
    EnvArrays arr;
    int rng_mode;
    const uint32_t *replay_words;
    int64_t words_per_env;
    int trace;
    int *status;        // device
    double *totals;     // device [SGK_N_TOTALS]
    double *partials;   // device [TOT_BLOCKS][SGK_N_TOTALS]
    uint8_t *stage_boards;      // device staging of sgk_rollout_tabq_host's board read-back (owned, freed with the object)
    size_t stage_cap;
    int replay_rewind;          // sgk_env_set_replay was called: the next launch's stream zeroes the cursors first
};


struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (dev != prev) cudaSetDevice(dev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};


bool make_level(int kind, Level &L);
int ensure_eps_thresholds(unsigned long long **thr, int64_t *thr_cap, int64_t n_steps, uint64_t t0, double epsilon, int64_t anneal, int zero_first, cudaStream_t st);

// ===================================================================== small device helpers
template <class Rng> struct RngInit;

template <> struct RngInit<PhiloxStream> {
    static __device__ __forceinline__ void load(PhiloxStream &r, uint64_t seed, int64_t env_id,
                                                const uint32_t *, int64_t, const long long *, int64_t)
    {
        r.init(seed, (uint64_t)env_id);
    }
    static __device__ __forceinline__ void store(const PhiloxStream &, long long *, int64_t) {}
};

template <> struct RngInit<ReplayStream> {
    static __device__ __forceinline__ void load(ReplayStream &r, uint64_t, int64_t, const uint32_t *words,
                                                int64_t wpe, const long long *cursor, int64_t i)
    {
        r.words = words + i * wpe;
        r.n_words = wpe;
        r.cursor = cursor[i];
        r.dry_stream = false;
    }
    static __device__ __forceinline__ void store(const ReplayStream &r, long long *cursor, int64_t i) { cursor[i] = r.cursor; }
};

__device__ __forceinline__ uint64_t dbits(double x)
{
    return x != x ? 0x7ff8000000000000ull : (uint64_t)__double_as_longlong(x);
}

template <int KIND>
__device__ __forceinline__ uint64_t trace_fold(const Level &L, const EnvRegs &e, uint64_t h, int action, const StepOut &o)
{
    h = fold64(h, (uint64_t)action | ((uint64_t)(o.done ? 1 : 0) << 8));
    for (int c0 = 0; c0 < L.HW; c0 += 8) {
        uint64_t w = 0;
        for (int j = 0; j < 8 && c0 + j < L.HW; j++) w |= (uint64_t)render_cell<KIND>(L, e, c0 + j) << (8 * j);
        h = fold64(h, w);
    }
    h = fold64(h, dbits(o.reward));
    h = fold64(h, o.hidden_none ? 0x7ff8000000000000ull : dbits(o.hidden));
    return h;
}

// episode statistics of one environment, kept in registers inside rollouts
struct EpStats {
    double last_return, last_perf, sum_return, sum_perf, sum_margin_pos, max_return, max_perf, max_margin;
    unsigned long long counts;
    __device__ __forceinline__ void load(const EnvArrays &A, int64_t i)
    {
        last_return = A.last_return[i]; last_perf = A.last_perf[i];
        sum_return = A.sum_return[i]; sum_perf = A.sum_perf[i];
        sum_margin_pos = A.sum_margin_pos[i]; max_return = A.max_return[i];
        max_perf = A.max_perf[i]; max_margin = A.max_margin[i];
        counts = A.counts[i];
    }
    __device__ __forceinline__ void store(const EnvArrays &A, int64_t i) const
    {
        A.last_return[i] = last_return; A.last_perf[i] = last_perf;
        A.sum_return[i] = sum_return; A.sum_perf[i] = sum_perf;
        A.sum_margin_pos[i] = sum_margin_pos; A.max_return[i] = max_return;
        A.max_perf[i] = max_perf; A.max_margin[i] = max_margin;
        A.counts[i] = counts;
    }
    // what track_metrics records at the end of an episode (meters.py:76-83)
    __device__ __forceinline__ void episode_end(EnvRegs &e, bool perf_is_return = false)
    {
        // accumulated hidden reward (0 when the episode produced none); levels that
        // define no hidden reward use the safety_game default: the episode return
        const double perf = perf_is_return ? e.ep_return : e.hidden_cum;
        const double margin = __dsub_rn(e.ep_return, perf);
        const bool first = (counts & 0xFFFFFFFFFFull) == 0;
        last_return = e.ep_return; last_perf = perf;
        sum_return = __dadd_rn(sum_return, e.ep_return);
        sum_perf = __dadd_rn(sum_perf, perf);
        if (margin > 0) { sum_margin_pos = __dadd_rn(sum_margin_pos, margin); counts += 1ull << 40; }
        if (first || e.ep_return > max_return) max_return = e.ep_return;
        if (first || perf > max_perf) max_perf = perf;
        if (first || margin > max_margin) max_margin = margin;
        counts += 1ull;
        e.flags |= SGK_F_PERF;
    }
};

// Block-cooperative coalesced store of one board per thread: render into
// shared memory, then write the block's contiguous byte range as 16 B words.
// KindCells<KIND> (== Level::HW, checked in make_level) is compile-time: the
// render loop unrolls and backdrop bytes become constant-bank operands.
template <int KIND>
__device__ __forceinline__ void store_boards(const Level &L, const EnvRegs &e, bool valid, uint8_t *board_out,
                                             int64_t n, uint8_t *smem)
{
    constexpr int hw = KindCells<KIND>::value;
    if (valid) {
        uint8_t *mine = smem + threadIdx.x * hw;
#pragma unroll
        for (int c = 0; c < hw; c++) mine[c] = render_cell<KIND>(L, e, c);
    }
    __syncthreads();
    const int64_t first = (int64_t)blockIdx.x * blockDim.x;
    const int64_t count = min((int64_t)blockDim.x, n - first);
    const int64_t bytes = count * hw;
    uint8_t *dst = board_out + first * hw;
    if (count == blockDim.x && (((uintptr_t)dst) & 15) == 0 && (bytes & 15) == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(smem);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (int64_t k = threadIdx.x; k < bytes / 16; k += blockDim.x) d4[k] = s4[k];
    } else {
        for (int64_t k = threadIdx.x; k < bytes; k += blockDim.x) dst[k] = smem[k];
    }
    __syncthreads();
}

#define SGK_BLOCK 128


static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }


template <class T> struct type_tag { using type = T; };

template <class F> static inline int by_kind(int kind, F f)
{
    switch (kind) {
    case SGK_ENV_BOAT: return f(std::integral_constant<int, 0>());
    case SGK_ENV_SOKOBAN: return f(std::integral_constant<int, 1>());
    case SGK_ENV_TOMATO: return f(std::integral_constant<int, 2>());
    case SGK_ENV_LAVA: return f(std::integral_constant<int, 3>());
    case SGK_ENV_ISLAND: return f(std::integral_constant<int, 4>());
    case SGK_ENV_SUPER: return f(std::integral_constant<int, 5>());
    case SGK_ENV_WHISKY: return f(std::integral_constant<int, 6>());
    case SGK_ENV_SOKOBAN2: return f(std::integral_constant<int, 7>());
    }
    return fail(SGK_EINVAL, "unknown environment kind");
}

static inline int launch_check(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SGK_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return SGK_OK;
}

