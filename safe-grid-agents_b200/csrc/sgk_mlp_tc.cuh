// sgk_mlp_tc.cuh -- fused MLP forward on the 5th-generation tensor cores.
//
// The deep-Q network of the reference (value.py:148-158, n_layers = 2):
//     Q(s) = W3 relu(W2 relu(W1 s + b1) + b2) + b3,   36 -> 100 -> 100 -> 4
// is the only dense contraction on the rollout path.  One CTA (4 warps) owns a
// tile of 128 boards and runs all three layers without leaving the SM:
//
//   * the three weight matrices are staged ONCE per CTA into shared memory as
//     TF32 operands in the canonical K-major, no-swizzle UMMA layout
//     (16-byte chunks along K; chunk-major so that LBO = rows * 16 B and
//     SBO = 128 B) and stay resident while the CTA walks its tiles;
//   * every layer is a short chain of tcgen05.mma (cta_group::1, kind::tf32,
//     M = 128, N = 112 / 112 / 16, K = 8 per instruction) issued by one elected
//     thread, accumulating in TMEM (fp32);
//   * completion is signalled with tcgen05.commit on an mbarrier; each thread
//     then reads ITS row of the accumulator with tcgen05.ld (32x32b: thread r
//     of the CTA <-> TMEM lane r <-> board r of the tile), applies bias + ReLU
//     and writes the activations straight back into shared memory as the next
//     layer's A operand -- activations never touch HBM unless the caller asks
//     for them (the fp32 backward pass wants H1 and H2).
//
// TF32 keeps 10 mantissa bits: boards (small integers) are exact, weights and
// hidden activations are rounded to nearest (cvt.rna.tf32), accumulation is
// fp32.  Results agree with the fp32 path to ~1e-3 relative; tests state the
// tolerance.  No cuBLAS/CUTLASS: descriptors and PTX are written out below.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

constexpr int TILE_M = 128;        // boards per tile = TMEM lanes = threads per CTA
constexpr int N_HID = 112;         // hidden width padded to a multiple of 16 (>= 100)
constexpr int N_OUT = 16;          // action count padded to the minimum UMMA N for M = 128
constexpr int K_HID = 104;         // hidden width padded to a multiple of 8 (UMMA K for tf32)
constexpr int MAX_K_IN = 64;       // boards up to 64 cells
constexpr int TMEM_COLS = 256;     // D1/D3 at column 0, D2 at column 128

struct Params {
    const float *w1, *b1, *w2, *b2, *w3, *b3;   // torch layout: W[out][in] row-major
    int n_in, n_hidden, n_out;                  // 36/25/63, <= 100, 4
    const uint8_t *boards;                      // [rows][n_in]
    int64_t rows;
    float *q_out;                               // [rows][n_out]
    float *h1_out, *h2_out;                     // [rows][n_hidden] or null
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor):
// bits [0,14) start >> 4, [16,30) leading byte offset >> 4 (between the two
// 16-byte K chunks of one instruction), [32,46) stride byte offset >> 4
// (between 8-row groups), [46,48) version = 1, [61,64) layout type = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}

// UMMA instruction descriptor (cute::UMMA::InstrDescriptor), kind::tf32:
// c_format F32 (1) at [4,6), a/b format TF32 (2) at [7,10)/[10,13), both
// K-major (bits 15,16 = 0), N >> 3 at [17,23), M >> 4 at [24,29).
__device__ __forceinline__ constexpr uint32_t make_idesc(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t *mbar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *mbar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}\n"
        :: "r"(smem_u32(mbar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}

// shared-memory plan (bytes); every operand block is a multiple of 16 B
struct Smem {
    static constexpr int CHUNK_A = TILE_M * 16;            // one 16-byte K chunk of all 128 rows
    static constexpr int CHUNK_H = N_HID * 16;             // ... of the 112 weight rows
    static constexpr int CHUNK_O = N_OUT * 16;
    static constexpr int W1 = 0;                            // [MAX_K_IN/4][112][16 B]
    static constexpr int W2 = W1 + (MAX_K_IN / 4) * CHUNK_H;
    static constexpr int W3 = W2 + (K_HID / 4) * CHUNK_H;   // [26][16][16 B]
    static constexpr int X = W3 + (K_HID / 4) * CHUNK_O;    // [MAX_K_IN/4][128][16 B]
    static constexpr int H = X + (MAX_K_IN / 4) * CHUNK_A;  // [26][128][16 B]
    static constexpr int BIAS = H + (K_HID / 4) * CHUNK_A;  // b1[112] b2[112] b3[16]
    static constexpr int BAR = BIAS + (2 * N_HID + N_OUT) * 4;
    static constexpr int TOTAL = BAR + 16;
};

// W[out][in] (global, row-major) -> chunk-major TF32 operand with `rows_pad`
// rows and `k_pad` columns, zero padded.  Consecutive threads take consecutive
// 16-byte pieces of a weight row (coalesced global reads, float4 when the row
// length allows), four pieces in flight per thread.
__device__ __forceinline__ void stage_weights(uint8_t *dst, const float *w, int n_out, int n_in, int rows_pad, int k_pad)
{
    const int chunks = k_pad / 4;
    const bool vec = (n_in & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0;
#pragma unroll 4
    for (int e = threadIdx.x; e < chunks * rows_pad; e += blockDim.x) {
        const int r = e / chunks, c = e - r * chunks;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = 4 * c;
        if (r < n_out && k < n_in) {
            const float *src = w + (size_t)r * n_in + k;
            if (vec) {
                v = __ldg(reinterpret_cast<const float4 *>(src));
            } else {
                v.x = __ldg(src);
                v.y = k + 1 < n_in ? __ldg(src + 1) : 0.f;
                v.z = k + 2 < n_in ? __ldg(src + 2) : 0.f;
                v.w = k + 3 < n_in ? __ldg(src + 3) : 0.f;
            }
            v = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        }
        *reinterpret_cast<float4 *>(dst + (size_t)c * rows_pad * 16 + r * 16) = v;
    }
}

// epilogue of a hidden layer: this thread's accumulator row -> bias, ReLU ->
// next layer's A operand in shared memory (+ optional fp32 copy in HBM)
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_row, const float *bias, uint8_t *h_smem, int row_in_tile,
                                                float *h_out_row, int n_hidden)
{
#pragma unroll 1
    for (int c8 = 0; c8 < K_HID / 8; c8++) {
        float v[8];
        tmem_ld8(tmem_row + c8 * 8, v);
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = fmaxf(v[i] + bias[c8 * 8 + i], 0.f);
        if (h_out_row) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (c8 * 8 + i < n_hidden) h_out_row[c8 * 8 + i] = v[i];
        }
        float4 lo = make_float4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
        float4 hi = make_float4(to_tf32(v[4]), to_tf32(v[5]), to_tf32(v[6]), to_tf32(v[7]));
        *reinterpret_cast<float4 *>(h_smem + (size_t)(2 * c8) * Smem::CHUNK_A + row_in_tile * 16) = lo;
        *reinterpret_cast<float4 *>(h_smem + (size_t)(2 * c8 + 1) * Smem::CHUNK_A + row_in_tile * 16) = hi;
    }
}

__global__ void __launch_bounds__(TILE_M, 1) k_mlp_forward_tc(const Params p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_slot;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + Smem::BAR);
    float *bias = reinterpret_cast<float *>(smem + Smem::BIAS);
    const int warp = threadIdx.x >> 5;
    const int k_in = (p.n_in + 7) & ~7;             // 36 -> 40, 25 -> 32, 63 -> 64

    // ---- one-time setup: weights, biases, barrier, TMEM
    stage_weights(smem + Smem::W1, p.w1, p.n_hidden, p.n_in, N_HID, k_in);
    stage_weights(smem + Smem::W2, p.w2, p.n_hidden, p.n_hidden, N_HID, K_HID);
    stage_weights(smem + Smem::W3, p.w3, p.n_out, p.n_hidden, N_OUT, K_HID);
    for (int i = threadIdx.x; i < 2 * N_HID + N_OUT; i += blockDim.x) {
        float b = 0.f;
        if (i < N_HID) b = i < p.n_hidden ? p.b1[i] : 0.f;
        else if (i < 2 * N_HID) b = i - N_HID < p.n_hidden ? p.b2[i - N_HID] : 0.f;
        else b = i - 2 * N_HID < p.n_out ? p.b3[i - 2 * N_HID] : 0.f;
        bias[i] = b;
    }
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes
    const uint32_t d1 = tmem, d2 = tmem + 128, d3 = tmem;
    const uint32_t a_x = smem_u32(smem + Smem::X), a_h = smem_u32(smem + Smem::H);
    const uint32_t b_w1 = smem_u32(smem + Smem::W1), b_w2 = smem_u32(smem + Smem::W2), b_w3 = smem_u32(smem + Smem::W3);
    constexpr uint32_t IDESC_HID = make_idesc(TILE_M, N_HID), IDESC_OUT = make_idesc(TILE_M, N_OUT);
    uint32_t phase = 0;

    const int64_t n_tiles = (p.rows + TILE_M - 1) / TILE_M;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * TILE_M + threadIdx.x;
        const bool valid = row < p.rows;
        // ---- stage this thread's board as A operand of layer 1
        for (int c = 0; c < k_in / 4; c++) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const uint8_t *b = p.boards + row * p.n_in + 4 * c;
                const int k = 4 * c;
                v.x = k + 0 < p.n_in ? (float)b[0] : 0.f;
                v.y = k + 1 < p.n_in ? (float)b[1] : 0.f;
                v.z = k + 2 < p.n_in ? (float)b[2] : 0.f;
                v.w = k + 3 < p.n_in ? (float)b[3] : 0.f;
            }
            *reinterpret_cast<float4 *>(smem + Smem::X + (size_t)c * Smem::CHUNK_A + threadIdx.x * 16) = v;
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        // ---- layer 1: D1[128 x 112] = X[128 x k_in] * W1^T
        if (threadIdx.x == 0) {
            for (int k = 0; k < k_in / 8; k++)
                mma_tf32(d1, make_desc(a_x + k * 2 * Smem::CHUNK_A, Smem::CHUNK_A, 128),
                         make_desc(b_w1 + k * 2 * Smem::CHUNK_H, Smem::CHUNK_H, 128), IDESC_HID, k > 0);
            mma_commit(mbar);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        hidden_epilogue(tmem_row + 0, bias, smem + Smem::H, threadIdx.x,
                        (valid && p.h1_out) ? p.h1_out + row * p.n_hidden : nullptr, p.n_hidden);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        // ---- layer 2: D2[128 x 112] = H1[128 x 104] * W2^T
        if (threadIdx.x == 0) {
            for (int k = 0; k < K_HID / 8; k++)
                mma_tf32(d2, make_desc(a_h + k * 2 * Smem::CHUNK_A, Smem::CHUNK_A, 128),
                         make_desc(b_w2 + k * 2 * Smem::CHUNK_H, Smem::CHUNK_H, 128), IDESC_HID, k > 0);
            mma_commit(mbar);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        hidden_epilogue(tmem_row + 128, bias + N_HID, smem + Smem::H, threadIdx.x,
                        (valid && p.h2_out) ? p.h2_out + row * p.n_hidden : nullptr, p.n_hidden);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        // ---- layer 3: D3[128 x 16] = H2[128 x 104] * W3^T
        if (threadIdx.x == 0) {
            for (int k = 0; k < K_HID / 8; k++)
                mma_tf32(d3, make_desc(a_h + k * 2 * Smem::CHUNK_A, Smem::CHUNK_A, 128),
                         make_desc(b_w3 + k * 2 * Smem::CHUNK_O, Smem::CHUNK_O, 128), IDESC_OUT, k > 0);
            mma_commit(mbar);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        {
            float v[8];
            tmem_ld8(tmem_row + 0, v);
            if (valid) {
                const float *b3 = bias + 2 * N_HID;
                if (p.n_out == 4) {
                    *reinterpret_cast<float4 *>(p.q_out + row * 4) = make_float4(v[0] + b3[0], v[1] + b3[1], v[2] + b3[2], v[3] + b3[3]);
                } else {
                    for (int i = 0; i < p.n_out && i < 8; i++) p.q_out[row * p.n_out + i] = v[i] + b3[i];
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS) : "memory");
}

}  // namespace tc
