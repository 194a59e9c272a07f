// sgk_mlp_tc.cuh -- the deep-Q network on the 5th-generation tensor cores: fused forward
// (k_mlp_forward_ts, at the end of this file), fused backward (k_mlp_backward_fused).
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
    const uint8_t *w_image;                     // W1|W2|W3 already in the shared-memory operand layout (k_pack_weights), or null
    // k_mlp_forward_ts: post-ReLU activations as FP16 operand images for the fused backward,
    // [tile][14 chunks][128 rows][16 B] with a 1.0 at feature n_hidden (bias-gradient column), or null
    uint8_t *h1_img, *h2_img;
};
constexpr int H_IMG_CHUNKS = N_HID / 8;                         // 14 chunks of 8 halfs
constexpr int H_IMG_TILE_BYTES = H_IMG_CHUNKS * TILE_M * 16;    // 28,672 bytes per 128-row tile

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// The same rounding (nearest, ties away from zero) in two integer instructions, for finite inputs:
// cvt.rna.tf32 expands to ~5 SASS instructions with its NaN handling, and the 3xTF32 epilogue needs two
// conversions per accumulator element.
__device__ __forceinline__ float to_tf32_fast(float x)
{
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
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

// One TMA bulk copy global -> shared (cp.async.bulk, no tensor map: the image is
// contiguous), completion signalled on `mbar` as a byte count.  Issued by ONE
// thread; `bytes` % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *mbar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
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

constexpr uint32_t FWD_IMAGE_BYTES = Smem::X;    // W1 | W2 | W3 regions, contiguous from offset 0

// W[out][in] (global, row-major) -> chunk-major TF32 operand with `rows_pad`
// rows and `k_pad` columns, zero padded.  Consecutive threads take consecutive
// 16-byte pieces of a weight row (coalesced global reads, float4 when the row
// length allows), four pieces in flight per thread.
// hi / lo split of an fp32 value into two TF32 numbers: x = hi + lo up to 2^-22 relative
__device__ __forceinline__ float tf32_part(float x, int lo_part)
{
    const float hi = to_tf32(x);
    return lo_part ? to_tf32(x - hi) : hi;
}

__device__ __forceinline__ void stage_weights(uint8_t *dst, const float *w, int n_out, int n_in, int rows_pad, int k_pad,
                                              int tid = threadIdx.x, int nth = blockDim.x, int lo_part = 0)
{
    const int chunks = k_pad / 4;
    const bool vec = (n_in & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0;
#pragma unroll 4
    for (int e = tid; e < chunks * rows_pad; e += nth) {
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
            v = make_float4(tf32_part(v.x, lo_part), tf32_part(v.y, lo_part), tf32_part(v.z, lo_part), tf32_part(v.w, lo_part));
        }
        *reinterpret_cast<float4 *>(dst + (size_t)c * rows_pad * 16 + r * 16) = v;
    }
}

// 16-byte row accesses with a scalar tail (rows are n_hidden floats, 16-byte
// aligned when n_hidden % 4 == 0): 4x fewer memory transactions than scalars.
__device__ __forceinline__ float4 load4(const float *row, int col, int n, bool vec)
{
    if (vec && col + 3 < n) return *reinterpret_cast<const float4 *>(row + col);
    float4 v;
    v.x = col + 0 < n ? row[col + 0] : 0.f;
    v.y = col + 1 < n ? row[col + 1] : 0.f;
    v.z = col + 2 < n ? row[col + 2] : 0.f;
    v.w = col + 3 < n ? row[col + 3] : 0.f;
    return v;
}

__device__ __forceinline__ void store4(float *row, int col, int n, bool vec, float4 v)
{
    if (vec && col + 3 < n) { *reinterpret_cast<float4 *>(row + col) = v; return; }
    if (col + 0 < n) row[col + 0] = v.x;
    if (col + 1 < n) row[col + 1] = v.y;
    if (col + 2 < n) row[col + 2] = v.z;
    if (col + 3 < n) row[col + 3] = v.w;
}

// ===================================================================== backward
// Backward pass of the same network on the tensor cores, in two kernels.
//
// k_mlp_backward_data_tc: the error chain dQ -> dH2 -> dH1, the forward's
// structure run with transposed weights:
//     dH2 = (dQ  * W3) .* (H2 > 0)        M 128, N 112, K 8   (one MMA)
//     dH1 = (dH2 * W2) .* (H1 > 0)        M 128, N 112, K 104
// W3^T and W2^T are staged once per CTA as K-major B operands; the ReLU masks
// come from the forward's H1 / H2 in HBM; dH2 goes back into shared memory as
// the next A operand and both error signals are written out in fp32 for the
// weight-gradient kernel.
struct BwdParams {
    const float *w2, *w3;          // W2[h][h], W3[n_out][h], torch layout
    int n_hidden, n_out;
    const float *dq;               // [rows][n_out]
    const float *h1, *h2;          // [rows][n_hidden], post-ReLU
    float *dh1, *dh2;              // [rows][n_hidden]
    int64_t rows;
    const uint8_t *w_image;        // W3^T|W2^T already in the operand layout, or null
};

struct SmemBwd {
    static constexpr int W3T = 0;                                   // [2][112][16 B]   (K = 8 actions padded)
    static constexpr int W2T = W3T + 2 * Smem::CHUNK_H;             // [26][112][16 B]
    static constexpr int DQ = W2T + (K_HID / 4) * Smem::CHUNK_H;    // [2][128][16 B]
    static constexpr int DH = DQ + 2 * Smem::CHUNK_A;               // [26][128][16 B]
    static constexpr int BAR = DH + (K_HID / 4) * Smem::CHUNK_A;
    static constexpr int TOTAL = BAR + 16;
};

constexpr uint32_t BWD_IMAGE_BYTES = SmemBwd::DQ;    // W3^T | W2^T regions, contiguous from offset 0

// B operand holding W^T: row r = input index (< n_in), K index = output index (< n_out)
__device__ __forceinline__ void stage_weights_T(uint8_t *dst, const float *w, int n_out, int n_in, int rows_pad, int k_pad,
                                                int tid = threadIdx.x, int nth = blockDim.x)
{
    const int chunks = k_pad / 4;
#pragma unroll 4
    for (int e = tid; e < chunks * rows_pad; e += nth) {
        const int c = e / rows_pad, r = e - c * rows_pad;      // consecutive threads: consecutive r (coalesced reads of W rows)
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n_in) {
            const int k = 4 * c;
            v.x = k + 0 < n_out ? to_tf32(__ldg(w + (size_t)(k + 0) * n_in + r)) : 0.f;
            v.y = k + 1 < n_out ? to_tf32(__ldg(w + (size_t)(k + 1) * n_in + r)) : 0.f;
            v.z = k + 2 < n_out ? to_tf32(__ldg(w + (size_t)(k + 2) * n_in + r)) : 0.f;
            v.w = k + 3 < n_out ? to_tf32(__ldg(w + (size_t)(k + 3) * n_in + r)) : 0.f;
        }
        *reinterpret_cast<float4 *>(dst + (size_t)c * rows_pad * 16 + r * 16) = v;
    }
}

__device__ __forceinline__ uint32_t tmem_alloc_and_sync(uint32_t *slot, int warp, uint32_t cols_256_or_128)
{
    if (warp == 0) {
        if (cols_256_or_128 == 256)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(slot)) : "memory");
        else
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *slot;
}

__global__ void __launch_bounds__(TILE_M, 1) k_mlp_backward_data_tc(const BwdParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_slot;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + SmemBwd::BAR);
    const int warp = threadIdx.x >> 5;
    uint64_t *mbar_w = mbar + 1;
    if (!p.w_image) {
        stage_weights_T(smem + SmemBwd::W3T, p.w3, p.n_out, p.n_hidden, N_HID, 8);
        stage_weights_T(smem + SmemBwd::W2T, p.w2, p.n_hidden, p.n_hidden, N_HID, K_HID);
    }
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        mbar_init(mbar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (p.w_image) bulk_load(smem + SmemBwd::W3T, p.w_image, BWD_IMAGE_BYTES, mbar_w);
    }
    const uint32_t tmem = tmem_alloc_and_sync(&tmem_base_slot, warp, 256);
    if (p.w_image) mbar_wait(mbar_w, 0);
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t d0 = tmem, d1 = tmem + 128;
    const uint32_t a_dq = smem_u32(smem + SmemBwd::DQ), a_dh = smem_u32(smem + SmemBwd::DH);
    const uint32_t b_w3t = smem_u32(smem + SmemBwd::W3T), b_w2t = smem_u32(smem + SmemBwd::W2T);
    constexpr uint32_t IDESC = make_idesc(TILE_M, N_HID);
    uint32_t phase = 0;
    const bool vec = (p.n_hidden & 3) == 0;
    const int64_t n_tiles = (p.rows + TILE_M - 1) / TILE_M;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * TILE_M + threadIdx.x;
        const bool valid = row < p.rows;
        {   // dQ row -> A operand (K = 8: 4 actions + zero padding)
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float *q = p.dq + row * p.n_out;
                v.x = q[0];
                v.y = p.n_out > 1 ? q[1] : 0.f;
                v.z = p.n_out > 2 ? q[2] : 0.f;
                v.w = p.n_out > 3 ? q[3] : 0.f;
                v = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
            }
            *reinterpret_cast<float4 *>(smem + SmemBwd::DQ + threadIdx.x * 16) = v;
            *reinterpret_cast<float4 *>(smem + SmemBwd::DQ + Smem::CHUNK_A + threadIdx.x * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (threadIdx.x == 0) {
            mma_tf32(d0, make_desc(a_dq, Smem::CHUNK_A, 128), make_desc(b_w3t, Smem::CHUNK_H, 128), IDESC, 0);
            mma_commit(mbar);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        // dH2 = D0 .* (H2 > 0) -> HBM and next A operand
#pragma unroll 1
        for (int c8 = 0; c8 < K_HID / 8; c8++) {
            float v[8];
            tmem_ld8(tmem_row + c8 * 8, v);
            if (valid) {
                const float4 m0 = load4(p.h2 + row * p.n_hidden, c8 * 8, p.n_hidden, vec);
                const float4 m1 = load4(p.h2 + row * p.n_hidden, c8 * 8 + 4, p.n_hidden, vec);
                v[0] = m0.x > 0.f ? v[0] : 0.f; v[1] = m0.y > 0.f ? v[1] : 0.f;
                v[2] = m0.z > 0.f ? v[2] : 0.f; v[3] = m0.w > 0.f ? v[3] : 0.f;
                v[4] = m1.x > 0.f ? v[4] : 0.f; v[5] = m1.y > 0.f ? v[5] : 0.f;
                v[6] = m1.z > 0.f ? v[6] : 0.f; v[7] = m1.w > 0.f ? v[7] : 0.f;
                store4(p.dh2 + row * p.n_hidden, c8 * 8, p.n_hidden, vec, make_float4(v[0], v[1], v[2], v[3]));
                store4(p.dh2 + row * p.n_hidden, c8 * 8 + 4, p.n_hidden, vec, make_float4(v[4], v[5], v[6], v[7]));
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = 0.f;
            }
            float4 lo = make_float4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
            float4 hi = make_float4(to_tf32(v[4]), to_tf32(v[5]), to_tf32(v[6]), to_tf32(v[7]));
            *reinterpret_cast<float4 *>(smem + SmemBwd::DH + (size_t)(2 * c8) * Smem::CHUNK_A + threadIdx.x * 16) = lo;
            *reinterpret_cast<float4 *>(smem + SmemBwd::DH + (size_t)(2 * c8 + 1) * Smem::CHUNK_A + threadIdx.x * 16) = hi;
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (threadIdx.x == 0) {
            for (int k = 0; k < K_HID / 8; k++)
                mma_tf32(d1, make_desc(a_dh + k * 2 * Smem::CHUNK_A, Smem::CHUNK_A, 128),
                         make_desc(b_w2t + k * 2 * Smem::CHUNK_H, Smem::CHUNK_H, 128), IDESC, k > 0);
            mma_commit(mbar);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c8 = 0; c8 < K_HID / 8; c8++) {
            float v[8];
            tmem_ld8(tmem_row + 128 + c8 * 8, v);
            if (valid) {
                const float4 m0 = load4(p.h1 + row * p.n_hidden, c8 * 8, p.n_hidden, vec);
                const float4 m1 = load4(p.h1 + row * p.n_hidden, c8 * 8 + 4, p.n_hidden, vec);
                store4(p.dh1 + row * p.n_hidden, c8 * 8, p.n_hidden, vec,
                       make_float4(m0.x > 0.f ? v[0] : 0.f, m0.y > 0.f ? v[1] : 0.f, m0.z > 0.f ? v[2] : 0.f, m0.w > 0.f ? v[3] : 0.f));
                store4(p.dh1 + row * p.n_hidden, c8 * 8 + 4, p.n_hidden, vec,
                       make_float4(m1.x > 0.f ? v[4] : 0.f, m1.y > 0.f ? v[5] : 0.f, m1.z > 0.f ? v[6] : 0.f, m1.w > 0.f ? v[7] : 0.f));
            }
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem) : "memory");
}

// k_wgrad_tc: weight (and bias) gradients, a reduction over samples:
//     D[m][n] = sum_b P[b][m] * Q[b][n]          (+ a column of ones appended
//     to Q, so D[m][ndim] = sum_b P[b][m] is the bias gradient)
// On the tensor core this is a GEMM whose K dimension is the sample index, so
// both operands must have samples contiguous in 16-byte chunks.  Each thread
// owns one row of a 64-sample tile (half the threads the P rows, half the Q
// rows) and scatters it TRANSPOSED into shared memory (K-major operands
// [m][b] and [n][b]); 8 tcgen05.mma (K = 8 samples each) accumulate the tile
// into TMEM, which keeps
// accumulating across all tiles of the CTA.  Each CTA finally writes its
// [128][npad] partial; k_wgrad_finish folds the partials in CTA order
// (deterministic).
struct WgradParams {
    const float *P; int ldp, mdim;     // A side: M = mdim (<= 128)
    const float *Q; int ldq, ndim;     // B side: N = ndim (+1 ones column), padded to npad
    int npad, add_ones;
    int64_t rows;
    float *partial;                    // [gridDim.x][128][npad]
    float scale_up;                    // k_wgrad_mn: P is multiplied by this before FP16 conversion (and divided out at read-out)
};

constexpr int WG_TILE = 64;    // samples per tile: 60 KB of operands => three CTAs per SM hide the row-load latency
struct SmemWg {
    static constexpr int AT = 0;                                  // [16][128][16 B]
    static constexpr int BT = AT + (WG_TILE / 4) * Smem::CHUNK_A; // [16][npad][16 B], npad <= 112
    static constexpr int BAR = BT + (WG_TILE / 4) * N_HID * 16;
    static constexpr int TOTAL = BAR + 16;
};

// one sample's row -> column b of a K-major operand (stride 16 B between rows)
__device__ __forceinline__ void scatter_row(uint8_t *dst, const float *row, int n, bool vec, bool valid)
{
    for (int c0 = 0; c0 < n; c0 += 32) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
            v[u] = (valid && c0 + 4 * u < n) ? load4(row, c0 + 4 * u, n, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int c = c0 + 4 * u;
            if (c + 0 < n) *reinterpret_cast<float *>(dst + (c + 0) * 16) = to_tf32(v[u].x);
            if (c + 1 < n) *reinterpret_cast<float *>(dst + (c + 1) * 16) = to_tf32(v[u].y);
            if (c + 2 < n) *reinterpret_cast<float *>(dst + (c + 2) * 16) = to_tf32(v[u].z);
            if (c + 3 < n) *reinterpret_cast<float *>(dst + (c + 3) * 16) = to_tf32(v[u].w);
        }
    }
}

// the three layers' reductions run side by side: blockIdx.y picks the layer
struct WgradBatch { WgradParams layer[3]; };

// body of k_wgrad_tc for one layer; `p` stays a reference into the kernel's
// parameter block, so its fields remain constant-bank operands (uniform)
__device__ __forceinline__ void wgrad_layer(const WgradParams &p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_slot;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + SmemWg::BAR);
    const int warp = threadIdx.x >> 5;
    const int chunk_b = p.npad * 16;
    for (int e = threadIdx.x; e < (SmemWg::BAR) / 16; e += blockDim.x)
        reinterpret_cast<float4 *>(smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t tmem = tmem_alloc_and_sync(&tmem_base_slot, warp, 128);
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t a_t = smem_u32(smem + SmemWg::AT), b_t = smem_u32(smem + SmemWg::BT);
    const uint32_t idesc = make_idesc(TILE_M, p.npad);
    uint32_t phase = 0;
    bool first = true;
    // threads 0..63 transpose the P rows of the tile's 64 samples, threads 64..127 the Q rows
    const int sample = threadIdx.x & (WG_TILE - 1), side = threadIdx.x >> 6;
    const int cb = sample >> 2, l4 = sample & 3;                // this sample's K chunk and position in it
    uint8_t *a_dst = smem + SmemWg::AT + (size_t)cb * Smem::CHUNK_A + l4 * 4;
    uint8_t *b_dst = smem + SmemWg::BT + (size_t)cb * chunk_b + l4 * 4;
    const bool pvec = (p.ldp & 3) == 0, qvec = (p.ldq & 3) == 0;
    const int64_t n_tiles = (p.rows + WG_TILE - 1) / WG_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * WG_TILE + sample;
        const bool valid = row < p.rows;
        // rows are read 8 x 16 bytes at a time (loads in flight together), then
        // scattered transposed: element (sample b, column m) -> operand row m, K slot b
        if (side == 0) {
            scatter_row(a_dst, p.P + row * p.ldp, p.mdim, pvec, valid);
        } else {
            scatter_row(b_dst, p.Q + row * p.ldq, p.ndim, qvec, valid);
            if (p.add_ones) *reinterpret_cast<float *>(b_dst + p.ndim * 16) = valid ? 1.f : 0.f;
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (threadIdx.x == 0) {
            for (int s = 0; s < WG_TILE / 8; s++)
                mma_tf32(tmem, make_desc(a_t + s * 2 * Smem::CHUNK_A, Smem::CHUNK_A, 128),
                         make_desc(b_t + s * 2 * chunk_b, chunk_b, 128), idesc, (!first || s > 0) ? 1u : 0u);
            mma_commit(mbar);
        }
        first = false;
        mbar_wait(mbar, phase); phase ^= 1;      // operands may be overwritten by the next tile
        tc_fence_after();
    }
    // this thread's accumulator row (m = threadIdx.x) -> partial
    float *out = p.partial + ((size_t)blockIdx.x * TILE_M + threadIdx.x) * p.npad;
    // rows >= mdim are never read back.  tcgen05.ld is warp-collective, so a warp
    // either loads as a whole or skips as a whole; only the store is per thread.
    const int n_c8 = warp * 32 < p.mdim ? p.npad / 8 : 0;
    const bool keep = (int)threadIdx.x < p.mdim;
    for (int c8 = 0; c8 < n_c8; c8++) {
        float v[8];
        tmem_ld8(tmem_row + c8 * 8, v);
        if (keep) {
#pragma unroll
            for (int i = 0; i < 8; i++) out[c8 * 8 + i] = v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tmem) : "memory");
}

__global__ void __launch_bounds__(TILE_M, 1) k_wgrad_tc(const __grid_constant__ WgradBatch batch)
{
    // three inlined copies selected by a uniform branch: indexing the parameter
    // block dynamically (or copying the selected entry into registers) makes every
    // pointer and loop bound a per-thread value -- measured 279 -> 492 us at batch 262,144
    if (blockIdx.y == 0) wgrad_layer(batch.layer[0]);
    else if (blockIdx.y == 1) wgrad_layer(batch.layer[1]);
    else wgrad_layer(batch.layer[2]);
}

// k_wgrad_mn: the same reduction WITHOUT the transposition.  A tile of 128
// samples is staged the way the forward stages its operands -- thread r copies
// row r in 16-byte pieces, piece c to chunk c (coalesced loads, conflict-free
// stores) -- and that very layout, read with MN-major descriptors, is the
// transposed operand the reduction needs: for element (feature f, sample r) at
// (f/8) * CHUNK + r * 16 + (f%8) * 2 the 8-feature groups are SBO = CHUNK bytes
// apart, consecutive samples (the MMA's K) 16 bytes apart and groups of 8
// samples LBO = 128 bytes apart: the canonical no-swizzle MN-major form.
// Operands are FP16 (kind::f16, fp32 accumulate): MN-major TF32 operands exist
// only in the 128B_BASE32B swizzled layout (which the K-major chain operands
// cannot share), FP16 has TF32's 10 mantissa bits, and its narrow range is
// handled by scaling the error signals up by `scale_up` (B/2, undoing the
// loss's 2/B) before conversion and back down at read-out.  8 tcgen05.mma
// (K = 16 samples each) fold a tile into the TMEM accumulator
// D[m][n] += sum_r P[r][m] * Q[r][n].  61 KB of shared memory and 128 TMEM
// columns per CTA: three CTAs per SM, loads of one overlapping MMAs of another.
constexpr int CHUNK_F16 = TILE_M * 16;      // one 16-byte piece (8 halfs) of all 128 rows
struct SmemWm {
    static constexpr int PA = 0;                              // [16][128][16 B]  P rows, 128 features (M side)
    static constexpr int QB = PA + 16 * CHUNK_F16;            // [14][128][16 B]  Q rows, 112 features (N side)
    static constexpr int BAR = QB + 14 * CHUNK_F16;
    static constexpr int TOTAL = BAR + 16;
};
constexpr int WM_THREADS = 256;

// kind::f16 instruction descriptor: F16 operands (format 0), F32 accumulate, both operands MN-major
__device__ __forceinline__ constexpr uint32_t make_idesc_f16_mn(int m, int n)
{
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// two floats -> packed half2 bits, round to nearest, saturating (no infinities)
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// row (fp32, n values) -> FP16 chunks 0 .. of the chunk-major operand, scaled; invalid rows are zeros;
// ones_at >= 0 plants a 1.0 at that feature (bias-gradient column)
__device__ __forceinline__ void stage_rows_f16(uint8_t *buf, const float *row, int n, bool vec, bool valid, int r, int ones_at,
                                               float scale)
{
    const int chunks = (n + (ones_at >= 0 ? 1 : 0) + 7) / 8;
    for (int c0 = 0; c0 < chunks; c0 += 4) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
            v[u] = (valid && 2 * c0 + u < 2 * chunks) ? load4(row, 8 * c0 + 4 * u, n, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int c = c0 + u;
            if (c >= chunks) break;
            float f[8] = {v[2 * u].x * scale, v[2 * u].y * scale, v[2 * u].z * scale, v[2 * u].w * scale,
                          v[2 * u + 1].x * scale, v[2 * u + 1].y * scale, v[2 * u + 1].z * scale, v[2 * u + 1].w * scale};
            if (valid && ones_at >= 0 && ones_at / 8 == c) f[ones_at & 7] = 1.f;
            uint4 w;
            w.x = pack_half2(f[0], f[1]); w.y = pack_half2(f[2], f[3]); w.z = pack_half2(f[4], f[5]); w.w = pack_half2(f[6], f[7]);
            *reinterpret_cast<uint4 *>(buf + (size_t)c * CHUNK_F16 + r * 16) = w;
        }
    }
}

__device__ __forceinline__ void wgrad_mn_layer(const WgradParams &p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_slot;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + SmemWm::BAR);
    const int warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < SmemWm::BAR / 16; e += blockDim.x)
        reinterpret_cast<float4 *>(smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t tmem = tmem_alloc_and_sync(&tmem_base_slot, warp, 128);
    const uint32_t a_p = smem_u32(smem + SmemWm::PA), b_q = smem_u32(smem + SmemWm::QB);
    const uint32_t idesc = make_idesc_f16_mn(TILE_M, p.npad);
    uint32_t phase = 0;
    bool first = true;
    const int r = threadIdx.x & (TILE_M - 1), side = threadIdx.x >> 7;          // threads 0..127: P rows, 128..255: Q rows
    const bool pvec = (p.ldp & 3) == 0, qvec = (p.ldq & 3) == 0;
    const int64_t n_tiles = (p.rows + TILE_M - 1) / TILE_M;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * TILE_M + r;
        const bool valid = row < p.rows;
        if (side == 0) stage_rows_f16(smem + SmemWm::PA, p.P + row * p.ldp, p.mdim, pvec, valid, r, -1, p.scale_up);
        else stage_rows_f16(smem + SmemWm::QB, p.Q + row * p.ldq, p.ndim, qvec, valid, r, p.add_ones ? p.ndim : -1, 1.f);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (threadIdx.x == 0) {
            for (int s = 0; s < TILE_M / 16; s++)         // K = 16 samples per instruction: 256 bytes further along the rows
                mma_f16(tmem, make_desc(a_p + s * 256, 128, CHUNK_F16), make_desc(b_q + s * 256, 128, CHUNK_F16), idesc,
                        (!first || s > 0) ? 1u : 0u);
            mma_commit(mbar);
        }
        first = false;
        mbar_wait(mbar, phase); phase ^= 1;               // the operands may be overwritten by the next tile
        tc_fence_after();
    }
    // accumulator row m (= feature m of P) -> partial; warps 0..3 own TMEM lanes 32 * warp ..
    if (warp < 4) {
        const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
        float *out = p.partial + ((size_t)blockIdx.x * TILE_M + threadIdx.x) * p.npad;
        const int n_c8 = warp * 32 < p.mdim ? p.npad / 8 : 0;
        const bool keep = (int)threadIdx.x < p.mdim;
        const float down = 1.f / p.scale_up;
        for (int c8 = 0; c8 < n_c8; c8++) {
            float v[8];
            tmem_ld8(tmem_row + c8 * 8, v);
            if (keep) {
#pragma unroll
                for (int i = 0; i < 8; i++) out[c8 * 8 + i] = v[i] * down;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tmem) : "memory");
}

__global__ void __launch_bounds__(WM_THREADS, 3) k_wgrad_mn(const __grid_constant__ WgradBatch batch)
{
    if (blockIdx.y == 0) wgrad_mn_layer(batch.layer[0]);
    else if (blockIdx.y == 1) wgrad_mn_layer(batch.layer[1]);
    else wgrad_mn_layer(batch.layer[2]);
}

// ===================================================================== fused backward
// k_mlp_backward_fused: the whole backward pass of a 128-sample tile without
// leaving the SM -- error chain AND the three weight / bias gradients:
//     dH2 = (dQ  * W3) .* (H2 > 0)      TF32 K-major chain (as k_mlp_backward_data_tc)
//     dH1 = (dH2 * W2) .* (H1 > 0)
//     dW3^T += [H2|1]^T dQ              FP16 MN-major reductions over the tile's samples,
//     dW2   += dH2^T [H1|1]             accumulators resident in TMEM across all tiles
//     dW1   += dH1^T [X|1]              of the CTA (the |1 column yields the bias gradient)
// H1 / H2 arrive as the FP16 operand images the forward wrote (one TMA bulk copy
// per tile and layer, 28 KB, landing directly in MMA operand layout; the next
// tile's copy is issued the moment the buffer is free), the ReLU masks are read
// from those images, dH2 / dH1 never exist outside shared memory, and each CTA
// writes one partial per gradient that k_bwd_fused_finish folds in CTA order
// (deterministic).  HBM traffic per tile: 2 x 28 KB of images + dQ + boards,
// against 3 x (105 + 105) MB read and 210 MB written per 262,144-sample batch by
// the unfused pair of kernels.
struct FusedBwdParams {
    const uint8_t *w_image;            // W3^T | W2^T, TF32 K-major B operands (BWD_IMAGE_BYTES)
    const float *dq;                   // [rows][n_out], already scaled by the loss (2/B)
    const uint8_t *h1_img, *h2_img;    // [tiles][14][128][16 B] FP16 images from k_mlp_forward_ts
    const uint8_t *boards;             // [rows][n_in] uint8
    int n_in, n_hidden, n_out, n1pad;  // n1pad: (n_in + 1) rounded up to 16 (N of the layer-1 reduction)
    int64_t rows;
    float scale_up;                    // FP16 copies of the error signals carry this factor (B/2)
    float *partial;                    // [gridDim.x][128][16 + 112 + n1pad]
};

struct SmemFb {
    static constexpr int W3T = 0;                                   // [2][112][16 B]
    static constexpr int W2T = W3T + 2 * Smem::CHUNK_H;             // [26][112][16 B]
    static constexpr int DQ = W2T + (K_HID / 4) * Smem::CHUNK_H;    // [2][128][16 B]   TF32 chain operand (K = 8)
    static constexpr int DH = DQ + 2 * Smem::CHUNK_A;               // [26][128][16 B]  TF32 chain operand (dH2)
    static constexpr int H2H = DH + (K_HID / 4) * Smem::CHUNK_A;    // [14][128][16 B]  FP16 image of H2 (M = 128 reads 2 chunks on, into H1H)
    static constexpr int H1H = H2H + H_IMG_TILE_BYTES;              // [14][128][16 B]  FP16 image of H1
    static constexpr int DHH = H1H + H_IMG_TILE_BYTES;              // [16][128][16 B]  FP16 dH2, then dH1 (M = 128)
    static constexpr int DQH = DHH + 16 * CHUNK_F16;                // [2][128][16 B]   FP16 dQ (N = 16)
    static constexpr int XH = DQH + 2 * CHUNK_F16;                  // [8][128][16 B]   FP16 boards | 1 (N <= 64)
    static constexpr int BAR = XH + 8 * CHUNK_F16;
    static constexpr int TOTAL = BAR + 64;
};
static_assert(SmemFb::DQ == (int)BWD_IMAGE_BYTES, "the backward weight image is W3^T | W2^T");
static_assert(SmemFb::TOTAL <= 227 * 1024, "fused backward: shared memory plan exceeds one SM");
constexpr int FB_THREADS = 384;      // 12 warps: 3 per TMEM lane quadrant, each masking a third of the columns (16 warps spilled)
constexpr uint32_t FB_D0 = 0, FB_D1 = 128, FB_ACC2 = 256, FB_ACC1 = 384, FB_ACC3 = 448;     // TMEM columns (512 allocated)

// error-signal epilogue of one thread: 8-column groups [g0, g1) of its accumulator row, masked by the
// FP16 activation image -> FP16 operand (x scale_up) and, for dH2, the TF32 chain operand
__device__ __forceinline__ void tmem_ld8x5(uint32_t taddr, uint32_t stride, float (&v)[40], int n_loads);

__device__ __forceinline__ void fb_mask_epilogue(uint32_t lane_base, uint32_t d_col, const uint8_t *act_img, uint8_t *dst_f16,
                                                 uint8_t *dst_tf32, int r, int g0, int g1, float scale_up)
{
    // at most five 8-column groups per warp: all of them in one TMEM round trip
    float all[40];
    tmem_ld8x5(lane_base + d_col + 8 * g0, 8, all, g1 - g0);
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const int g = g0 + j;
        if (g >= g1) break;                                    // warp-uniform
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = all[8 * j + i];
        const uint4 m = *reinterpret_cast<const uint4 *>(act_img + (size_t)g * CHUNK_F16 + r * 16);
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t hbits = (mw[i >> 1] >> (16 * (i & 1))) & 0x7FFFu;      // post-ReLU activation: > 0 iff any magnitude bit
            v[i] = hbits ? v[i] : 0.f;
        }
        uint4 w;
        w.x = pack_half2(v[0] * scale_up, v[1] * scale_up); w.y = pack_half2(v[2] * scale_up, v[3] * scale_up);
        w.z = pack_half2(v[4] * scale_up, v[5] * scale_up); w.w = pack_half2(v[6] * scale_up, v[7] * scale_up);
        *reinterpret_cast<uint4 *>(dst_f16 + (size_t)g * CHUNK_F16 + r * 16) = w;
        if (dst_tf32) {
            *reinterpret_cast<float4 *>(dst_tf32 + (size_t)(2 * g) * Smem::CHUNK_A + r * 16) =
                make_float4(to_tf32_fast(v[0]), to_tf32_fast(v[1]), to_tf32_fast(v[2]), to_tf32_fast(v[3]));
            *reinterpret_cast<float4 *>(dst_tf32 + (size_t)(2 * g + 1) * Smem::CHUNK_A + r * 16) =
                make_float4(to_tf32_fast(v[4]), to_tf32_fast(v[5]), to_tf32_fast(v[6]), to_tf32_fast(v[7]));
        }
    }
}

__global__ void __launch_bounds__(FB_THREADS, 1) k_mlp_backward_fused(const __grid_constant__ FusedBwdParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_slot;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + SmemFb::BAR);        // MMA groups
    uint64_t *mbar_w = mbar + 1, *mbar_h2 = mbar + 2, *mbar_h1 = mbar + 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quad = warp & 3, half = warp >> 2;                              // TMEM lane quadrant; column third (0..2)
    const int r = quad * 32 + lane;                                           // this thread's row of the tile
    // the 13 column groups of 8, split 5 / 4 / 4 over the three warps of a quadrant
    const int g_lo = half == 0 ? 0 : 1 + 4 * half, g_hi = half == 0 ? 5 : 5 + 4 * half;
    const int64_t n_tiles = (p.rows + TILE_M - 1) / TILE_M;
    // activation buffers start out zero: padding chunks / rows are never written again
    for (int e = threadIdx.x; e < (SmemFb::BAR - SmemFb::DQ) / 16; e += blockDim.x)
        reinterpret_cast<float4 *>(smem + SmemFb::DQ)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1); mbar_init(mbar_w, 1); mbar_init(mbar_h2, 1); mbar_init(mbar_h1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        bulk_load(smem + SmemFb::W3T, p.w_image, BWD_IMAGE_BYTES, mbar_w);
        if ((int64_t)blockIdx.x < n_tiles) {
            bulk_load(smem + SmemFb::H2H, p.h2_img + (size_t)blockIdx.x * H_IMG_TILE_BYTES, H_IMG_TILE_BYTES, mbar_h2);
            bulk_load(smem + SmemFb::H1H, p.h1_img + (size_t)blockIdx.x * H_IMG_TILE_BYTES, H_IMG_TILE_BYTES, mbar_h1);
        }
    }
    const uint32_t tmem = tmem_base_slot;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t s_w3t = smem_u32(smem + SmemFb::W3T), s_w2t = smem_u32(smem + SmemFb::W2T);
    const uint32_t s_dq = smem_u32(smem + SmemFb::DQ), s_dh = smem_u32(smem + SmemFb::DH);
    const uint32_t s_h2h = smem_u32(smem + SmemFb::H2H), s_h1h = smem_u32(smem + SmemFb::H1H);
    const uint32_t s_dhh = smem_u32(smem + SmemFb::DHH), s_dqh = smem_u32(smem + SmemFb::DQH), s_xh = smem_u32(smem + SmemFb::XH);
    constexpr uint32_t IDESC_CHAIN = make_idesc(TILE_M, N_HID);
    const uint32_t idesc3 = make_idesc_f16_mn(TILE_M, N_OUT), idesc2 = make_idesc_f16_mn(TILE_M, N_HID);
    const uint32_t idesc1 = make_idesc_f16_mn(TILE_M, p.n1pad);
    uint32_t ph_mma = 0, ph_h2 = 0, ph_h1 = 0;
    bool first = true, g3_pending = false;
    const int x_chunks = p.n1pad / 8;
    mbar_wait(mbar_w, 0);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * TILE_M + r;
        const bool valid = row < p.rows;
        // ---- this tile's dQ row (half 0) / board row (half 1) into registers, then wait for the previous
        //      tile's last reduction (it reads XH and DHH) before overwriting its operands
        float4 dqv = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t cells[16];
        if (half == 0) {
            if (valid) {
                const float *q = p.dq + row * p.n_out;
                dqv.x = q[0];
                dqv.y = p.n_out > 1 ? q[1] : 0.f;
                dqv.z = p.n_out > 2 ? q[2] : 0.f;
                dqv.w = p.n_out > 3 ? q[3] : 0.f;
            }
        } else if (half == 1) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                uint32_t w = 0;
                if (valid && 4 * j < p.n_in) {
                    const uint8_t *b = p.boards + row * p.n_in + 4 * j;
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (4 * j + i < p.n_in) w |= (uint32_t)__ldg(b + i) << (8 * i);
                }
                cells[j] = w;
            }
        }
        if (g3_pending) { mbar_wait(mbar, ph_mma); ph_mma ^= 1; tc_fence_after(); g3_pending = false; }
        if (half == 0) {
            *reinterpret_cast<float4 *>(smem + SmemFb::DQ + r * 16) = make_float4(to_tf32(dqv.x), to_tf32(dqv.y), to_tf32(dqv.z), to_tf32(dqv.w));
            uint4 w = make_uint4(pack_half2(dqv.x * p.scale_up, dqv.y * p.scale_up), pack_half2(dqv.z * p.scale_up, dqv.w * p.scale_up), 0u, 0u);
            *reinterpret_cast<uint4 *>(smem + SmemFb::DQH + r * 16) = w;
        } else if (half == 1) {
#pragma unroll
            for (int c = 0; c < 8; c++) {                 // 8 cells per FP16 chunk = two packed words
                if (c >= x_chunks) continue;
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; i++) f[i] = (float)((cells[2 * c + (i >> 2)] >> (8 * (i & 3))) & 0xFFu);
                if (valid && p.n_in / 8 == c) f[p.n_in & 7] = 1.f;
                uint4 w;
                w.x = pack_half2(f[0], f[1]); w.y = pack_half2(f[2], f[3]); w.z = pack_half2(f[4], f[5]); w.w = pack_half2(f[6], f[7]);
                *reinterpret_cast<uint4 *>(smem + SmemFb::XH + (size_t)c * CHUNK_F16 + r * 16) = w;
            }
        }
        mbar_wait(mbar_h2, ph_h2); ph_h2 ^= 1;            // H2 image of this tile has landed
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        // ---- group 1: dH2 pre-activation, and dW3^T += [H2|1]^T dQ
        if (threadIdx.x == 0) {
            mma_tf32(tmem + FB_D0, make_desc(s_dq, Smem::CHUNK_A, 128), make_desc(s_w3t, Smem::CHUNK_H, 128), IDESC_CHAIN, 0);
            for (int s = 0; s < TILE_M / 16; s++)
                mma_f16(tmem + FB_ACC3, make_desc(s_h2h + s * 256, 128, CHUNK_F16), make_desc(s_dqh + s * 256, 128, CHUNK_F16), idesc3,
                        (!first || s > 0) ? 1u : 0u);
            mma_commit(mbar);
        }
        mbar_wait(mbar, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        fb_mask_epilogue(lane_base, FB_D0, smem + SmemFb::H2H, smem + SmemFb::DHH, smem + SmemFb::DH, r, g_lo, g_hi, p.scale_up);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        const int64_t next = tile + gridDim.x;
        // ---- group 2: dH1 pre-activation, and dW2 += dH2^T [H1|1]; the H2 buffer is free for the next tile
        if (threadIdx.x == 0) {
            if (next < n_tiles) bulk_load(smem + SmemFb::H2H, p.h2_img + (size_t)next * H_IMG_TILE_BYTES, H_IMG_TILE_BYTES, mbar_h2);
            mbar_wait(mbar_h1, ph_h1);
            for (int k = 0; k < K_HID / 8; k++)
                mma_tf32(tmem + FB_D1, make_desc(s_dh + k * 2 * Smem::CHUNK_A, Smem::CHUNK_A, 128),
                         make_desc(s_w2t + k * 2 * Smem::CHUNK_H, Smem::CHUNK_H, 128), IDESC_CHAIN, k > 0);
            for (int s = 0; s < TILE_M / 16; s++)
                mma_f16(tmem + FB_ACC2, make_desc(s_dhh + s * 256, 128, CHUNK_F16), make_desc(s_h1h + s * 256, 128, CHUNK_F16), idesc2,
                        (!first || s > 0) ? 1u : 0u);
            mma_commit(mbar);
        }
        mbar_wait(mbar_h1, ph_h1); ph_h1 ^= 1;            // everyone reads the H1 image below
        mbar_wait(mbar, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        fb_mask_epilogue(lane_base, FB_D1, smem + SmemFb::H1H, smem + SmemFb::DHH, nullptr, r, g_lo, g_hi, p.scale_up);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        // ---- group 3: dW1 += dH1^T [X|1]; the H1 buffer is free for the next tile
        if (threadIdx.x == 0) {
            if (next < n_tiles) bulk_load(smem + SmemFb::H1H, p.h1_img + (size_t)next * H_IMG_TILE_BYTES, H_IMG_TILE_BYTES, mbar_h1);
            for (int s = 0; s < TILE_M / 16; s++)
                mma_f16(tmem + FB_ACC1, make_desc(s_dhh + s * 256, 128, CHUNK_F16), make_desc(s_xh + s * 256, 128, CHUNK_F16), idesc1,
                        (!first || s > 0) ? 1u : 0u);
            mma_commit(mbar);
        }
        g3_pending = true;
        first = false;
    }
    if (g3_pending) { mbar_wait(mbar, ph_mma); ph_mma ^= 1; tc_fence_after(); }
    // ---- accumulators -> this CTA's partial: row m, columns [acc3 (16) | acc2 (112) | acc1 (n1pad)], scaled back
    if (!first) {
        const int stride = N_OUT + N_HID + p.n1pad;
        float *out = p.partial + ((size_t)blockIdx.x * TILE_M + r) * stride;
        const float down = 1.f / p.scale_up;
        const int c_lo = half >= 2 ? stride : half ? (N_OUT + N_HID) / 2 : 0;                            // multiples of 8
        const int c_hi = half >= 2 ? stride : half ? stride : (N_OUT + N_HID) / 2;                       // (quarters 2, 3 idle)
        for (int c = c_lo; c < c_hi; c += 8) {
            const uint32_t col = c < N_OUT ? FB_ACC3 + c : c < N_OUT + N_HID ? FB_ACC2 + (c - N_OUT) : FB_ACC1 + (c - N_OUT - N_HID);
            float v[8];
            tmem_ld8(lane_base + col, v);
            *reinterpret_cast<float4 *>(out + c) = make_float4(v[0] * down, v[1] * down, v[2] * down, v[3] * down);
            *reinterpret_cast<float4 *>(out + c + 4) = make_float4(v[4] * down, v[5] * down, v[6] * down, v[7] * down);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

constexpr int WGF_ELEMS = 32, WGF_LANES = 8;
// partials -> gradients, torch layout.  A block folds 32 parameters: 8 lanes of partials (g = lane, lane + 8,
// ...) per parameter, then the 8 lane sums in lane order -- a fixed order, so the result does not depend on
// scheduling (same scheme as k_wgrad_finish).
struct FusedFinish {
    const float *partial; int n_partials, n_in, n_hidden, n_out, n1pad;
    float *dW1, *db1, *dW2, *db2, *dW3, *db3;
};
__global__ void __launch_bounds__(WGF_ELEMS * WGF_LANES) k_bwd_fused_finish(const FusedFinish f)
{
    __shared__ float part[WGF_LANES][WGF_ELEMS];
    const int H = f.n_hidden, A = f.n_out, I = f.n_in;
    const int n1 = H * (I + 1), n2 = H * (H + 1), n3 = A * (H + 1);
    const int lane_g = threadIdx.x / WGF_ELEMS, el = threadIdx.x & (WGF_ELEMS - 1);
    const int e_raw = blockIdx.x * WGF_ELEMS + el;
    const bool live = e_raw < n1 + n2 + n3;
    const int e = live ? e_raw : 0;
    const int stride = N_OUT + N_HID + f.n1pad;
    int m, col;
    float *dst;
    if (e < n1) {                                     // dW1[h][i] | db1[h]: accumulator 1, row h, column i (bias: column n_in)
        m = e / (I + 1); const int i = e - m * (I + 1);
        col = N_OUT + N_HID + i;
        dst = i < I ? f.dW1 + (size_t)m * I + i : f.db1 + m;
    } else if (e < n1 + n2) {                         // dW2[o][h] | db2[o]
        const int k = e - n1;
        m = k / (H + 1); const int h = k - m * (H + 1);
        col = N_OUT + h;
        dst = h < H ? f.dW2 + (size_t)m * H + h : f.db2 + m;
    } else {                                          // dW3[a][h] | db3[a]: accumulator 3 holds the TRANSPOSE, row h (bias: row n_hidden)
        const int k = e - n1 - n2;
        const int a = k / (H + 1), h = k - a * (H + 1);
        m = h; col = a;
        dst = h < H ? f.dW3 + (size_t)a * H + h : f.db3 + a;
    }
    const float *src = f.partial + (size_t)m * stride + col;
    const size_t step = (size_t)TILE_M * stride;
    float s0 = 0.f, s1 = 0.f;
    if (live) {
        int g = lane_g;
        for (; g + WGF_LANES < f.n_partials; g += 2 * WGF_LANES) {
            s0 += src[(size_t)g * step];
            s1 += src[(size_t)(g + WGF_LANES) * step];
        }
        if (g < f.n_partials) s0 += src[(size_t)g * step];
    }
    part[lane_g][el] = s0 + s1;
    __syncthreads();
    if (lane_g == 0 && live) {
        float s = 0.f;
#pragma unroll
        for (int l = 0; l < WGF_LANES; l++) s += part[l][el];
        *dst = s;
    }
}

// dW[m][n] = sum_g partial[g][m][n], db[m] = sum_g partial[g][m][ndim].
// A block folds 32 elements: 8 lanes of partials (g = lane, lane + 8, ...) per
// element, then the 8 lane sums in lane order -- a fixed order, so the result
// does not depend on scheduling.
struct WgradFinish { const float *partial; int npad, mdim, ndim; float *dW, *db; };
struct WgradFinishBatch { WgradFinish layer[3]; };

__global__ void __launch_bounds__(WGF_ELEMS * WGF_LANES)
k_wgrad_finish(const WgradFinishBatch batch, int n_partials)
{
    __shared__ float part[WGF_LANES][WGF_ELEMS];
    WgradFinish f = batch.layer[0];
    if (blockIdx.y == 1) f = batch.layer[1];
    if (blockIdx.y == 2) f = batch.layer[2];
    const int e = threadIdx.x & (WGF_ELEMS - 1), lane = threadIdx.x / WGF_ELEMS;
    const int idx = blockIdx.x * WGF_ELEMS + e;
    const int total = f.mdim * (f.ndim + 1);
    if (blockIdx.x * WGF_ELEMS >= total) return;          // whole block beyond this layer's elements
    const bool live = idx < total;
    const int m = live ? idx / (f.ndim + 1) : 0, n = live ? idx - m * (f.ndim + 1) : 0;
    const float *src = f.partial + (size_t)m * f.npad + n;
    const size_t stride = (size_t)TILE_M * f.npad;
    float s0 = 0.f, s1 = 0.f;
    if (live) {
        int g = lane;
        for (; g + WGF_LANES < n_partials; g += 2 * WGF_LANES) {
            s0 += src[g * stride];
            s1 += src[(g + WGF_LANES) * stride];
        }
        if (g < n_partials) s0 += src[g * stride];
    }
    part[lane][e] = s0 + s1;
    __syncthreads();
    if (lane == 0 && live) {
        float s = 0.f;
#pragma unroll
        for (int l = 0; l < WGF_LANES; l++) s += part[l][e];
        if (n < f.ndim) f.dW[(size_t)m * f.ndim + n] = s;
        else if (f.db) f.db[m] = s;
    }
}

// Weights -> operand images in global memory, once per parameter update; the
// MLP kernels then fetch an image with one bulk copy instead of re-staging the
// matrices in every CTA.  Block b packs one matrix.
// The forward image is [hi | lo]: two copies of the W1|W2|W3 operand layout, the
// TF32-rounded weights and the TF32-rounded remainders (3xTF32, k_mlp_forward_ts).
__global__ void __launch_bounds__(256) k_pack_weights(const float *w1, const float *w2, const float *w3, int n_in, int n_hidden,
                                                      int n_out, uint8_t *fwd_image, uint8_t *bwd_image)
{
    const int k_in = (n_in + 7) & ~7;
    const int tid = blockIdx.y * blockDim.x + threadIdx.x, nth = gridDim.y * blockDim.x;    // gridDim.y blocks share a matrix
    uint8_t *lo = fwd_image + FWD_IMAGE_BYTES;
    switch (blockIdx.x) {
    case 0: stage_weights(fwd_image + Smem::W1, w1, n_hidden, n_in, N_HID, k_in, tid, nth); break;
    case 1: stage_weights(fwd_image + Smem::W2, w2, n_hidden, n_hidden, N_HID, K_HID, tid, nth); break;
    case 2: stage_weights(fwd_image + Smem::W3, w3, n_out, n_hidden, N_OUT, K_HID, tid, nth); break;
    case 3: stage_weights(lo + Smem::W1, w1, n_hidden, n_in, N_HID, k_in, tid, nth, 1); break;
    case 4: stage_weights(lo + Smem::W2, w2, n_hidden, n_hidden, N_HID, K_HID, tid, nth, 1); break;
    case 5: stage_weights(lo + Smem::W3, w3, n_out, n_hidden, N_OUT, K_HID, tid, nth, 1); break;
    case 6: if (bwd_image) stage_weights_T(bwd_image + SmemBwd::W3T, w3, n_out, n_hidden, N_HID, 8, tid, nth); break;
    default: if (bwd_image) stage_weights_T(bwd_image + SmemBwd::W2T, w2, n_hidden, n_hidden, N_HID, K_HID, tid, nth); break;
    }
}

// ===================================================================== forward, fp32-accurate (3xTF32)
// k_mlp_forward_ts: the same fused three-layer forward with
//   * 3xTF32: every fp32 operand is split into two TF32 numbers, x = hi + lo,
//     and a product a*b is accumulated as a_hi*b_hi + a_hi*b_lo + a_lo*b_hi in
//     the fp32 TMEM accumulator (the dropped lo*lo term is 2^-22 relative):
//     Q values agree with torch fp32 to ~1e-6, inside the 1e-5 of north_star.
//     Boards are small integers, exact in TF32, so layer 1 needs two passes;
//   * activations that never leave tensor memory: the epilogue reads its
//     accumulator row (tcgen05.ld), applies bias + ReLU, splits hi / lo and
//     writes both back to TMEM (tcgen05.st), where the next layer's MMAs take
//     them as the A operand (tcgen05.mma with A in TMEM) -- shared memory
//     holds only the two weight images (160 KB) and the board tile;
//   * 16 warps drain the accumulator: warp w owns TMEM lanes 32*(w%4).., and
//     the four warps of a lane quadrant split the columns 32 / 32 / 32 / 16;
//     each issues all its tcgen05.ld before one wait, so the epilogue is one
//     TMEM round trip deep; the next tile's boards are fetched while the
//     current tile computes.  (Tried and rejected: a 17th warp issuing the
//     next layer's K steps as the epilogue delivers its 8-column groups, one
//     mbarrier per group.  The per-group tcgen05.wait::st + arrive and the
//     fewer epilogue warps that fit the register file made the epilogue, the
//     long pole at 3.8 of 5.7 us per tile, slower than the hidden MMA time
//     gained: 2.8e9 boards/s against 3.3e9.)
struct SmemTs {
    static constexpr int HI = 0;                                   // W1|W2|W3 (layout of Smem), TF32-rounded
    static constexpr int LO = FWD_IMAGE_BYTES;                     // ... remainders
    static constexpr int X = 2 * FWD_IMAGE_BYTES;                  // [MAX_K_IN/4][128][16 B]
    static constexpr int BIAS = X + (MAX_K_IN / 4) * Smem::CHUNK_A;
    static constexpr int BAR = BIAS + (2 * N_HID + N_OUT) * 4;
    static constexpr int TOTAL = BAR + 32;
};
constexpr int TS_THREADS = 512;      // 16 warps: 4 per TMEM lane quadrant, each draining a quarter of the columns
constexpr uint32_t TS_D_A = 0, TS_D_B = 128, TS_H_HI = 256, TS_H_LO = 384;     // TMEM columns (512 allocated)

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                    "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                    "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

// up to five 8-column loads (`stride` columns apart) in flight, then one wait (a second, already
// satisfied one carries the last eight registers: an asm statement takes at most 32 of them)
__device__ __forceinline__ void tmem_ld8x5(uint32_t taddr, uint32_t stride, float (&v)[40], int n_loads)
{
    uint32_t r[40];
#pragma unroll
    for (int i = 0; i < 40; i++) r[i] = 0u;
#pragma unroll
    for (int j = 0; j < 5; j++)
        if (j < n_loads)          // warp-uniform
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[8 * j + 0]), "=r"(r[8 * j + 1]), "=r"(r[8 * j + 2]), "=r"(r[8 * j + 3]),
                           "=r"(r[8 * j + 4]), "=r"(r[8 * j + 5]), "=r"(r[8 * j + 6]), "=r"(r[8 * j + 7])
                         : "r"(taddr + stride * j));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[32]), "+r"(r[33]), "+r"(r[34]), "+r"(r[35]), "+r"(r[36]), "+r"(r[37]), "+r"(r[38]), "+r"(r[39])
                 :: "memory");
#pragma unroll
    for (int i = 0; i < 40; i++) v[i] = __uint_as_float(r[i]);
}

// hidden-layer epilogue of one thread: columns [c0, c1) of its accumulator row ->
// bias, ReLU -> hi / lo TF32 parts into the A-operand region of TMEM (+ fp32 copy in HBM)
// two 16-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32], bool second)
{
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = 0u;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    if (second)           // warp-uniform
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr + 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi);

template <bool X3>
__device__ __forceinline__ void ts_hidden_epilogue(uint32_t lane_base, uint32_t d_col, const float *bias, int quarter,
                                                   float *h_out_row, int n_hidden, uint8_t *img_tile = nullptr,
                                                   int row_in_tile = 0, bool valid = true)
{
    // this warp's columns: [32 * quarter, +32), the last quarter [96, 112)
    const int c0 = 32 * quarter;
    const bool wide = quarter < 3;
    const bool vec = (n_hidden & 3) == 0;
    float v[32];
    tmem_ld32(lane_base + d_col + c0, v, wide);
    // bias as four-wide shared-memory loads (the last quarter reads 16 floats past its 16 columns: still
    // inside the bias block, and those lanes of v are never stored)
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 b4 = *reinterpret_cast<const float4 *>(bias + c0 + i);
        v[i] = fmaxf(v[i] + b4.x, 0.f); v[i + 1] = fmaxf(v[i + 1] + b4.y, 0.f);
        v[i + 2] = fmaxf(v[i + 2] + b4.z, 0.f); v[i + 3] = fmaxf(v[i + 3] + b4.w, 0.f);
    }
    if (h_out_row) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
            if (wide || i < 16) store4(h_out_row, c0 + i, n_hidden, vec, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
    }
    if (img_tile) {
        // FP16 image of the tile for the fused backward: chunk (c0 / 8 + g) holds features c0 + 8g .. + 7 of
        // all 128 rows, 16 bytes per row -- a warp's 32 rows are 512 contiguous bytes (coalesced)
#pragma unroll
        for (int g = 0; g < 4; g++) {
            if (c0 + 8 * g >= N_HID) break;
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = valid ? v[8 * g + i] : 0.f;
            if (valid && n_hidden >= c0 + 8 * g && n_hidden < c0 + 8 * g + 8) f[n_hidden - (c0 + 8 * g)] = 1.f;
            uint4 w;
            w.x = pack_half2(f[0], f[1]); w.y = pack_half2(f[2], f[3]); w.z = pack_half2(f[4], f[5]); w.w = pack_half2(f[6], f[7]);
            *reinterpret_cast<uint4 *>(img_tile + (size_t)(c0 / 8 + g) * (TILE_M * 16) + row_in_tile * 16) = w;
        }
    }
#pragma unroll
    for (int g = 0; g < 4; g++) {
        if (c0 + 8 * g >= K_HID) break;               // the 112-column accumulator has 104 operand columns (warp-uniform)
        float hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            hi[i] = to_tf32_fast(v[8 * g + i]);
            lo[i] = X3 ? to_tf32_fast(v[8 * g + i] - hi[i]) : 0.f;
        }
        tmem_st8(lane_base + TS_H_HI + c0 + 8 * g, hi);
        if (X3) tmem_st8(lane_base + TS_H_LO + c0 + 8 * g, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <bool X3>
__global__ void __launch_bounds__(TS_THREADS, 1) k_mlp_forward_ts(const Params p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_slot;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + SmemTs::BAR);
    uint64_t *mbar_w = mbar + 1;
    float *bias = reinterpret_cast<float *>(smem + SmemTs::BIAS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quad = warp & 3, half = warp >> 2;          // TMEM lane quadrant; column quarter (0..3)
    const int row_in_tile = quad * 32 + lane;
    const int k_in = (p.n_in + 7) & ~7;                   // 36 -> 40, 25 -> 32, 63 -> 64
    const int n_chunks = k_in / 4;

    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        mbar_init(mbar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        bulk_load(smem + SmemTs::HI, p.w_image, X3 ? 2 * FWD_IMAGE_BYTES : FWD_IMAGE_BYTES, mbar_w);
    }
    for (int i = threadIdx.x; i < 2 * N_HID + N_OUT; i += blockDim.x) {
        float b = 0.f;
        if (i < N_HID) b = i < p.n_hidden ? p.b1[i] : 0.f;
        else if (i < 2 * N_HID) b = i - N_HID < p.n_hidden ? p.b2[i - N_HID] : 0.f;
        else b = i - 2 * N_HID < p.n_out ? p.b3[i - 2 * N_HID] : 0.f;
        bias[i] = b;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t a_x = smem_u32(smem + SmemTs::X);
    const uint32_t w_hi = smem_u32(smem + SmemTs::HI), w_lo = smem_u32(smem + SmemTs::LO);
    constexpr uint32_t IDESC_HID = make_idesc(TILE_M, N_HID), IDESC_OUT = make_idesc(TILE_M, N_OUT);
    uint32_t phase = 0;

    // this thread's share of a board row: chunks quarter, quarter + 4, ... (4 cells each), kept in registers
    uint32_t cells[MAX_K_IN / 16];
    auto fetch = [&](int64_t tile) {
        const int64_t row = tile * TILE_M + row_in_tile;
#pragma unroll
        for (int j = 0; j < MAX_K_IN / 16; j++) {
            const int c = half + 4 * j;
            uint32_t w = 0;
            if (c < n_chunks && row < p.rows) {
                const uint8_t *b = p.boards + row * p.n_in + 4 * c;
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (4 * c + i < p.n_in) w |= (uint32_t)__ldg(b + i) << (8 * i);
            }
            cells[j] = w;
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int j = 0; j < MAX_K_IN / 16; j++) {
            const int c = half + 4 * j;
            if (c < n_chunks) {
                const uint32_t w = cells[j];
                *reinterpret_cast<float4 *>(smem + SmemTs::X + (size_t)c * Smem::CHUNK_A + row_in_tile * 16) =
                    make_float4((float)(w & 0xFFu), (float)((w >> 8) & 0xFFu), (float)((w >> 16) & 0xFFu), (float)(w >> 24));
            }
        }
    };
    const int64_t n_tiles = (p.rows + TILE_M - 1) / TILE_M;
    if ((int64_t)blockIdx.x < n_tiles) { fetch(blockIdx.x); stage(); }
    mbar_wait(mbar_w, 0);                                  // weight images have landed
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * TILE_M + row_in_tile;
        const bool valid = row < p.rows;
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        // ---- layer 1: D_a = X * W1^T  (X exact in TF32: hi and lo weight passes)
        if (threadIdx.x == 0) {
            for (int k = 0; k < k_in / 8; k++)
                mma_tf32(tmem + TS_D_A, make_desc(a_x + k * 2 * Smem::CHUNK_A, Smem::CHUNK_A, 128),
                         make_desc(w_hi + Smem::W1 + k * 2 * Smem::CHUNK_H, Smem::CHUNK_H, 128), IDESC_HID, k > 0);
            if (X3)
                for (int k = 0; k < k_in / 8; k++)
                    mma_tf32(tmem + TS_D_A, make_desc(a_x + k * 2 * Smem::CHUNK_A, Smem::CHUNK_A, 128),
                             make_desc(w_lo + Smem::W1 + k * 2 * Smem::CHUNK_H, Smem::CHUNK_H, 128), IDESC_HID, 1);
            mma_commit(mbar);
        }
        const bool more = tile + gridDim.x < n_tiles;
        if (more) fetch(tile + gridDim.x);                 // global loads in flight under the MMAs and epilogues
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        if (more) stage();                                 // layer 1 is done with the board tile
        ts_hidden_epilogue<X3>(lane_base, TS_D_A, bias, half, (valid && p.h1_out) ? p.h1_out + row * p.n_hidden : nullptr, p.n_hidden,
                               p.h1_img ? p.h1_img + tile * H_IMG_TILE_BYTES : nullptr, row_in_tile, valid);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        // ---- layer 2: D_b = H1 * W2^T, A from tensor memory
        if (threadIdx.x == 0) {
            for (int k = 0; k < K_HID / 8; k++)
                mma_tf32_ts(tmem + TS_D_B, tmem + TS_H_HI + 8 * k, make_desc(w_hi + Smem::W2 + k * 2 * Smem::CHUNK_H, Smem::CHUNK_H, 128), IDESC_HID, k > 0);
            if (X3) {
                for (int k = 0; k < K_HID / 8; k++)
                    mma_tf32_ts(tmem + TS_D_B, tmem + TS_H_HI + 8 * k, make_desc(w_lo + Smem::W2 + k * 2 * Smem::CHUNK_H, Smem::CHUNK_H, 128), IDESC_HID, 1);
                for (int k = 0; k < K_HID / 8; k++)
                    mma_tf32_ts(tmem + TS_D_B, tmem + TS_H_LO + 8 * k, make_desc(w_hi + Smem::W2 + k * 2 * Smem::CHUNK_H, Smem::CHUNK_H, 128), IDESC_HID, 1);
            }
            mma_commit(mbar);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        ts_hidden_epilogue<X3>(lane_base, TS_D_B, bias + N_HID, half, (valid && p.h2_out) ? p.h2_out + row * p.n_hidden : nullptr, p.n_hidden,
                               p.h2_img ? p.h2_img + tile * H_IMG_TILE_BYTES : nullptr, row_in_tile, valid);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        // ---- layer 3: D_a[:, 0:16] = H2 * W3^T
        if (threadIdx.x == 0) {
            for (int k = 0; k < K_HID / 8; k++)
                mma_tf32_ts(tmem + TS_D_A, tmem + TS_H_HI + 8 * k, make_desc(w_hi + Smem::W3 + k * 2 * Smem::CHUNK_O, Smem::CHUNK_O, 128), IDESC_OUT, k > 0);
            if (X3) {
                for (int k = 0; k < K_HID / 8; k++)
                    mma_tf32_ts(tmem + TS_D_A, tmem + TS_H_HI + 8 * k, make_desc(w_lo + Smem::W3 + k * 2 * Smem::CHUNK_O, Smem::CHUNK_O, 128), IDESC_OUT, 1);
                for (int k = 0; k < K_HID / 8; k++)
                    mma_tf32_ts(tmem + TS_D_A, tmem + TS_H_LO + 8 * k, make_desc(w_hi + Smem::W3 + k * 2 * Smem::CHUNK_O, Smem::CHUNK_O, 128), IDESC_OUT, 1);
            }
            mma_commit(mbar);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        if (half == 0) {
            float v[8];
            tmem_ld8(lane_base + TS_D_A, v);
            if (valid) {
                const float *b3 = bias + 2 * N_HID;
                if (p.n_out == 4) {
                    *reinterpret_cast<float4 *>(p.q_out + row * 4) = make_float4(v[0] + b3[0], v[1] + b3[1], v[2] + b3[2], v[3] + b3[3]);
                } else {
                    for (int i = 0; i < p.n_out && i < 8; i++) p.q_out[row * p.n_out + i] = v[i] + b3[i];
                }
            }
        }
        // the loop's first barrier orders this read of D_a before the next tile's layer 1
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

}  // namespace tc
