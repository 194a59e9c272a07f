// sgk_dqn.cu -- the deep-Q agent of the rollout path.
//
// Replaces DeepQAgent (safe_grid_agents/common/agents/value.py:61-187), its
// ReplayBuffer (common/utils/contain.py:8-22), dqn_warmup
// (common/warmup.py:8-23) and the dqn_learn body (common/learn.py:29-58) for N
// lock-step environments sharing ONE Q network:
//
//   per lock-step:  Q(s) for all N boards -> epsilon-greedy action ->
//                   env.step -> append N transitions to the HBM replay ring ->
//                   one optimiser step on a batch sampled from the ring
//                   (online forward, target forward, loss, backward,
//                   clip_grad_norm_(10), Adam(amsgrad)) -> target sync every
//                   sync_every lock-steps.
//
// This file is the fp32 reference-semantics implementation: every GEMM is a
// plain FFMA tile kernel, so results track torch's fp32 to rounding.  The
// tcgen05 tensor-core forward lives in sgk_mlp_tc.cuh and is checked against
// this path.
#include <math.h>
#include <string.h>

#include <new>
#include <vector>

#include "sgk_internal.cuh"
#include "sgk_mlp_tc.cuh"

#define DQN_MAX_LAYERS 6   // linear layers

// What changes from one lock-step to the next.  The direct path passes these by
// value; the CUDA-graph path (sgk_rollout_dqn) uploads one entry per lock-step
// up front and every replay of the captured graph reads entry `*cursor`.
struct StepVars {
    uint64_t step;             // agent step t
    unsigned long long thr;    // explore threshold of step t
    int64_t pos0;              // ring position the step's transitions are written at
    int64_t fill;              // filled part of the ring when the step samples
    float bc1, bc2_sqrt;       // Adam bias corrections of the step's update
};

// ===================================================================== object
struct sgk_dqn {
    int device, kind, hw, n_actions;
    int n_linear;                     // n_layers + 1 (value.py:148-158)
    int dims[DQN_MAX_LAYERS + 1];     // n_in, hidden..., n_actions
    int64_t w_off[DQN_MAX_LAYERS], b_off[DQN_MAX_LAYERS], n_params;
    float *params[2];                 // 0 = Q, 1 = target_Q
    float *grads, *adam_m, *adam_v, *adam_vmax;
    int64_t adam_step;
    // replay ring (contain.py:8-22): packed uint8 boards, SoA
    int64_t cap, count;
    uint8_t *r_s, *r_s2, *r_a, *r_term;
    float *r_r;
    // hyper-parameters (value.py:64-87, agent_parser_configs.yaml:26-63)
    double lr, discount, epsilon;
    int64_t anneal, sync_every, batch;
    int bxb_loss;                     // reproduce the reference's B x B broadcast MSE
    uint64_t seed;
    // work buffers
    int64_t rows_cap;                 // rows the activation buffers can hold
    float *x, *x2, *act[DQN_MAX_LAYERS], *act_t[DQN_MAX_LAYERS], *dact[3], *y, *scalars;
    uint8_t *b_a, *b_term; float *b_r; int64_t *b_idx;
    float *partials; int64_t partials_cap;
    float *q_env; uint8_t *boards_env; int64_t env_rows;
    unsigned long long *thr; int64_t thr_cap;
    int *status;
    double *loss_partial;             // [LOSS_BLOCKS][5]
    int use_tc;                       // forward passes on tcgen05 (TF32) instead of fp32 FFMA
    uint8_t *xb, *xb2;                // uint8 copies of the staged batch (tensor-core input)
    int sm_count;
    // graph replay of the lock-step: per-step variables on the device + cursor
    StepVars *sv_table; int64_t sv_cap;
    int *sv_cursor;
    const StepVars *sv_use;           // non-null while a lock-step is being captured: kernels read sv_use[*sv_cursor]
    cudaStream_t cap_stream;
    // tensor-core path: weights pre-packed into the kernels' shared-memory operand layout
    uint8_t *w_image[2];              // forward image of Q / target_Q
    uint8_t *w_image_bwd;             // W3^T | W2^T of Q
    int w_image_dirty[2];             // parameters changed since the image was packed
    uint8_t *h_img[2];                // FP16 operand images of H1 / H2 (forward -> fused backward)
    int64_t h_img_tiles;
    float *fb_partial; int64_t fb_partial_cap;     // per-CTA gradient partials of the fused backward
};

static const int SPLITS = 256;   // upper bound of the batch splits of the weight-gradient GEMMs

// ===================================================================== kernels
// C[M,N] = opA[M,K] * opB[K,N], fp32 FFMA, 128x128x8 block tiles, 8x8 per thread.
//   MODE 0 (NT): A[M,K] row-major, B given as W[N,K] row-major        (forward)
//   MODE 1 (NN): A[M,K] row-major, B[K,N] row-major                   (dX = dY W)
//   MODE 2 (TN): A given as [K,M] row-major, B[K,N] row-major         (dW = dY^T X)
// blockIdx.z splits K; split z writes C + z * M * ldc (partials, reduced later
// in index order => deterministic).  Epilogue: + bias[n]; relu; * (mask[m,n] > 0).
#define GEMM_BM 128
#define GEMM_BN 128
#define GEMM_BK 8
template <int MODE>
__global__ void __launch_bounds__(256) k_gemm(int M, int N, int K, const float *__restrict__ A, int lda,
                                              const float *__restrict__ B, int ldb, float *__restrict__ C, int ldc,
                                              const float *__restrict__ bias, int relu, const float *__restrict__ mask,
                                              int ldmask, int k_chunk)
{
    __shared__ __align__(16) float As[GEMM_BK][GEMM_BM + 4], Bs[GEMM_BK][GEMM_BN + 4];
    const int bm = blockIdx.y * GEMM_BM, bn = blockIdx.x * GEMM_BN;
    const int k0 = blockIdx.z * k_chunk, k1 = min(K, k0 + k_chunk);
    C += (size_t)blockIdx.z * M * ldc;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[8][8] = {};
    for (int kt = k0; kt < k1; kt += GEMM_BK) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int e = threadIdx.x + r * 256;      // 0 .. 1023 = 8 x 128
            int mm, kk;
            if (MODE == 2) { mm = e & 127; kk = e >> 7; } else { mm = e >> 3; kk = e & 7; }
            const int m = bm + mm, k = kt + kk;
            float v = 0.f;
            if (m < M && k < k1) v = MODE == 2 ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k];
            As[kk][mm] = v;
            int nn, kb;
            if (MODE == 0) { nn = e >> 3; kb = e & 7; } else { nn = e & 127; kb = e >> 7; }
            const int n = bn + nn, k2 = kt + kb;
            float w = 0.f;
            if (n < N && k2 < k1) w = MODE == 0 ? B[(size_t)n * ldb + k2] : B[(size_t)k2 * ldb + n];
            Bs[kb][nn] = w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GEMM_BK; kk++) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int m = bm + ty * 8 + i, n = bn + tx * 8 + j;
            if (m < M && n < N) {
                float v = acc[i][j];
                if (bias) v += bias[n];
                if (relu) v = fmaxf(v, 0.f);
                if (mask) v = mask[(size_t)m * ldmask + n] > 0.f ? v : 0.f;
                C[(size_t)m * ldc + n] = v;
            }
        }
}

// Small problems (few 128x128 tiles): 64x64x16 tiles, 4x4 per thread, so that
// enough CTAs exist to fill the machine.
template <int MODE>
__global__ void __launch_bounds__(256) k_gemm_small(int M, int N, int K, const float *__restrict__ A, int lda,
                                                    const float *__restrict__ B, int ldb, float *__restrict__ C, int ldc,
                                                    const float *__restrict__ bias, int relu, const float *__restrict__ mask,
                                                    int ldmask, int k_chunk)
{
    __shared__ float As[16][65], Bs[16][65];
    const int bm = blockIdx.y * 64, bn = blockIdx.x * 64;
    const int k0 = blockIdx.z * k_chunk, k1 = min(K, k0 + k_chunk);
    C += (size_t)blockIdx.z * M * ldc;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int kt = k0; kt < k1; kt += 16) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int e = threadIdx.x + r * 256;
            int mm, kk;
            if (MODE == 2) { mm = e & 63; kk = e >> 6; } else { mm = e >> 4; kk = e & 15; }
            const int m = bm + mm, k = kt + kk;
            float v = 0.f;
            if (m < M && k < k1) v = MODE == 2 ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k];
            As[kk][mm] = v;
            int nn, kb;
            if (MODE == 0) { nn = e >> 4; kb = e & 15; } else { nn = e & 63; kb = e >> 6; }
            const int n = bn + nn, k2 = kt + kb;
            float w = 0.f;
            if (n < N && k2 < k1) w = MODE == 0 ? B[(size_t)n * ldb + k2] : B[(size_t)k2 * ldb + n];
            Bs[kb][nn] = w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int m = bm + ty * 4 + i, n = bn + tx * 4 + j;
            if (m < M && n < N) {
                float v = acc[i][j];
                if (bias) v += bias[n];
                if (relu) v = fmaxf(v, 0.f);
                if (mask) v = mask[(size_t)m * ldmask + n] > 0.f ? v : 0.f;
                C[(size_t)m * ldc + n] = v;
            }
        }
}

// out[i] = sum over z of partial[z][i] in index order (deterministic)
__global__ void k_reduce_splits(const float *partial, float *out, int64_t n, int splits)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int z = 0;
    for (; z + 3 < splits; z += 4) {
        s0 += partial[(size_t)z * n + i];
        s1 += partial[(size_t)(z + 1) * n + i];
        s2 += partial[(size_t)(z + 2) * n + i];
        s3 += partial[(size_t)(z + 3) * n + i];
    }
    for (; z < splits; z++) s0 += partial[(size_t)z * n + i];
    out[i] = (s0 + s1) + (s2 + s3);
}

// column sums of dY[rows, n] for the bias gradients: partial[z][n].  Many row
// chunks (COL_SPLITS) and four independent accumulators per thread: the loop
// is latency-bound otherwise.  Fixed chunking and order => deterministic.
static const int COL_SPLITS = 1024;
__global__ void k_colsum_partial(const float *dy, int rows, int n, int ld, float *partial, int chunk)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int r0 = blockIdx.y * chunk, r1 = min(rows, r0 + chunk);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int r = r0;
    for (; r + 3 < r1; r += 4) {
        s0 += dy[(size_t)r * ld + c];
        s1 += dy[(size_t)(r + 1) * ld + c];
        s2 += dy[(size_t)(r + 2) * ld + c];
        s3 += dy[(size_t)(r + 3) * ld + c];
    }
    for (; r < r1; r++) s0 += dy[(size_t)r * ld + c];
    partial[(size_t)blockIdx.y * n + c] = (s0 + s1) + (s2 + s3);
}

__global__ void k_f64_to_f32(const double *in, float *out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

__global__ void k_u8_to_f32(const uint8_t *in, float *out, int64_t n)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        out[k] = (float)in[k];
}

// ReplayBuffer.sample (contain.py:19-22): `batch` indices uniform with
// replacement over the filled part of the ring; gathers the batch.
__global__ void k_replay_sample(uint64_t seed, uint64_t step, int64_t fill, int64_t batch, int hw,
                                const uint8_t *r_s, const uint8_t *r_s2, const uint8_t *r_a, const float *r_r,
                                const uint8_t *r_term, float *x, float *x2, uint8_t *xb, uint8_t *xb2, uint8_t *b_a, float *b_r,
                                uint8_t *b_term, int64_t *b_idx, int copy_boards, const StepVars *sv, const int *cursor)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (sv) { step = sv[*cursor].step; fill = sv[*cursor].fill; }
    uint32_t w[4];
    philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), (uint32_t)step, 0x40u | ((uint32_t)((step >> 32) & 0xFFFFFF) << 8),
                  (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x5bd1e995u, w);
    const uint64_t u = ((uint64_t)w[0] << 32) | w[1];
    const int64_t idx = (int64_t)(u % (uint64_t)fill);
    b_idx[b] = idx;
    for (int c = 0; copy_boards && c < hw; c++) {
        const uint8_t v = r_s[idx * hw + c], v2 = r_s2[idx * hw + c];
        x[b * hw + c] = (float)v; x2[b * hw + c] = (float)v2;
        xb[b * hw + c] = v; xb2[b * hw + c] = v2;
    }
    b_a[b] = r_a[idx];
    b_r[b] = r_r[idx];
    b_term[b] = r_term[idx];
}

// second stage of the sample for boards whose size is a multiple of 4 bytes:
// one thread per (sample, 4-byte word) -- coalesced word copies instead of a
// per-thread byte loop.  b_idx holds the ring index of every sample.
__global__ void k_replay_gather_words(const int64_t *b_idx, int64_t batch, int words, const uint32_t *r_s, const uint32_t *r_s2,
                                      uint32_t *xb, uint32_t *xb2, float4 *x, float4 *x2)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * words) return;
    const int64_t b = e / words;
    const int w = (int)(e - b * words);
    const int64_t idx = b_idx[b];
    const uint32_t v = r_s[idx * words + w], v2 = r_s2[idx * words + w];
    xb[e] = v; xb2[e] = v2;
    x[e] = make_float4((float)(v & 0xFF), (float)((v >> 8) & 0xFF), (float)((v >> 16) & 0xFF), (float)(v >> 24));
    x2[e] = make_float4((float)(v2 & 0xFF), (float)((v2 >> 8) & 0xFF), (float)((v2 >> 16) & 0xFF), (float)(v2 >> 24));
}

__global__ void k_replay_add(int64_t cap, int64_t pos0, int64_t n, int hw, const uint8_t *s, const uint8_t *a,
                             const double *r, const uint8_t *s2, const uint8_t *term, uint8_t *r_s, uint8_t *r_s2,
                             uint8_t *r_a, float *r_r, uint8_t *r_term)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t slot = (pos0 + i) % cap;
    for (int c = 0; c < hw; c++) { r_s[slot * hw + c] = s[i * hw + c]; r_s2[slot * hw + c] = s2[i * hw + c]; }
    r_a[slot] = a[i] & 3;
    r_r[slot] = (float)r[i];
    r_term[slot] = term[i] ? 1 : 0;
}

// expected_Qs = discount * max_a target_Q(s') [terminals -> 0] + rewards (value.py:120-122)
__global__ void k_td_target(const float *qt, int n_actions, const float *r, const uint8_t *term, float discount, float *y,
                            int64_t batch)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    float m = qt[b * n_actions];
    for (int a = 1; a < n_actions; a++) m = fmaxf(m, qt[b * n_actions + a]);
    if (term[b]) m = 0.f;
    y[b] = discount * m + r[b];
}

// loss and dL/dQ.  bxb: F.mse_loss(Qs[B,1], y[B]) broadcasts to B x B
// (value.py:119-123): L = mean_ij (q_i - y_j)^2, dL/dq_i = (2/B)(q_i - mean(y)).
// Otherwise the per-sample TD loss.  Three stages, all with fixed chunking and
// order (deterministic): per-block partial sums, one-block fold -> scalars[0] =
// loss, scalars[3] = mean(y); elementwise dQ.
static const int LOSS_BLOCKS = 256;
__global__ void __launch_bounds__(256) k_loss_partial(const float *q, const uint8_t *a, const float *y, int n_actions,
                                                      int64_t batch, double *partial)
{
    __shared__ double sh[5][256];
    const int64_t chunk = (batch + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(batch, lo + chunk);
    double v[5] = {0, 0, 0, 0, 0};
    for (int64_t b = lo + threadIdx.x; b < hi; b += blockDim.x) {
        const double qv = q[b * n_actions + a[b]], yv = y[b];
        v[0] += qv; v[1] += qv * qv; v[2] += yv; v[3] += yv * yv; v[4] += (qv - yv) * (qv - yv);
    }
    for (int k = 0; k < 5; k++) sh[k][threadIdx.x] = v[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) for (int k = 0; k < 5; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 5) partial[blockIdx.x * 5 + threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void __launch_bounds__(256) k_loss_final(const double *partial, int n_partials, int64_t batch, int bxb, float *scalars)
{
    __shared__ double sh[5][256];
    for (int k = 0; k < 5; k++) sh[k][threadIdx.x] = (int)threadIdx.x < n_partials ? partial[threadIdx.x * 5 + k] : 0.0;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) for (int k = 0; k < 5; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double B = (double)batch;
        const double mq = sh[0][0] / B, mq2 = sh[1][0] / B, my = sh[2][0] / B, my2 = sh[3][0] / B, md2 = sh[4][0] / B;
        scalars[0] = (float)(bxb ? (mq2 - 2.0 * mq * my + my2) : md2);
        scalars[3] = (float)my;
    }
}

__global__ void k_loss_dq(const float *q, const uint8_t *a, const float *y, int n_actions, int64_t batch, int bxb,
                          const float *scalars, float *dq)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const float qv = q[b * n_actions + a[b]];
    const float tgt = bxb ? scalars[3] : y[b];
    for (int k = 0; k < n_actions; k++) dq[b * n_actions + k] = 0.f;
    dq[b * n_actions + a[b]] = (float)(2.0 / (double)batch) * (qv - tgt);
}

// total gradient norm (clip_grad_norm_, value.py:128): scalars[1] = norm,
// scalars[2] = clip coefficient min(1, 10 / (norm + 1e-6)).
__global__ void __launch_bounds__(1024) k_grad_norm(const float *g, int64_t n, float max_norm, float *scalars)
{
    __shared__ double sh[1024];
    double s = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)g[i] * g[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 512; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float norm = (float)sqrt(sh[0]);
        scalars[1] = norm;
        scalars[2] = fminf(1.f, max_norm / (norm + 1e-6f));
    }
}

// torch.optim.Adam(amsgrad=True) single-tensor update, defaults betas
// (0.9, 0.999), eps 1e-8, no weight decay (value.py:87).
__global__ void k_adam_amsgrad(float *p, const float *g, float *m, float *v, float *vmax, int64_t n, const float *scalars,
                               float lr, float bc1, float bc2_sqrt, const StepVars *sv, const int *cursor)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (sv) { bc1 = sv[*cursor].bc1; bc2_sqrt = sv[*cursor].bc2_sqrt; }
    const float grad = g[i] * scalars[2];
    const float mi = m[i] + (grad - m[i]) * (float)(1.0 - 0.9);    // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * 0.999f + (float)(1.0 - 0.999) * grad * grad;
    const float vm = fmaxf(vmax[i], vi);
    m[i] = mi; v[i] = vi; vmax[i] = vm;
    const float denom = sqrtf(vm) / bc2_sqrt + 1e-8f;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
}

// nn.Linear default init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias
__global__ void k_init_params(float *p, int64_t off, int64_t n, float bound, uint64_t seed)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[4];
    const uint64_t c = (uint64_t)(off + i);
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), 0x9e3779b9u, 0x51u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    const float u = (float)(w[0] >> 8) * (1.0f / 16777216.0f);
    p[off + i] = (2.f * u - 1.f) * bound;
}

// One lock-step of acting for every environment: epsilon-greedy on the Q row
// (value.py:94-111: argmax w.p. 1-eps+eps/A, else uniform == "u < eps ->
// uniform action"), env.step, transition appended to the replay ring, reset
// when done (train.py:62-70).  random_policy: dqn_warmup (warmup.py:8-23).
struct DqnStepArgs {
    Level level;
    EnvArrays arr;
    int64_t n, env_id0;
    uint64_t seed, step;
    const uint32_t *words;
    int64_t wpe;
    int *status;
    const float *q;
    unsigned long long thr;
    int random_policy;
    int cheat;              // learn.py:39-47: store the hidden reward and the action really taken
    const StepVars *sv; const int *cursor;     // graph replay: step, thr and pos0 come from sv[*cursor]
    int64_t cap, pos0;
    uint8_t *r_s, *r_s2, *r_a, *r_term;
    float *r_r;
};

template <int KIND, class Rng>
__global__ void __launch_bounds__(SGK_BLOCK) k_dqn_step(const __grid_constant__ DqnStepArgs p)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const Level &L = p.level;
    EnvRegs e;
    unpack_env<KIND>(p.arr.core[i], e);
    e.ep_return = p.arr.ep_return[i];
    e.hidden_cum = p.arr.hidden_cum[i];
    uint64_t step = p.step;
    unsigned long long thr = p.thr;
    int64_t pos0 = p.pos0;
    if (p.sv) { const StepVars v = p.sv[*p.cursor]; step = v.step; thr = v.thr; pos0 = v.pos0; }
    Rng rng;
    RngInit<Rng>::load(rng, p.seed, p.env_id0 + i, p.words, p.wpe, p.arr.replay_cursor, i);
    rng.set_step(step);
    int a;
    if (p.random_policy) {
        a = rng.random_action();
    } else {
        const float4 q = *reinterpret_cast<const float4 *>(p.q + i * 4);
        a = 0; float m = q.x;
        if (q.y > m) { m = q.y; a = 1; }
        if (q.z > m) { m = q.z; a = 2; }
        if (q.w > m) { m = q.w; a = 3; }
        if (rng.agent_uniform() < thr) a = rng.agent_choice();
    }
    const int64_t slot = (pos0 + i) % p.cap;
    uint8_t *s = p.r_s + slot * L.HW, *s2 = p.r_s2 + slot * L.HW;
    for (int c = 0; c < L.HW; c++) s[c] = render_cell<KIND>(L, e, c);
    const StepOut o = env_step<KIND>(L, e, a, rng);
    for (int c = 0; c < L.HW; c++) s2[c] = render_cell<KIND>(L, e, c);
    p.r_a[slot] = (uint8_t)(p.cheat ? o.actual : a);
    p.r_r[slot] = (float)(p.cheat ? (o.hidden_none ? 0.0 : o.hidden) : o.reward);
    p.r_term[slot] = o.done ? 1 : 0;
    if (o.done) {
        EpStats st;
        st.load(p.arr, i);
        st.episode_end(e, p.level.perf_is_return != 0);
        st.store(p.arr, i);
        rng.set_step(step + 1);
        env_reset<KIND>(L, e, rng);
    }
    RngInit<Rng>::store(rng, p.arr.replay_cursor, i);
    if (rng.overflowed()) *p.status = SGK_ST_REPLAY_DRY;
    p.arr.core[i] = pack_core(e);
    p.arr.ep_return[i] = e.ep_return;
    p.arr.hidden_cum[i] = e.hidden_cum;
}

__global__ void k_advance_cursor(int *cursor) { *cursor += 1; }

#define SGK_DQN_GRAPH_MIN_STEPS 16

template <int KIND>
__global__ void __launch_bounds__(SGK_BLOCK) k_dqn_render_f32(const __grid_constant__ Level L, const uint64_t *core, int64_t n, float *x)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvRegs e;
    unpack_env<KIND>(core[i], e);
    for (int c = 0; c < L.HW; c++) x[i * L.HW + c] = (float)render_cell<KIND>(L, e, c);
}

template <int KIND>
__global__ void __launch_bounds__(SGK_BLOCK) k_dqn_render_u8(const __grid_constant__ Level L, const uint64_t *core, int64_t n, uint8_t *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvRegs e;
    unpack_env<KIND>(core[i], e);
    for (int c = 0; c < L.HW; c++) out[i * L.HW + c] = render_cell<KIND>(L, e, c);
}

// ===================================================================== host helpers
static int gemm(int mode, int M, int N, int K, const float *A, int lda, const float *B, int ldb, float *C, int ldc,
                const float *bias, int relu, const float *mask, int ldmask, int splits, cudaStream_t st)
{
    const long big_blocks = (long)((N + GEMM_BN - 1) / GEMM_BN) * ((M + GEMM_BM - 1) / GEMM_BM) * splits;
    if (big_blocks >= 120) {
        const int k_chunk = ((K + splits - 1) / splits + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
        const dim3 grid((N + GEMM_BN - 1) / GEMM_BN, (M + GEMM_BM - 1) / GEMM_BM, splits);
        if (mode == 0) k_gemm<0><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, relu, mask, ldmask, k_chunk);
        else if (mode == 1) k_gemm<1><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, relu, mask, ldmask, k_chunk);
        else k_gemm<2><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, relu, mask, ldmask, k_chunk);
    } else {
        const int k_chunk = ((K + splits - 1) / splits + 15) / 16 * 16;
        const dim3 grid((N + 63) / 64, (M + 63) / 64, splits);
        if (mode == 0) k_gemm_small<0><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, relu, mask, ldmask, k_chunk);
        else if (mode == 1) k_gemm_small<1><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, relu, mask, ldmask, k_chunk);
        else k_gemm_small<2><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, relu, mask, ldmask, k_chunk);
    }
    return launch_check("k_gemm");
}

static int ensure_rows(sgk_dqn *d, int64_t rows)
{
    if (d->rows_cap >= rows) return SGK_OK;
    auto re = [&](float **p, size_t n) -> int {
        if (*p) cudaFree(*p);
        *p = nullptr;
        CU(cudaMalloc(p, n * sizeof(float)));
        return SGK_OK;
    };
    int rc;
    int widest = 0;
    for (int l = 0; l <= d->n_linear; l++) widest = d->dims[l] > widest ? d->dims[l] : widest;
    if ((rc = re(&d->x, (size_t)rows * d->dims[0]))) return rc;
    if ((rc = re(&d->x2, (size_t)rows * d->dims[0]))) return rc;
    for (int l = 0; l < d->n_linear; l++) {
        if ((rc = re(&d->act[l], (size_t)rows * d->dims[l + 1]))) return rc;
        if ((rc = re(&d->act_t[l], (size_t)rows * d->dims[l + 1]))) return rc;
    }
    for (int k = 0; k < 3; k++) if ((rc = re(&d->dact[k], (size_t)rows * widest))) return rc;
    if ((rc = re(&d->y, (size_t)rows))) return rc;
    if ((rc = re(&d->b_r, (size_t)rows))) return rc;
    if (d->b_a) cudaFree(d->b_a);
    if (d->b_term) cudaFree(d->b_term);
    if (d->b_idx) cudaFree(d->b_idx);
    CU(cudaMalloc(&d->b_a, (size_t)rows));
    CU(cudaMalloc(&d->b_term, (size_t)rows));
    CU(cudaMalloc(&d->b_idx, (size_t)rows * 8));
    if (d->xb) cudaFree(d->xb);
    if (d->xb2) cudaFree(d->xb2);
    CU(cudaMalloc(&d->xb, (size_t)rows * d->dims[0]));
    CU(cudaMalloc(&d->xb2, (size_t)rows * d->dims[0]));
    d->rows_cap = rows;
    return SGK_OK;
}

// forward through network `which` on x[rows, n_in] -> acts[l]; returns Q in acts[n_linear-1]
static int forward(sgk_dqn *d, int which, const float *x, int64_t rows, float *const *acts, cudaStream_t st)
{
    const float *in = x;
    for (int l = 0; l < d->n_linear; l++) {
        const int K = d->dims[l], N = d->dims[l + 1];
        const float *W = d->params[which] + d->w_off[l], *b = d->params[which] + d->b_off[l];
        int rc = gemm(0, (int)rows, N, K, in, K, W, K, acts[l], N, b, l + 1 < d->n_linear, nullptr, 0, 1, st);
        if (rc != SGK_OK) return rc;
        in = acts[l];
    }
    return SGK_OK;
}

static bool tc_supported(const sgk_dqn *d)
{
    return d->n_linear == 3 && d->dims[1] == d->dims[2] && d->dims[1] <= 100 && d->dims[0] <= tc::MAX_K_IN &&
           d->n_actions <= 8;
}

// (re)pack network `which` into the operand images if its parameters changed
static int ensure_packed(sgk_dqn *d, int which, cudaStream_t st)
{
    if (!d->w_image[which]) {
        CU(cudaMalloc(&d->w_image[which], 2 * tc::FWD_IMAGE_BYTES));        // [hi | lo] images
        CU(cudaMemsetAsync(d->w_image[which], 0, 2 * tc::FWD_IMAGE_BYTES, st));
        d->w_image_dirty[which] = 1;
    }
    if (which == 0 && !d->w_image_bwd) {
        CU(cudaMalloc(&d->w_image_bwd, tc::BWD_IMAGE_BYTES));
        d->w_image_dirty[0] = 1;
    }
    if (!d->w_image_dirty[which]) return SGK_OK;
    const float *P = d->params[which];
    tc::k_pack_weights<<<dim3(which == 0 ? 8 : 6, 6), 256, 0, st>>>(P + d->w_off[0], P + d->w_off[1], P + d->w_off[2], d->dims[0], d->dims[1],
                                                         d->n_actions, d->w_image[which], which == 0 ? d->w_image_bwd : nullptr);
    d->w_image_dirty[which] = 0;
    return launch_check("k_pack_weights");
}

// the fused tensor-core forward: boards (uint8) -> Q, optionally H1 / H2 in fp32
static int forward_tc(sgk_dqn *d, int which, const uint8_t *boards, int64_t rows, float *q_out, float *h1, float *h2,
                      cudaStream_t st, bool want_images = false)
{
    static bool attr_set = false;
    if (!attr_set) {
        CU(cudaFuncSetAttribute(tc::k_mlp_forward_ts<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemTs::TOTAL));
        CU(cudaFuncSetAttribute(tc::k_mlp_forward_ts<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemTs::TOTAL));
        attr_set = true;
    }
    int prc = ensure_packed(d, which, st);
    if (prc != SGK_OK) return prc;
    tc::Params p;
    p.w_image = d->w_image[which];
    const float *P = d->params[which];
    p.w1 = P + d->w_off[0]; p.b1 = P + d->b_off[0];
    p.w2 = P + d->w_off[1]; p.b2 = P + d->b_off[1];
    p.w3 = P + d->w_off[2]; p.b3 = P + d->b_off[2];
    p.n_in = d->dims[0]; p.n_hidden = d->dims[1]; p.n_out = d->n_actions;
    p.boards = boards; p.rows = rows; p.q_out = q_out; p.h1_out = h1; p.h2_out = h2;
    const int64_t tiles = (rows + tc::TILE_M - 1) / tc::TILE_M;
    p.h1_img = p.h2_img = nullptr;
    if (want_images) {
        if (d->h_img_tiles < tiles) {
            for (int k = 0; k < 2; k++) { if (d->h_img[k]) cudaFree(d->h_img[k]); d->h_img[k] = nullptr; }
            d->h_img_tiles = 0;
            for (int k = 0; k < 2; k++) CU(cudaMalloc(&d->h_img[k], (size_t)tiles * tc::H_IMG_TILE_BYTES));
            d->h_img_tiles = tiles;
        }
        p.h1_img = d->h_img[0]; p.h2_img = d->h_img[1];
    }
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, d->sm_count);
    if (d->use_tc >= 3) tc::k_mlp_forward_ts<true><<<grid, tc::TS_THREADS, tc::SmemTs::TOTAL, st>>>(p);
    else tc::k_mlp_forward_ts<false><<<grid, tc::TS_THREADS, tc::SmemTs::TOTAL, st>>>(p);
    return launch_check("k_mlp_forward_ts");
}

// backward pass fused into one kernel (+ one folding the per-CTA partials): sgk_mlp_tc.cuh k_mlp_backward_fused
static int backward_fused(sgk_dqn *d, int64_t B, const float *dq, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        CU(cudaFuncSetAttribute(tc::k_mlp_backward_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemFb::TOTAL));
        attr_set = true;
    }
    int prc = ensure_packed(d, 0, st);
    if (prc != SGK_OK) return prc;
    const int H = d->dims[1], A = d->n_actions, n_in = d->dims[0];
    const int64_t tiles = (B + tc::TILE_M - 1) / tc::TILE_M;
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, d->sm_count);
    tc::FusedBwdParams p;
    p.w_image = d->w_image_bwd; p.dq = dq; p.h1_img = d->h_img[0]; p.h2_img = d->h_img[1]; p.boards = d->xb;
    p.n_in = n_in; p.n_hidden = H; p.n_out = A; p.n1pad = (n_in + 1 + 15) / 16 * 16; p.rows = B;
    p.scale_up = 0.5f * (float)B;          // the loss put 2/B into dQ: brings the FP16 copies into range
    const int stride = tc::N_OUT + tc::N_HID + p.n1pad;
    const int64_t need = (int64_t)grid * tc::TILE_M * stride;
    if (d->fb_partial_cap < need) {
        if (d->fb_partial) cudaFree(d->fb_partial);
        d->fb_partial = nullptr; d->fb_partial_cap = 0;
        CU(cudaMalloc(&d->fb_partial, (size_t)need * 4));
        d->fb_partial_cap = need;
    }
    p.partial = d->fb_partial;
    tc::k_mlp_backward_fused<<<grid, tc::FB_THREADS, tc::SmemFb::TOTAL, st>>>(p);
    int rc = launch_check("k_mlp_backward_fused");
    if (rc != SGK_OK) return rc;
    tc::FusedFinish f;
    f.partial = d->fb_partial; f.n_partials = (int)grid; f.n_in = n_in; f.n_hidden = H; f.n_out = A; f.n1pad = p.n1pad;
    f.dW1 = d->grads + d->w_off[0]; f.db1 = d->grads + d->b_off[0];
    f.dW2 = d->grads + d->w_off[1]; f.db2 = d->grads + d->b_off[1];
    f.dW3 = d->grads + d->w_off[2]; f.db3 = d->grads + d->b_off[2];
    tc::k_bwd_fused_finish<<<grid_for(d->n_params, tc::WGF_ELEMS), tc::WGF_ELEMS * tc::WGF_LANES, 0, st>>>(f);
    return launch_check("k_bwd_fused_finish");
}

// the unfused tensor-core backward (SGK_BWD_UNFUSED=1, kept for A/B measurements): error chain, then the
// three weight / bias gradients as sample-reductions (sgk_mlp_tc.cuh)
static int backward_tc(sgk_dqn *d, int64_t B, const float *dq, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        CU(cudaFuncSetAttribute(tc::k_mlp_backward_data_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemBwd::TOTAL));
        CU(cudaFuncSetAttribute(tc::k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemWg::TOTAL));
        CU(cudaFuncSetAttribute(tc::k_wgrad_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemWm::TOTAL));
        attr_set = true;
    }
    const int H = d->dims[1], A = d->n_actions, n_in = d->dims[0];
    const int64_t tiles = (B + tc::TILE_M - 1) / tc::TILE_M;
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, d->sm_count);
    float *dh2 = d->dact[1], *dh1 = d->dact[2];
    const float *P0 = d->params[0];
    tc::BwdParams bp;
    bp.w2 = P0 + d->w_off[1]; bp.w3 = P0 + d->w_off[2]; bp.n_hidden = H; bp.n_out = A;
    bp.dq = dq; bp.h1 = d->act[0]; bp.h2 = d->act[1]; bp.dh1 = dh1; bp.dh2 = dh2; bp.rows = B;
    int prc = ensure_packed(d, 0, st);
    if (prc != SGK_OK) return prc;
    bp.w_image = d->w_image_bwd;
    // 107 KB of shared memory and 256 TMEM columns per CTA: two CTAs fit on an SM,
    // so one CTA's epilogue overlaps the other's MMAs
    const unsigned grid_bwd = (unsigned)std::min<int64_t>(tiles, 2 * (int64_t)d->sm_count);
    tc::k_mlp_backward_data_tc<<<grid_bwd, tc::TILE_M, tc::SmemBwd::TOTAL, st>>>(bp);
    int rc = launch_check("k_mlp_backward_data_tc");
    if (rc != SGK_OK) return rc;
    // weight gradients: k_wgrad_mn (128-sample tiles, MN-major operands, no transposition) unless
    // SGK_WGRAD_SCATTER selects the older scatter-transposing kernel (64-sample tiles) for A/B runs
    static const bool scatter = getenv("SGK_WGRAD_SCATTER") != nullptr;
    const int64_t wg_tiles = scatter ? (B + tc::WG_TILE - 1) / tc::WG_TILE : tiles;
    const unsigned wg_grid = (unsigned)std::min<int64_t>(wg_tiles, 3 * (int64_t)d->sm_count);
    const int64_t per_layer = (int64_t)wg_grid * tc::TILE_M * tc::N_HID;
    const int64_t need = 3 * per_layer;
    if (d->partials_cap < need) {
        if (d->partials) cudaFree(d->partials);
        d->partials = nullptr; d->partials_cap = 0;
        CU(cudaMalloc(&d->partials, (size_t)need * 4));
        d->partials_cap = need;
    }
    // the three weight/bias gradients are independent sample-reductions: ONE launch
    // reduces them side by side (blockIdx.y = layer), one more folds the partials
    tc::WgradBatch wb;
    tc::WgradFinishBatch fb;
    int max_total = 0;
    auto set = [&](int y, const float *P, int ldp, int mdim, const float *Q, int ldq, int ndim, int layer) {
        tc::WgradParams &wp = wb.layer[y];
        wp.P = P; wp.ldp = ldp; wp.mdim = mdim; wp.Q = Q; wp.ldq = ldq; wp.ndim = ndim;
        wp.npad = (ndim + 1 + 15) / 16 * 16; wp.add_ones = 1; wp.rows = B; wp.partial = d->partials + (size_t)y * per_layer;
        wp.scale_up = 0.5f * (float)B;      // the loss put 2/B into dQ: brings the error signals into FP16's range
        tc::WgradFinish &f = fb.layer[y];
        f.partial = wp.partial; f.npad = wp.npad; f.mdim = mdim; f.ndim = ndim;
        f.dW = d->grads + d->w_off[layer]; f.db = d->grads + d->b_off[layer];
        max_total = std::max(max_total, mdim * (ndim + 1));
    };
    set(0, dq, A, A, d->act[1], H, H, 2);           // dW3, db3
    set(1, dh2, H, H, d->act[0], H, H, 1);          // dW2, db2
    set(2, dh1, H, H, d->x, n_in, n_in, 0);         // dW1, db1
    if (wg_tiles <= 3 * (int64_t)d->sm_count) {
        // small batches: one layer alone cannot fill the GPU (batch 4,096 = 64 tiles)
        if (scatter) tc::k_wgrad_tc<<<dim3(wg_grid, 3), tc::TILE_M, tc::SmemWg::TOTAL, st>>>(wb);
        else tc::k_wgrad_mn<<<dim3(wg_grid, 3), tc::WM_THREADS, tc::SmemWm::TOTAL, st>>>(wb);
        tc::k_wgrad_finish<<<dim3((max_total + tc::WGF_ELEMS - 1) / tc::WGF_ELEMS, 3), tc::WGF_ELEMS * tc::WGF_LANES, 0, st>>>(fb, (int)wg_grid);
    } else {
        // large batches: every layer fills the GPU by itself; side by side they only
        // compete for L2 (measured 802 -> 1007 us per lock-step at batch 262,144)
        for (int y = 0; y < 3; y++) {
            tc::WgradBatch one_w; one_w.layer[0] = wb.layer[y]; one_w.layer[1] = one_w.layer[2] = wb.layer[y];
            tc::WgradFinishBatch one_f; one_f.layer[0] = fb.layer[y]; one_f.layer[1] = one_f.layer[2] = fb.layer[y];
            const int total = fb.layer[y].mdim * (fb.layer[y].ndim + 1);
            if (scatter) tc::k_wgrad_tc<<<dim3(wg_grid, 1), tc::TILE_M, tc::SmemWg::TOTAL, st>>>(one_w);
            else tc::k_wgrad_mn<<<dim3(wg_grid, 1), tc::WM_THREADS, tc::SmemWm::TOTAL, st>>>(one_w);
            tc::k_wgrad_finish<<<dim3((total + tc::WGF_ELEMS - 1) / tc::WGF_ELEMS, 1), tc::WGF_ELEMS * tc::WGF_LANES, 0, st>>>(one_f, (int)wg_grid);
        }
    }
    return launch_check("k_wgrad_tc");
}

// ===================================================================== C ABI
extern "C" int sgk_dqn_destroy(sgk_dqn *d)
{
    if (!d) return SGK_OK;
    DeviceGuard g(d->device);
    void *ptrs[] = {d->params[0], d->params[1], d->grads, d->adam_m, d->adam_v, d->adam_vmax, d->r_s, d->r_s2, d->r_a,
                    d->r_term, d->r_r, d->x, d->x2, d->dact[0], d->dact[1], d->dact[2], d->y, d->scalars, d->b_a, d->b_term, d->b_r,
                    d->b_idx, d->partials, d->q_env, d->boards_env, d->thr, d->status, d->xb, d->xb2, d->loss_partial};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (int l = 0; l < DQN_MAX_LAYERS; l++) { if (d->act[l]) cudaFree(d->act[l]); if (d->act_t[l]) cudaFree(d->act_t[l]); }
    for (uint8_t *img : {d->w_image[0], d->w_image[1], d->w_image_bwd, d->h_img[0], d->h_img[1]}) if (img) cudaFree(img);
    if (d->fb_partial) cudaFree(d->fb_partial);
    if (d->sv_table) cudaFree(d->sv_table);
    if (d->sv_cursor) cudaFree(d->sv_cursor);
    if (d->cap_stream) cudaStreamDestroy(d->cap_stream);
    delete d;
    return SGK_OK;
}

extern "C" int sgk_dqn_create(const sgk_env *env, int n_layers, int n_hidden, int64_t replay_capacity, int64_t batch_size,
                              uint64_t seed, sgk_dqn **out)
{
    REQUIRE(env != nullptr && out != nullptr, "bad argument");
    REQUIRE(n_layers >= 1 && n_layers + 1 <= DQN_MAX_LAYERS, "n_layers out of range");
    REQUIRE(n_hidden >= 1 && n_hidden <= 4096, "n_hidden out of range");
    REQUIRE(replay_capacity >= 1 && batch_size >= 1, "replay_capacity and batch_size must be positive");
    DeviceGuard g(env->device);
    sgk_dqn *d = new (std::nothrow) sgk_dqn();
    REQUIRE(d != nullptr, "out of host memory");
    memset(d, 0, sizeof(*d));
    d->device = env->device; d->kind = env->level.kind; d->hw = env->level.HW; d->n_actions = SGK_NA;
    d->n_linear = n_layers + 1;
    // n_input: the reference multiplies only two of the three (C,H,W) dims
    // (value.py:66-67, a shape bug); the board has C*H*W cells
    d->dims[0] = d->hw;
    for (int l = 1; l < d->n_linear; l++) d->dims[l] = n_hidden;
    d->dims[d->n_linear] = d->n_actions;
    int64_t off = 0;
    for (int l = 0; l < d->n_linear; l++) {
        d->w_off[l] = off; off += (int64_t)d->dims[l] * d->dims[l + 1];
        d->b_off[l] = off; off += d->dims[l + 1];
    }
    d->n_params = off;
    d->cap = replay_capacity; d->batch = batch_size; d->seed = seed;
    // the reference's default architecture runs on the tensor cores at fp32 accuracy (3xTF32)
    d->use_tc = tc_supported(d) ? 3 : 0;
    cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, env->device);
    d->lr = 1e-3; d->discount = 0.99; d->epsilon = 0.01; d->anneal = 100000; d->sync_every = 10000; d->bxb_loss = 1;
    const size_t pb = (size_t)d->n_params * sizeof(float);
    bool ok = true;
    for (int k = 0; k < 2; k++) ok = ok && cudaMalloc(&d->params[k], pb) == cudaSuccess;
    float **zs[] = {&d->grads, &d->adam_m, &d->adam_v, &d->adam_vmax};
    for (float **z : zs) ok = ok && cudaMalloc(z, pb) == cudaSuccess && cudaMemset(*z, 0, pb) == cudaSuccess;
    const size_t rb = (size_t)d->cap * d->hw;
    ok = ok && cudaMalloc(&d->r_s, rb) == cudaSuccess && cudaMalloc(&d->r_s2, rb) == cudaSuccess &&
         cudaMalloc(&d->r_a, (size_t)d->cap) == cudaSuccess && cudaMalloc(&d->r_term, (size_t)d->cap) == cudaSuccess &&
         cudaMalloc(&d->r_r, (size_t)d->cap * 4) == cudaSuccess && cudaMalloc(&d->scalars, 8 * sizeof(float)) == cudaSuccess &&
         cudaMalloc(&d->status, sizeof(int)) == cudaSuccess && cudaMemset(d->status, 0, sizeof(int)) == cudaSuccess &&
         cudaMalloc(&d->loss_partial, 256 * 5 * sizeof(double)) == cudaSuccess;
    if (!ok) { sgk_dqn_destroy(d); return fail(SGK_ECUDA, "cudaMalloc failed for the deep-Q agent"); }
    for (int l = 0; l < d->n_linear; l++) {
        const float bound = 1.0f / sqrtf((float)d->dims[l]);
        const int64_t nw = (int64_t)d->dims[l] * d->dims[l + 1], nb = d->dims[l + 1];
        k_init_params<<<grid_for(nw, 256), 256>>>(d->params[0], d->w_off[l], nw, bound, seed);
        k_init_params<<<grid_for(nb, 256), 256>>>(d->params[0], d->b_off[l], nb, bound, seed);
        // target_Q is an independently initialised network (value.py:83)
        k_init_params<<<grid_for(nw, 256), 256>>>(d->params[1], d->w_off[l], nw, bound, seed ^ 0xabcdef12345ull);
        k_init_params<<<grid_for(nb, 256), 256>>>(d->params[1], d->b_off[l], nb, bound, seed ^ 0xabcdef12345ull);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { sgk_dqn_destroy(d); return fail(SGK_ECUDA, "parameter init failed"); }
    *out = d;
    return SGK_OK;
}

extern "C" int sgk_dqn_configure(sgk_dqn *d, double lr, double discount, double epsilon, int64_t epsilon_anneal,
                                 int64_t sync_every, int bxb_loss)
{
    REQUIRE(d != nullptr, "d is NULL");
    REQUIRE(epsilon_anneal >= 1 && sync_every >= 1, "epsilon_anneal and sync_every must be >= 1");
    d->lr = lr; d->discount = discount; d->epsilon = epsilon; d->anneal = epsilon_anneal; d->sync_every = sync_every;
    d->bxb_loss = bxb_loss ? 1 : 0;
    return SGK_OK;
}

extern "C" int64_t sgk_dqn_param_count(const sgk_dqn *d) { return d ? d->n_params : 0; }
extern "C" int64_t sgk_dqn_replay_count(const sgk_dqn *d) { return d ? (d->count < d->cap ? d->count : d->cap) : 0; }

extern "C" int sgk_dqn_get_params(const sgk_dqn *d, int which, float *out, void *stream)
{
    REQUIRE(d != nullptr && out != nullptr && (which == 0 || which == 1), "bad argument");
    DeviceGuard g(d->device);
    CU(cudaMemcpyAsync(out, d->params[which], (size_t)d->n_params * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SGK_OK;
}

extern "C" int sgk_dqn_set_params(sgk_dqn *d, int which, const float *in, void *stream)
{
    REQUIRE(d != nullptr && in != nullptr && (which == 0 || which == 1), "bad argument");
    DeviceGuard g(d->device);
    CU(cudaMemcpyAsync(d->params[which], in, (size_t)d->n_params * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    d->w_image_dirty[which] = 1;
    return SGK_OK;
}

extern "C" int sgk_dqn_get_grads(const sgk_dqn *d, float *out, void *stream)
{
    REQUIRE(d != nullptr && out != nullptr, "bad argument");
    DeviceGuard g(d->device);
    CU(cudaMemcpyAsync(out, d->grads, (size_t)d->n_params * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SGK_OK;
}

extern "C" int sgk_dqn_sync_target(sgk_dqn *d, void *stream)
{
    REQUIRE(d != nullptr, "d is NULL");
    DeviceGuard g(d->device);
    CU(cudaMemcpyAsync(d->params[1], d->params[0], (size_t)d->n_params * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    d->w_image_dirty[1] = 1;
    // re-pack right away: the target image must never be stale inside a captured lock-step
    if (d->use_tc) return ensure_packed(d, 1, (cudaStream_t)stream);
    return SGK_OK;
}

extern "C" int sgk_dqn_qvalues(sgk_dqn *d, int which, const uint8_t *boards, int64_t n, float *q_out, void *stream)
{
    REQUIRE(d != nullptr && boards && q_out && n > 0 && (which == 0 || which == 1), "bad argument");
    DeviceGuard g(d->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_rows(d, n);
    if (rc != SGK_OK) return rc;
    if (d->use_tc) return forward_tc(d, which, boards, n, q_out, nullptr, nullptr, st);
    k_u8_to_f32<<<(unsigned)std::min<int64_t>((n * d->hw + 255) / 256, 148 * 16), 256, 0, st>>>(boards, d->x, n * d->hw);
    rc = forward(d, which, d->x, n, d->act, st);
    if (rc != SGK_OK) return rc;
    CU(cudaMemcpyAsync(q_out, d->act[d->n_linear - 1], (size_t)n * d->n_actions * 4, cudaMemcpyDeviceToDevice, st));
    return SGK_OK;
}

extern "C" int sgk_dqn_replay_add(sgk_dqn *d, const uint8_t *s, const uint8_t *a, const double *r, const uint8_t *s2,
                                  const uint8_t *term, int64_t n, void *stream)
{
    REQUIRE(d != nullptr && s && a && r && s2 && term && n > 0, "bad argument");
    REQUIRE(n <= d->cap, "more transitions than the ring holds");
    DeviceGuard g(d->device);
    k_replay_add<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(d->cap, d->count % d->cap, n, d->hw, s, a, r, s2, term,
                                                                     d->r_s, d->r_s2, d->r_a, d->r_r, d->r_term);
    d->count += n;
    return launch_check("k_replay_add");
}

extern "C" int sgk_dqn_replay_get(const sgk_dqn *d, int64_t first, int64_t n, uint8_t *s, uint8_t *a, float *r,
                                  uint8_t *s2, uint8_t *term, void *stream)
{
    REQUIRE(d != nullptr && first >= 0 && n > 0 && first + n <= d->cap, "rows outside the ring");
    DeviceGuard g(d->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hw = (size_t)d->hw;
    if (s) CU(cudaMemcpyAsync(s, d->r_s + (size_t)first * hw, (size_t)n * hw, cudaMemcpyDeviceToDevice, st));
    if (s2) CU(cudaMemcpyAsync(s2, d->r_s2 + (size_t)first * hw, (size_t)n * hw, cudaMemcpyDeviceToDevice, st));
    if (a) CU(cudaMemcpyAsync(a, d->r_a + first, (size_t)n, cudaMemcpyDeviceToDevice, st));
    if (r) CU(cudaMemcpyAsync(r, d->r_r + first, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (term) CU(cudaMemcpyAsync(term, d->r_term + first, (size_t)n, cudaMemcpyDeviceToDevice, st));
    return SGK_OK;
}

// loss, backward, clip, Adam on the batch already staged in x / x2 / b_a / b_r / b_term
static int learn_staged(sgk_dqn *d, int64_t B, float *loss_out, cudaStream_t st)
{
    const int L = d->n_linear, A = d->n_actions;
    int rc;
    static const bool unfused = getenv("SGK_BWD_UNFUSED") != nullptr;
    if (d->use_tc) {
        // tensor-core forwards; H1 / H2 come back as FP16 operand images for the fused backward
        // (in fp32 for the unfused A/B path)
        if (unfused) rc = forward_tc(d, 0, d->xb, B, d->act[2], d->act[0], d->act[1], st);
        else rc = forward_tc(d, 0, d->xb, B, d->act[2], nullptr, nullptr, st, true);
        if (rc) return rc;
        if ((rc = forward_tc(d, 1, d->xb2, B, d->act_t[2], nullptr, nullptr, st))) return rc;
    } else {
        if ((rc = forward(d, 0, d->x, B, d->act, st))) return rc;        // Qs = Q(states)
        if ((rc = forward(d, 1, d->x2, B, d->act_t, st))) return rc;     // target_Q(successors)
    }
    k_td_target<<<grid_for(B, 256), 256, 0, st>>>(d->act_t[L - 1], A, d->b_r, d->b_term, (float)d->discount, d->y, B);
    float *dcur = d->dact[0], *dnext = d->dact[1];
    k_loss_partial<<<LOSS_BLOCKS, 256, 0, st>>>(d->act[L - 1], d->b_a, d->y, A, B, d->loss_partial);
    k_loss_final<<<1, 256, 0, st>>>(d->loss_partial, LOSS_BLOCKS, B, d->bxb_loss, d->scalars);
    k_loss_dq<<<grid_for(B, 256), 256, 0, st>>>(d->act[L - 1], d->b_a, d->y, A, B, d->bxb_loss, d->scalars, dcur);
    if ((rc = launch_check("k_loss"))) return rc;
    if (d->use_tc) {
        if ((rc = unfused ? backward_tc(d, B, dcur, st) : backward_fused(d, B, dcur, st))) return rc;
    } else {
        // partial buffer for the split-K weight / bias gradients
        int64_t need = 0;
        for (int l = 0; l < L; l++)
            need = std::max<int64_t>(need, std::max<int64_t>((int64_t)SPLITS * d->dims[l] * d->dims[l + 1], (int64_t)COL_SPLITS * d->dims[l + 1]));
        if (d->partials_cap < need) {
            if (d->partials) cudaFree(d->partials);
            d->partials = nullptr; d->partials_cap = 0;
            CU(cudaMalloc(&d->partials, (size_t)need * 4));
            d->partials_cap = need;
        }
        for (int l = L - 1; l >= 0; l--) {
            const int K = d->dims[l], N = d->dims[l + 1];
            const float *in = l == 0 ? d->x : d->act[l - 1];
            // dW[N,K] = dY^T[N,B] * in[B,K], split over the batch
            const int w_splits = (int)std::min<int64_t>(SPLITS, (B + 63) / 64);
            if ((rc = gemm(2, N, K, (int)B, dcur, N, in, K, d->partials, K, nullptr, 0, nullptr, 0, w_splits, st))) return rc;
            k_reduce_splits<<<grid_for((int64_t)N * K, 256), 256, 0, st>>>(d->partials, d->grads + d->w_off[l], (int64_t)N * K, w_splits);
            const int col_splits = (int)std::min<int64_t>(COL_SPLITS, (B + 31) / 32);
            const int chunk = (int)((B + col_splits - 1) / col_splits);
            k_colsum_partial<<<dim3((N + 127) / 128, col_splits), 128, 0, st>>>(dcur, (int)B, N, N, d->partials, chunk);
            k_reduce_splits<<<grid_for(N, 256), 256, 0, st>>>(d->partials, d->grads + d->b_off[l], N, col_splits);
            if (l > 0) {
                // dX[B,K] = dY[B,N] * W[N,K], masked by relu'(in)
                const float *W = d->params[0] + d->w_off[l];
                if ((rc = gemm(1, (int)B, K, N, dcur, N, W, K, dnext, K, nullptr, 0, in, K, 1, st))) return rc;
                float *t = dcur; dcur = dnext; dnext = t;
            }
        }
    }
    k_grad_norm<<<1, 1024, 0, st>>>(d->grads, d->n_params, 10.f, d->scalars);
    d->adam_step += 1;
    const double bc1 = 1.0 - pow(0.9, (double)d->adam_step), bc2 = 1.0 - pow(0.999, (double)d->adam_step);
    k_adam_amsgrad<<<grid_for(d->n_params, 256), 256, 0, st>>>(d->params[0], d->grads, d->adam_m, d->adam_v, d->adam_vmax,
                                                               d->n_params, d->scalars, (float)d->lr, (float)bc1, (float)sqrt(bc2),
                                                               d->sv_use, d->sv_cursor);
    d->w_image_dirty[0] = 1;
    if ((rc = launch_check("k_adam_amsgrad"))) return rc;
    if (loss_out) CU(cudaMemcpyAsync(loss_out, d->scalars, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SGK_OK;
}

extern "C" int sgk_dqn_learn_batch(sgk_dqn *d, const uint8_t *s, const uint8_t *a, const double *r, const uint8_t *s2,
                                   const uint8_t *term, int64_t n, float *loss_out, void *stream)
{
    REQUIRE(d != nullptr && s && a && r && s2 && term && n > 0, "bad argument");
    DeviceGuard g(d->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_rows(d, n);
    if (rc != SGK_OK) return rc;
    k_u8_to_f32<<<(unsigned)std::min<int64_t>((n * d->hw + 255) / 256, 148 * 16), 256, 0, st>>>(s, d->x, n * d->hw);
    k_u8_to_f32<<<(unsigned)std::min<int64_t>((n * d->hw + 255) / 256, 148 * 16), 256, 0, st>>>(s2, d->x2, n * d->hw);
    CU(cudaMemcpyAsync(d->xb, s, (size_t)n * d->hw, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(d->xb2, s2, (size_t)n * d->hw, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(d->b_a, a, (size_t)n, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(d->b_term, term, (size_t)n, cudaMemcpyDeviceToDevice, st));
    // rewards arrive as float64 (the env's dtype); the network works in float32
    k_f64_to_f32<<<grid_for(n, 256), 256, 0, st>>>(r, d->b_r, n);
    return learn_staged(d, n, loss_out, st);
}

extern "C" int sgk_dqn_learn(sgk_dqn *d, uint64_t step, float *loss_out, void *stream)
{
    REQUIRE(d != nullptr, "d is NULL");
    REQUIRE(d->count > 0, "the replay ring is empty");
    DeviceGuard g(d->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t B = d->batch, fill = d->count < d->cap ? d->count : d->cap;
    int rc = ensure_rows(d, B);
    if (rc != SGK_OK) return rc;
    k_replay_sample<<<grid_for(B, 128), 128, 0, st>>>(d->seed, step, fill, B, d->hw, d->r_s, d->r_s2, d->r_a, d->r_r, d->r_term,
                                                      d->x, d->x2, d->xb, d->xb2, d->b_a, d->b_r, d->b_term, d->b_idx, (d->hw & 3) != 0,
                                                      d->sv_use, d->sv_cursor);
    if ((d->hw & 3) == 0) {
        const int words = d->hw / 4;
        k_replay_gather_words<<<grid_for(B * words, 256), 256, 0, st>>>(
            d->b_idx, B, words, reinterpret_cast<const uint32_t *>(d->r_s), reinterpret_cast<const uint32_t *>(d->r_s2),
            reinterpret_cast<uint32_t *>(d->xb), reinterpret_cast<uint32_t *>(d->xb2), reinterpret_cast<float4 *>(d->x),
            reinterpret_cast<float4 *>(d->x2));
    }
    if ((rc = launch_check("k_replay_sample"))) return rc;
    return learn_staged(d, B, loss_out, st);
}

extern "C" int sgk_rollout_dqn(sgk_env *env, sgk_dqn *d, int64_t n_steps, uint64_t t0, int mode, void *stream)
{
    REQUIRE(env != nullptr && d != nullptr && n_steps > 0, "bad argument");
    REQUIRE((mode & ~(SGK_DQN_LEARN | SGK_DQN_CHEAT)) == 0, "unknown mode bits");
    const bool learn = (mode & SGK_DQN_LEARN) != 0;
    REQUIRE(env->device == d->device && env->level.kind == d->kind, "agent was created for a different environment kind");
    REQUIRE(env->n <= d->cap, "replay capacity is smaller than one lock-step of transitions");
    DeviceGuard g(env->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_rows(d, std::max<int64_t>(env->n, d->batch));
    if (rc != SGK_OK) return rc;
    if (d->env_rows < env->n) {
        if (d->q_env) cudaFree(d->q_env);
        d->q_env = nullptr; d->env_rows = 0;
        CU(cudaMalloc(&d->q_env, (size_t)env->n * d->n_actions * 4));
        d->env_rows = env->n;
    }
    const bool replay = env->rng_mode == SGK_RNG_REPLAY;
    // epsilon of agent-step t: DeepQAgent keeps entry 0 (value.py:72-76)
    auto threshold_at = [&](uint64_t t) -> unsigned long long {
        const int64_t last = d->anneal > 1 ? d->anneal - 1 : 0;
        const int64_t idx = (int64_t)t < last ? (int64_t)t : last;
        const volatile double scaled = (1 - d->epsilon) * (double)idx;
        const volatile double frac = scaled / (double)d->anneal;
        const double eps = 1.0 - frac;
        return eps <= 0.0 ? 0ull : (unsigned long long)ceil(eps * 9007199254740992.0);
    };
    // one lock-step enqueued on `st`; host bookkeeping (ring count, Adam step) advances with it
    auto lockstep = [&](uint64_t t, cudaStream_t st) -> int {
        int rc;
        DqnStepArgs a;
        a.level = env->level; a.arr = env->arr; a.n = env->n; a.env_id0 = env->env_id0; a.seed = env->seed; a.step = t;
        a.words = env->replay_words; a.wpe = env->words_per_env; a.status = env->status;
        a.random_policy = learn ? 0 : 1;
        a.cheat = (mode & SGK_DQN_CHEAT) ? 1 : 0;
        a.sv = d->sv_use; a.cursor = d->sv_cursor;
        a.cap = d->cap; a.pos0 = d->count % d->cap;
        a.r_s = d->r_s; a.r_s2 = d->r_s2; a.r_a = d->r_a; a.r_term = d->r_term; a.r_r = d->r_r;
        a.q = d->q_env; a.thr = 0;
        if (learn) {
            // act: Q(s) for every environment's current board
            const bool use_tc = d->use_tc != 0;
            rc = by_kind(env->level.kind, [&](auto K) {
                constexpr int KIND = decltype(K)::value;
                if (use_tc) k_dqn_render_u8<KIND><<<grid_for(env->n, SGK_BLOCK), SGK_BLOCK, 0, st>>>(env->level, env->arr.core, env->n, d->xb);
                else k_dqn_render_f32<KIND><<<grid_for(env->n, SGK_BLOCK), SGK_BLOCK, 0, st>>>(env->level, env->arr.core, env->n, d->x);
                return launch_check("k_dqn_render");
            });
            if (rc != SGK_OK) return rc;
            if (use_tc) {
                if ((rc = forward_tc(d, 0, d->xb, env->n, d->q_env, nullptr, nullptr, st))) return rc;
            } else {
                if ((rc = forward(d, 0, d->x, env->n, d->act, st))) return rc;
                CU(cudaMemcpyAsync(d->q_env, d->act[d->n_linear - 1], (size_t)env->n * d->n_actions * 4, cudaMemcpyDeviceToDevice, st));
            }
            a.thr = threshold_at(t);
        }
        rc = by_kind(env->level.kind, [&](auto K) {
            constexpr int KIND = decltype(K)::value;
            if (replay) k_dqn_step<KIND, ReplayStream><<<grid_for(env->n, SGK_BLOCK), SGK_BLOCK, 0, st>>>(a);
            else k_dqn_step<KIND, PhiloxStream><<<grid_for(env->n, SGK_BLOCK), SGK_BLOCK, 0, st>>>(a);
            return launch_check("k_dqn_step");
        });
        if (rc != SGK_OK) return rc;
        d->count += env->n;
        if (learn && (rc = sgk_dqn_learn(d, t, nullptr, st))) return rc;
        return SGK_OK;
    };
    auto sync_if_due = [&](uint64_t t) -> int {
        if (learn && t % (uint64_t)d->sync_every == (uint64_t)d->sync_every - 1)   // learn.py:55-56
            return sgk_dqn_sync_target(d, stream);
        return SGK_OK;
    };

    // Long learning rollouts replay ONE captured CUDA graph of the lock-step
    // (about 20 kernels) instead of enqueueing every kernel again: the host
    // cost per lock-step drops from ~160 us of launches to one graph launch.
    // Per-step scalars are uploaded once and indexed by a device cursor.
    int64_t k = 0;
    const bool use_graph = learn && n_steps >= SGK_DQN_GRAPH_MIN_STEPS && getenv("SGK_DQN_NO_GRAPH") == nullptr;
    if (use_graph) {
        // first lock-step directly: settles every lazy allocation outside the capture
        if ((rc = lockstep(t0, st))) return rc;
        if ((rc = sync_if_due(t0))) return rc;
        k = 1;
        const int64_t m = n_steps - 1;
        if (d->sv_cap < m) {
            if (d->sv_table) cudaFree(d->sv_table);
            d->sv_table = nullptr; d->sv_cap = 0;
            CU(cudaMalloc(&d->sv_table, (size_t)m * sizeof(StepVars)));
            d->sv_cap = m;
        }
        if (!d->sv_cursor) CU(cudaMalloc(&d->sv_cursor, sizeof(int)));
        if (!d->cap_stream) CU(cudaStreamCreateWithFlags(&d->cap_stream, cudaStreamNonBlocking));
        std::vector<StepVars> host((size_t)m);
        for (int64_t j = 0; j < m; j++) {
            StepVars &v = host[(size_t)j];
            const int64_t count_before = d->count + j * env->n, count_after = count_before + env->n;
            const int64_t adam = d->adam_step + j + 1;
            v.step = t0 + (uint64_t)(j + 1);
            v.thr = threshold_at(v.step);
            v.pos0 = count_before % d->cap;
            v.fill = count_after < d->cap ? count_after : d->cap;
            v.bc1 = (float)(1.0 - pow(0.9, (double)adam));
            v.bc2_sqrt = (float)sqrt(1.0 - pow(0.999, (double)adam));
        }
        CU(cudaMemcpyAsync(d->sv_table, host.data(), (size_t)m * sizeof(StepVars), cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(d->sv_cursor, 0, sizeof(int), st));
        CU(cudaStreamSynchronize(st));            // `host` goes out of scope; the capture below is on another stream
        // capture one lock-step reading sv_table[*sv_cursor]
        const int64_t count0 = d->count, adam0 = d->adam_step;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        d->sv_use = d->sv_table;
        CU(cudaStreamBeginCapture(d->cap_stream, cudaStreamCaptureModeRelaxed));
        rc = lockstep(t0 + 1, d->cap_stream);
        if (rc == SGK_OK) k_advance_cursor<<<1, 1, 0, d->cap_stream>>>(d->sv_cursor);
        const cudaError_t end = cudaStreamEndCapture(d->cap_stream, &graph);
        d->sv_use = nullptr;
        d->count = count0; d->adam_step = adam0;      // the capture enqueued nothing
        if (rc != SGK_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (end != cudaSuccess) return fail(SGK_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(end));
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        if (ie != cudaSuccess) {
            cudaGraphDestroy(graph);
            return fail(SGK_ECUDA, std::string("cudaGraphInstantiate (deep-Q lock-step): ") + cudaGetErrorString(ie));
        }
        for (; k < n_steps; k++) {
            const cudaError_t le = cudaGraphLaunch(exec, st);
            if (le != cudaSuccess) { rc = fail(SGK_ECUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(le)); break; }
            d->count += env->n;
            d->adam_step += 1;
            if ((rc = sync_if_due(t0 + (uint64_t)k))) break;
        }
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
        return rc;
    }
    for (; k < n_steps; k++) {
        const uint64_t t = t0 + (uint64_t)k;
        if ((rc = lockstep(t, st))) return rc;
        if ((rc = sync_if_due(t))) return rc;
    }
    return SGK_OK;
}

extern "C" int sgk_dqn_set_tensor_cores(sgk_dqn *d, int enabled)
{
    REQUIRE(d != nullptr, "d is NULL");
    REQUIRE(!enabled || tc_supported(d), "the tensor-core forward covers n_layers == 2, n_hidden <= 100, boards <= 64 cells");
    REQUIRE(enabled == 0 || enabled == 1 || enabled == 3, "enabled must be 0 (fp32 FFMA), 1 (single-pass TF32) or 3 (3xTF32)");
    d->use_tc = enabled;
    return SGK_OK;
}

extern "C" int sgk_dqn_get_tensor_cores(const sgk_dqn *d) { return d ? d->use_tc : 0; }

extern "C" int sgk_dqn_last_scalars(const sgk_dqn *d, float *out3, void *stream)
{
    REQUIRE(d != nullptr && out3 != nullptr, "bad argument");
    DeviceGuard g(d->device);
    CU(cudaMemcpyAsync(out3, d->scalars, 3 * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SGK_OK;
}
