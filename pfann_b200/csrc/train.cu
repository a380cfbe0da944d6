// train.cu -- first pieces of the training step (BASELINE configs[4]): the contrastive loss of
// train.py:41-52 forward + backward, and the batch side of datautil/specaug.py:13-42 and datautil/noise.py:96-109.
//
// similarity_loss (NT-Xent over a batch y[N][d] of unit fingerprints, rows 2i / 2i+1 are a positive pair):
//     a = y y^T / tau;   L = -1/N sum_i ( a[i, p(i)] - logsumexp_{j != i} a[i, j] ),   p(i) = i ^ 1
//     dL/dy_i = 1/(tau N) sum_{j != i} ( P_i[j] + P_j[i] - 2 [j = p(i)] ) y_j,         P_i[j] = exp(a[i,j] - lse_i)
// N = 640, d = 64 (configs/n640d64.json): 26 MFLOP per Gram matrix.  bf16 tensor-core products are NOT usable here:
// logits are inner products divided by tau = 0.05, so a 4e-3 operand-rounding error becomes 8 % in exp(); the kernels
// below are fp32 CUDA-core code, two launches (row statistics, then gradient), bit-reproducible (fixed summation order).
#include <math.h>

#include "common.cuh"
#include "pfann_b200.h"

using namespace pfann;

namespace {

constexpr int NX_ROWS = 16;      // rows of the Gram matrix per CTA
constexpr int NX_THREADS = 256;
constexpr int NX_TILE = 64;      // columns (rows of y) staged per step

// S[r][j] = <y_{row0 + r}, y_j> / tau for all j, into shared memory.  ys: [NX_ROWS][d + 1] this CTA's rows,
// yt: [NX_TILE][d + 1] staging, S: [NX_ROWS][N]
__device__ void gram_rows(const float *__restrict__ y, int N, int d, int row0, float inv_tau, float *ys, float *yt, float *S) {
    const int tid = threadIdx.x, ld = d + 1;
    for (int i = tid; i < NX_ROWS * d; i += NX_THREADS) {
        const int r = i / d, c = i - r * d;
        ys[r * ld + c] = row0 + r < N ? y[(size_t)(row0 + r) * d + c] : 0.f;
    }
    for (int j0 = 0; j0 < N; j0 += NX_TILE) {
        __syncthreads();
        for (int i = tid; i < NX_TILE * d; i += NX_THREADS) {
            const int r = i / d, c = i - r * d;
            yt[r * ld + c] = j0 + r < N ? y[(size_t)(j0 + r) * d + c] : 0.f;
        }
        __syncthreads();
        // thread -> (column jl, 4 rows): 64 columns x 4 row groups
        const int jl = tid & (NX_TILE - 1), rg = tid / NX_TILE;
        float acc[NX_ROWS / 4];
#pragma unroll
        for (int q = 0; q < NX_ROWS / 4; q++) acc[q] = 0.f;
        for (int c = 0; c < d; c++) {
            const float v = yt[jl * ld + c];
#pragma unroll
            for (int q = 0; q < NX_ROWS / 4; q++) acc[q] = fmaf(ys[(rg * (NX_ROWS / 4) + q) * ld + c], v, acc[q]);
        }
        if (j0 + jl < N) {
#pragma unroll
            for (int q = 0; q < NX_ROWS / 4; q++) S[(size_t)(rg * (NX_ROWS / 4) + q) * N + j0 + jl] = acc[q] * inv_tau;
        }
    }
    __syncthreads();
}

// lse[i] = logsumexp_{j != i} a[i][j];  rowloss[i] = lse[i] - a[i][p(i)]
__global__ void __launch_bounds__(NX_THREADS) ntxent_stats_kernel(const float *y, int N, int d, float inv_tau, float *lse,
                                                                  float *rowloss) {
    extern __shared__ float sm[];
    float *ys = sm, *yt = ys + NX_ROWS * (d + 1), *S = yt + NX_TILE * (d + 1);
    const int row0 = blockIdx.x * NX_ROWS;
    gram_rows(y, N, d, row0, inv_tau, ys, yt, S);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < NX_ROWS; r += NX_THREADS / 32) {
        const int i = row0 + r;
        if (i >= N) continue;
        const float *s = S + (size_t)r * N;
        float m = -INFINITY;
        for (int j = lane; j < N; j += 32)
            if (j != i) m = fmaxf(m, s[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float e = 0.f;
        for (int j = lane; j < N; j += 32)
            if (j != i) e += expf(s[j] - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if (lane == 0) {
            const float l = m + logf(e);
            lse[i] = l;
            rowloss[i] = l - s[i ^ 1];
        }
    }
}

// dy[i] = 1/(tau N) sum_{j != i} (exp(a_ij - lse_i) + exp(a_ij - lse_j) - 2 [j == p(i)]) y_j;  CTA 0 also sums the loss
__global__ void __launch_bounds__(NX_THREADS) ntxent_grad_kernel(const float *y, int N, int d, float inv_tau, const float *lse,
                                                                 const float *rowloss, float *loss, float *dy) {
    extern __shared__ float sm[];
    float *ys = sm, *yt = ys + NX_ROWS * (d + 1), *S = yt + NX_TILE * (d + 1);
    const int row0 = blockIdx.x * NX_ROWS, tid = threadIdx.x;
    if (blockIdx.x == 0 && tid < 32) {   // loss = mean of the row losses, fixed order
        float t = 0.f;
        for (int j = tid; j < N; j += 32) t += rowloss[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (tid == 0) *loss = t / (float)N;
    }
    if (dy == nullptr) return;
    gram_rows(y, N, d, row0, inv_tau, ys, yt, S);
    // S := the weights W
    for (int i = tid; i < NX_ROWS * N; i += NX_THREADS) {
        const int r = i / N, j = i - r * N, gi = row0 + r;
        float w = 0.f;
        if (gi < N && j != gi) {
            const float s = S[i];
            w = expf(s - lse[gi]) + expf(s - lse[j]) - (j == (gi ^ 1) ? 2.f : 0.f);
        }
        S[i] = w;
    }
    __syncthreads();
    // dy rows: thread -> (row r, channels c, c + stride ...); y_j streamed from global (L2-resident: N d 4 bytes)
    const float scale = inv_tau / (float)N;
    for (int o = tid; o < NX_ROWS * d; o += NX_THREADS) {
        const int r = o / d, c = o - r * d;
        if (row0 + r >= N) continue;
        const float *w = S + (size_t)r * N;
        float acc = 0.f;
        for (int j = 0; j < N; j++) acc = fmaf(w[j], __ldg(y + (size_t)j * d + c), acc);
        dy[(size_t)(row0 + r) * d + c] = acc * scale;
    }
}

// x[b][f][t] *= 1 - mask_b, mask_b = union of three rectangles (cutout, frequency band, time band): specaug.py:13-42
__global__ void specaug_kernel(float *x, const int *rects /*[B][8]: f0 f1 t0 t1 | fb0 fb1 | tb0 tb1*/, long long total,
                               int F, int T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int t = (int)(i % T);
    const long long r = i / T;
    const int f = (int)(r % F);
    const int *q = rects + (r / F) * 8;
    const bool m = (f >= q[0] && f < q[1] && t >= q[2] && t < q[3]) || (f >= q[4] && f < q[5]) || (t >= q[6] && t < q[7]);
    if (m) x[i] = 0.f;   // x * (1 - 1)
}

// noise.py:96-109: x_aug = x + ratio * noise, ratio = sqrt(clamp(mean x^2)) / sqrt(clamp(mean n^2)) * 10^(-snr / 20)
__global__ void __launch_bounds__(256) snr_mix_kernel(const float *x, const float *noise, const float *snr_db, float *out,
                                                      int n) {
    __shared__ float red[2][8];
    const long long b = blockIdx.x;
    const float *xb = x + b * n, *nb = noise + b * n;
    float sx = 0.f, sn = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) {
        sx = fmaf(xb[i], xb[i], sx);
        sn = fmaf(nb[i], nb[i], sn);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sn += __shfl_xor_sync(0xffffffffu, sn, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = sx;
        red[1][threadIdx.x >> 5] = sn;
    }
    __syncthreads();
    float tx = 0.f, tn = 0.f;
    for (int w = 0; w < 8; w++) {
        tx += red[0][w];
        tn += red[1][w];
    }
    const float vol_x = sqrtf(fmaxf(tx / (float)n, 1e-12f)), vol_n = sqrtf(fmaxf(tn / (float)n, 1e-12f));
    const float ratio = vol_x / vol_n * powf(10.f, -(snr_db[b] / 20.f));
    for (int i = threadIdx.x; i < n; i += 256) out[b * n + i] = fmaf(ratio, nb[i], xb[i]);
}


// ---- impulse-response convolution (datautil/dataset_v2.py:157-163) ----
// The reference multiplies rfft(x, n) by the spectra of a room and a microphone response and keeps samples
// [pad_start, segment_size) of the inverse transform; n is chosen >= len(x) + len(h1) + len(h2) (dataset_v2.py:51-58), so
// the circular product IS the causal linear convolution.  Here: the direct form, register-tiled FIR --
//     out[b][i] = sum_{k < L} h[b][k] * x[b][out_start + i - k]        (x = 0 outside [0, n))
// CTA = 1024 outputs of one row, thread = 8 consecutive outputs; per chunk of 256 taps the CTA stages the taps and the
// 1280-sample window they touch; per 8 taps a thread reads 16 window samples (4 LDS.128) + 8 taps for 64 FMAs.
constexpr int IR_OUT = 1024, IR_TK = 256;
__global__ void __launch_bounds__(128) ir_conv_kernel(const float *x, int n, const float *h, int L, float *out, int out_start,
                                                      int out_len) {
    __shared__ __align__(16) float h_s[IR_TK];
    __shared__ __align__(16) float x_s[IR_OUT + IR_TK];
    const int b = blockIdx.y, base = blockIdx.x * IR_OUT, tid = threadIdx.x;
    const float *xb = x + (long long)b * n, *hb = h + (long long)b * L;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
    const int g_last = out_start + min(base + IR_OUT, out_len) - 1;   // last output this CTA owns
    const int k_end = min(L, g_last + 1);                             // taps beyond it only see x[< 0]
    for (int k0 = 0; k0 < k_end; k0 += IR_TK) {
        const int ws = out_start + base - k0 - IR_TK;
        __syncthreads();
        for (int j = tid; j < IR_TK; j += 128) h_s[j] = (k0 + j < L) ? __ldg(hb + k0 + j) : 0.f;
        for (int j = tid; j < IR_OUT + IR_TK; j += 128) {
            const int gi = ws + j;
            x_s[j] = (gi >= 0 && gi < n) ? __ldg(xb + gi) : 0.f;
        }
        __syncthreads();
#pragma unroll 2
        for (int kk8 = 0; kk8 < IR_TK; kk8 += 8) {
            float xs[16], hs[8];
            const float4 *xp = reinterpret_cast<const float4 *>(x_s + 8 * tid - kk8 + IR_TK - 8);
            const float4 *hp = reinterpret_cast<const float4 *>(h_s + kk8);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 v = xp[q];
                xs[4 * q] = v.x; xs[4 * q + 1] = v.y; xs[4 * q + 2] = v.z; xs[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const float4 v = hp[q];
                hs[4 * q] = v.x; hs[4 * q + 1] = v.y; hs[4 * q + 2] = v.z; hs[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int e = 0; e < 8; e++)
#pragma unroll
                for (int i = 0; i < 8; i++) acc[i] = fmaf(hs[e], xs[i - e + 8], acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int o = base + 8 * tid + i;
        if (o < out_len) out[(long long)b * out_len + o] = acc[i];
    }
}

}  // namespace

extern "C" {

int pfann_ntxent(pfann_ctx *hctx, const float *y, int N, int d, float tau, float *loss, float *dy) {
    PF_CHECK(hctx && y && loss && N >= 2 && (N % 2) == 0 && d > 0 && tau > 0.f, PFANN_ERR_ARG,
             "pfann_ntxent: need an even batch of at least 2 rows and tau > 0");
    PF_CHECK(is_device_ptr(y) && is_device_ptr(loss) && (!dy || is_device_ptr(dy)), PFANN_ERR_ARG,
             "pfann_ntxent: device pointers only");
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    PF_CUDA(cudaSetDevice(ctx->device));
    const size_t smem = ((size_t)(NX_ROWS + NX_TILE) * (d + 1) + (size_t)NX_ROWS * N) * sizeof(float);
    PF_CHECK(smem <= 200 * 1024, PFANN_ERR_UNSUPPORTED, "pfann_ntxent: batch %d x %d needs %zu B of shared memory", N, d, smem);
    PF_TRY(ctx->stage_out[3].ensure((size_t)2 * N * sizeof(float)));
    float *lse = ctx->stage_out[3].as<float>(), *rowloss = lse + N;
    PF_CUDA(cudaFuncSetAttribute(ntxent_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PF_CUDA(cudaFuncSetAttribute(ntxent_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = cdiv(N, NX_ROWS);
    ProfScope ps(ctx, K_MISC);
    ntxent_stats_kernel<<<grid, NX_THREADS, smem, ctx->stream>>>(y, N, d, 1.f / tau, lse, rowloss);
    ntxent_grad_kernel<<<dy ? grid : 1u, NX_THREADS, smem, ctx->stream>>>(y, N, d, 1.f / tau, lse, rowloss, loss, dy);
    ctx->launches += 2;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

int pfann_specaug_apply(pfann_ctx *hctx, float *x, const int32_t *rects, int64_t B, int F, int T) {
    PF_CHECK(hctx && B >= 0 && F > 0 && T > 0 && (B == 0 || (x && rects)), PFANN_ERR_ARG, "pfann_specaug_apply: bad argument");
    if (B == 0) return PFANN_OK;
    PF_CHECK(is_device_ptr(x) && is_device_ptr(rects), PFANN_ERR_ARG, "pfann_specaug_apply: device pointers only");
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    PF_CUDA(cudaSetDevice(ctx->device));
    const long long total = (long long)B * F * T;
    ProfScope ps(ctx, K_MISC);
    specaug_kernel<<<cdiv(total, 256), 256, 0, ctx->stream>>>(x, rects, total, F, T);
    ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

int pfann_snr_mix(pfann_ctx *hctx, const float *x, const float *noise, const float *snr_db, int64_t B, int n, float *out) {
    PF_CHECK(hctx && B >= 0 && n > 0 && (B == 0 || (x && noise && snr_db && out)), PFANN_ERR_ARG, "pfann_snr_mix: bad argument");
    if (B == 0) return PFANN_OK;
    PF_CHECK(is_device_ptr(x) && is_device_ptr(noise) && is_device_ptr(snr_db) && is_device_ptr(out), PFANN_ERR_ARG,
             "pfann_snr_mix: device pointers only");
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    PF_CUDA(cudaSetDevice(ctx->device));
    ProfScope ps(ctx, K_MISC);
    snr_mix_kernel<<<(unsigned)B, 256, 0, ctx->stream>>>(x, noise, snr_db, out, n);
    ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

int pfann_ir_conv(pfann_ctx *hctx, const float *x, int64_t B, int n, const float *h, int L, float *out, int out_start,
                  int out_len) {
    PF_CHECK(hctx && B >= 0 && B <= 65535 && n > 0 && L > 0 && out_start >= 0 && out_len > 0 && (B == 0 || (x && h && out)),
             PFANN_ERR_ARG, "pfann_ir_conv: bad argument");
    if (B == 0) return PFANN_OK;
    PF_CHECK(is_device_ptr(x) && is_device_ptr(h) && is_device_ptr(out), PFANN_ERR_ARG, "pfann_ir_conv: device pointers only");
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    PF_CUDA(cudaSetDevice(ctx->device));
    ProfScope ps(ctx, K_MISC);
    ir_conv_kernel<<<dim3(cdiv(out_len, IR_OUT), (unsigned)B), 128, 0, ctx->stream>>>(x, n, h, L, out, out_start, out_len);
    ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

}  // extern "C"
