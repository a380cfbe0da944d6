// encoder_train.cu -- training step of the fingerprint network (BASELINE configs[4]): FpNetwork forward
// with the activations kept (model.py:148-153) and its backward, i.e. what torch autograd does for the reference in
// train.py:99-103.  fp32 CUDA-core kernels throughout (the reference trains in fp32 / AMP; gradients feed an optimizer, a
// bf16 tensor-core backward is future work -- DESIGN.md section 8); activations are channels-last like the inference
// path.  Every option set of FpNetwork: ReLU / ELU, relu_after_bn True / False, any strides, dense or depthwise conv2.
//
// Per convolution i (conv -> LayerNorm over (C,F,T) with per-element affine -> ReLU), given dA = dL/d(output):
//     dn   = dA * [A > 0]                                   ReLU'
//     dgamma[e] += dn[b,e] * xhat[b,e],  dbeta[e] += dn[b,e]  xhat = (Y - mean_b) * rstd_b      (sum over the batch)
//     g    = dn * gamma
//     dY   = rstd_b * (g - mean_e(g) - xhat * mean_e(g * xhat))                                   LayerNorm backward
//     dW[(j,c)][o] = sum_m X_in[m @ tap j][c] * dY[m][o],   db[o] = sum_m dY[m][o]                 weight gradient
//     dX_in = transposed convolution of dY with W                                                 data gradient
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "encoder.cuh"
#include "pfann_b200.h"

using namespace pfann;

namespace pfann {
int enc_conv_f32(Model *m, const ConvWeights &cw, const float *X, float *Y, int nb);
int enc_ln_stats_f32(Model *m, const ConvWeights &cw, const float *Y, int nb, float2 *stats);
int enc_ln_apply_f32(Model *m, const ConvWeights &cw, const float *Y, const float2 *stats, float *X, int nb);
int enc_conv_bwd_data_f32(Model *m, const ConvGeom &g, const float *dY, const float *Wt, float *dX, int nb);
}  // namespace pfann

namespace {

struct TrainState {
    int B = 0;
    DevBuf Y[16], X[16], stats[16];        // raw conv outputs, post-activation outputs, (mean, rstd) per sample
    DevBuf lnred;                          // [B] (mean g, mean g xhat) of the LayerNorm backward
    DevBuf dA, dB;                         // gradient ping-pong buffers (largest activation)
    DevBuf hp, hr, hdr;                    // head: pre-ELU [B][d u], pre-normalisation output [B][d], its gradient
    DevBuf gW[16], gB[16], gG[16], gBe[16], gw1, gb1, gw2, gb2;   // parameter gradients (kernel layouts)
    float *wt[16] = {};                    // transposed weights [(tap, o)][c] for the data gradient
    bool wt_valid[16] = {};
    const float *mel = nullptr;
    bool have_grads = false;
};

__device__ __forceinline__ double block_sum_d(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; i++) t += red[i];
    return t;
}

// ---- head (model.py:122-130), training forward: x = activated encoder output [B][h] ----
__global__ void head_train_fwd_kernel(const float *x, const float *w1, const float *b1, const float *w2, const float *b2,
                                      float *p, float *r, float *z, int d, int h, int u, int norm) {
    extern __shared__ float red[];
    const long long b = blockIdx.x;
    const int g = threadIdx.x, v = h / d;
    float out = 0.f;
    if (g < d) {
        out = b2[g];
        for (int j = 0; j < u; j++) {
            float acc = b1[g * u + j];
            for (int i = 0; i < v; i++) acc = fmaf(w1[(size_t)(g * u + j) * v + i], x[b * h + g * v + i], acc);
            p[(b * d + g) * u + j] = acc;
            out = fmaf(w2[g * u + j], acc > 0.f ? acc : expm1f(acc), out);
        }
        r[b * d + g] = out;
    }
    float s = g < d ? out * out : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[i];
    if (g < d) z[b * d + g] = norm ? out / fmaxf(sqrtf(t), 1e-12f) : out;
}

// dr = d(loss)/d(r) from dz: F.normalize backward  dr = (dz - z (z . dz)) / max(|r|, eps)   (identity when !norm)
__global__ void normalize_bwd_kernel(const float *dz, const float *r, float *dr, int d, int norm) {
    extern __shared__ float red[];
    const long long b = blockIdx.x;
    const int g = threadIdx.x;
    const float rv = g < d ? r[b * d + g] : 0.f, gv = g < d ? dz[b * d + g] : 0.f;
    if (!norm) {
        if (g < d) dr[b * d + g] = gv;
        return;
    }
    float s1 = rv * rv, s2 = rv * gv;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const int nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s1; red[nw + (threadIdx.x >> 5)] = s2; }
    __syncthreads();
    float t1 = 0.f, t2 = 0.f;
    for (int i = 0; i < nw; i++) { t1 += red[i]; t2 += red[nw + i]; }
    const float nrm = fmaxf(sqrtf(t1), 1e-12f);
    if (g < d) dr[b * d + g] = (gv - rv * (t2 / (nrm * nrm))) / nrm;
}

// CTA = (output dimension g, slice of the batch), one thread per hidden unit j: sums over the slice in registers, then
// one atomic per parameter element and CTA;
// dx[b][g v + i] = sum_j dp[b][g][j] w1[g][j][i] via a warp reduction
__global__ void head_bwd_kernel(const float *x, const float *p, const float *dr, const float *w1, const float *w2, int B, int d,
                                int h, int u, float *gw1, float *gb1, float *gw2, float *gb2, float *dx) {
    const int g = blockIdx.x, j = threadIdx.x, v = h / d;   // blockDim.x = 32 >= u
    float aw1[16], ab1 = 0.f, aw2 = 0.f, ab2 = 0.f, w1r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        aw1[i] = 0.f;
        w1r[i] = (j < u && i < v) ? w1[(size_t)(g * u + j) * v + i] : 0.f;
    }
    const float w2v = j < u ? w2[g * u + j] : 0.f;
    const int per = (B + gridDim.y - 1) / gridDim.y;
    const int bb0 = blockIdx.y * per, bb1 = (bb0 + per) < B ? (bb0 + per) : B;
    for (int b = bb0; b < bb1; b++) {
        const float drv = dr[(size_t)b * d + g];
        const float pv = j < u ? p[((size_t)b * d + g) * u + j] : 0.f;
        const float e = pv > 0.f ? pv : expm1f(pv);
        const float dp = j < u ? drv * w2v * (pv > 0.f ? 1.f : e + 1.f) : 0.f;   // ELU' = 1 or exp(p)
        aw2 = fmaf(drv, e, aw2);
        ab2 += drv;
        ab1 += dp;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (i < v) {
                const float xv = x[(size_t)b * h + g * v + i];
                aw1[i] = fmaf(dp, xv, aw1[i]);
                float c = dp * w1r[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                if (j == 0) dx[(size_t)b * h + g * v + i] = c;
            }
        }
    }
    if (j < u) {
        for (int i = 0; i < v && i < 16; i++) atomicAdd(gw1 + (size_t)(g * u + j) * v + i, aw1[i]);
        atomicAdd(gb1 + g * u + j, ab1);
        atomicAdd(gw2 + g * u + j, aw2);
    }
    if (j == 0) atomicAdd(gb2 + g, ab2);
}

// ---- LayerNorm + activation backward, all option sets of model.py:58-72 ----
// mode bits 0-1: 0 ReLU, 1 ELU; bit 2: activation BEFORE the LayerNorm (relu_after_bn == False).
//   after  (default): A = act(n), n = xhat gamma + beta, xhat = (Y - mean) rstd:  dn = dA act'(n), dY = LNbwd(dn gamma)
//   before          : A = zhat gamma + beta, zhat = (act(Y) - mean) rstd:          dn = dA, dY = LNbwd(dn gamma) act'(Y)
// act' from the stored output: ReLU' = [out > 0]; ELU' = out > 0 ? 1 : out + 1 (= exp(in) for in <= 0).
__device__ __forceinline__ float act_val(float y, int act) { return act == 1 ? (y > 0.f ? y : expm1f(y)) : fmaxf(y, 0.f); }
__device__ __forceinline__ float act_grad_from_out(float out, int act) {
    return act == 1 ? (out > 0.f ? 1.f : out + 1.f) : (out > 0.f ? 1.f : 0.f);
}
// (dn, normalised input of the LayerNorm) for one element
__device__ __forceinline__ void ln_bwd_terms(float dA, float A, float Y, float2 st, int mode, float &dn, float &xh) {
    if (mode & 4) {
        dn = dA;
        xh = (act_val(Y, mode & 3) - st.x) * st.y;
    } else {
        dn = dA * act_grad_from_out(A, mode & 3);
        xh = (Y - st.x) * st.y;
    }
}

// red[b] = (mean_e g, mean_e g xhat), g = dn gamma
__global__ void __launch_bounds__(512) ln_bwd_reduce_kernel(const float *dA, const float *A, const float *Y, const float2 *stats,
                                                            const float *gamma, long long E, float2 *red, int mode) {
    __shared__ double sh[16];
    const long long b = blockIdx.x;
    const float2 st = stats[b];
    double s1 = 0.0, s2 = 0.0;
    if ((E & 3) == 0) {   // 16-byte loads, fp32 within a group of four, double across groups
        const float4 *a4 = reinterpret_cast<const float4 *>(A + b * E), *d4 = reinterpret_cast<const float4 *>(dA + b * E);
        const float4 *y4 = reinterpret_cast<const float4 *>(Y + b * E), *g4 = reinterpret_cast<const float4 *>(gamma);
        for (long long e = threadIdx.x; e < E / 4; e += blockDim.x) {
            const float4 a = a4[e], d = d4[e], y = y4[e], gm = __ldg(g4 + e);
            float dn[4], xh[4];
            ln_bwd_terms(d.x, a.x, y.x, st, mode, dn[0], xh[0]);
            ln_bwd_terms(d.y, a.y, y.y, st, mode, dn[1], xh[1]);
            ln_bwd_terms(d.z, a.z, y.z, st, mode, dn[2], xh[2]);
            ln_bwd_terms(d.w, a.w, y.w, st, mode, dn[3], xh[3]);
            const float g0 = dn[0] * gm.x, g1 = dn[1] * gm.y, g2 = dn[2] * gm.z, g3 = dn[3] * gm.w;
            s1 += (double)((g0 + g1) + (g2 + g3));
            s2 += (double)(fmaf(g0, xh[0], g1 * xh[1]) + fmaf(g2, xh[2], g3 * xh[3]));
        }
    } else {
        for (long long e = threadIdx.x; e < E; e += blockDim.x) {
            float dn, xh;
            ln_bwd_terms(dA[b * E + e], A[b * E + e], Y[b * E + e], st, mode, dn, xh);
            const float g = dn * gamma[e];
            s1 += (double)g;
            s2 += (double)g * (double)xh;
        }
    }
    const double t1 = block_sum_d(s1, sh), t2 = block_sum_d(s2, sh);
    if (threadIdx.x == 0) red[b] = make_float2((float)(t1 / (double)E), (float)(t2 / (double)E));
}

// dY (in place over dA) and the affine gradients; a thread owns 4 consecutive elements and walks `group` samples
__global__ void __launch_bounds__(256) ln_bwd_apply_kernel(float *dA, const float *A, const float *Y, const float2 *stats,
                                                           const float2 *red, const float *gamma, long long E, int nb, int group,
                                                           float *gG, float *gBe, int mode) {
    const long long e0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (e0 >= E) return;
    const int b0 = blockIdx.y * group, b1 = (b0 + group) < nb ? (b0 + group) : nb;
    const int n = (int)((E - e0) < 4 ? (E - e0) : 4);
    float gam[4], ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < n; i++) gam[i] = gamma[e0 + i];
    const bool vec = (E & 3) == 0;   // n == 4 and 16-byte aligned rows
    for (int b = b0; b < b1; b++) {
        const float2 st = __ldg(stats + b), rd = __ldg(red + b);
        const long long idx = (long long)b * E + e0;
        float av[4], yv[4], dv[4];
        if (vec) {
            const float4 a = *reinterpret_cast<const float4 *>(A + idx), y = *reinterpret_cast<const float4 *>(Y + idx);
            const float4 d = *reinterpret_cast<const float4 *>(dA + idx);
            av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w;
            yv[0] = y.x; yv[1] = y.y; yv[2] = y.z; yv[3] = y.w;
            dv[0] = d.x; dv[1] = d.y; dv[2] = d.z; dv[3] = d.w;
        } else {
            for (int i = 0; i < n; i++) { av[i] = A[idx + i]; yv[i] = Y[idx + i]; dv[i] = dA[idx + i]; }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < n) {
                float dn, xh;
                ln_bwd_terms(dv[i], av[i], yv[i], st, mode, dn, xh);
                ag[i] = fmaf(dn, xh, ag[i]);
                ab[i] += dn;
                float dz = st.y * (dn * gam[i] - rd.x - xh * rd.y);
                // activation before the LayerNorm: its derivative at the conv output (ReLU: Y > 0; ELU: exp(Y) for Y <= 0)
                if (mode & 4) dz *= (mode & 3) == 1 ? (yv[i] > 0.f ? 1.f : expf(yv[i])) : (yv[i] > 0.f ? 1.f : 0.f);
                dv[i] = dz;
            }
        }
        if (vec) *reinterpret_cast<float4 *>(dA + idx) = make_float4(dv[0], dv[1], dv[2], dv[3]);
        else for (int i = 0; i < n; i++) dA[idx + i] = dv[i];
    }
    for (int i = 0; i < n; i++) {
        atomicAdd(gG + e0 + i, ag[i]);
        atomicAdd(gBe + e0 + i, ab[i]);
    }
}

// ---- dense convolution, weight gradient: dW[k][o] += sum_{m in slice} A[m][k] dY[m][o], k = (tap, c) ----
struct BwdWArgs {
    const float *X;     // conv input [nb][Fi][Ti][Ci]
    const float *dY;    // [M][Co]
    float *dW;          // [K][Co]
    float *db;          // [Co]
    long long M, m_per_split;
    int Ci, Co, Fi, Ti, Fo, To, axis, ntaps, K, stride;
    int off[3];
};

__global__ void __launch_bounds__(256) conv_bwd_w_kernel(const BwdWArgs a) {
    __shared__ float As[16][64 + 4];   // [m][k]
    __shared__ float Gs[16][64];       // [m][o]
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const long long m_begin = (long long)blockIdx.z * a.m_per_split;
    long long m_end = m_begin + a.m_per_split;
    if (m_end > a.M) m_end = a.M;
    // loader roles: 16 rows x 16 threads, each 4 consecutive k (A) / 4 consecutive o (G)
    const int lm = tid >> 4, lq = (tid & 15) * 4;
    const int tx = tid & 15, ty = tid >> 4;   // compute: k = ty*4.., o = tx*4..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    float bsum = 0.f;   // bias gradient: k-tile 0 only, thread column o = n0 + tid (tid < 64)
    for (long long m0 = m_begin; m0 < m_end; m0 += 16) {
        const long long m = m0 + lm;
        float av[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < m_end) {
            const int to = (int)(m % a.To);
            const long long r = m / a.To;
            const int fo = (int)(r % a.Fo);
            const long long b = r / a.Fo;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = k0 + lq + e;
                if (k < a.K) {
                    const int tap = k / a.Ci, c = k - tap * a.Ci;
                    int fi = fo, ti = to;
                    if (a.axis == 0) ti = a.stride * to + a.off[tap]; else fi = a.stride * fo + a.off[tap];
                    if (fi >= 0 && fi < a.Fi && ti >= 0 && ti < a.Ti) av[e] = __ldg(a.X + ((b * a.Fi + fi) * a.Ti + ti) * (long long)a.Ci + c);
                }
                const int o = n0 + lq + e;
                if (o < a.Co) gv[e] = __ldg(a.dY + m * a.Co + o);
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; e++) {
            As[lm][lq + e] = av[e];
            Gs[lm][lq + e] = gv[e];
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < 16; mm++) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[mm][ty * 4]);
            const float4 g4 = *reinterpret_cast<const float4 *>(&Gs[mm][tx * 4]);
            const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(aa[i], gg[j], acc[i][j]);
        }
        if (blockIdx.x == 0 && tid < 64) {
#pragma unroll
            for (int mm = 0; mm < 16; mm++) bsum += Gs[mm][tid];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int k = k0 + ty * 4 + i;
        if (k >= a.K) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int o = n0 + tx * 4 + j;
            if (o < a.Co) atomicAdd(a.dW + (size_t)k * a.Co + o, acc[i][j]);
        }
    }
    if (blockIdx.x == 0 && tid < 64 && n0 + tid < a.Co) atomicAdd(a.db + n0 + tid, bsum);
}

// ---- depthwise conv2 (fuller == false): data and weight gradients ----
__global__ void dw_bwd_data_kernel(const float *dY, const float *W /*[C][ntaps]*/, float *dX, long long total, int C, int Fi,
                                   int Ti, int Fo, int ntaps, int off0, int off1, int off2, int stride) {
    // thread = VEC consecutive channels of one input position; total = elements of dX [nb][Fi][Ti][C]
    const int VEC = (C & 3) == 0 ? 4 : 1;
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i >= total) return;
    const int c = (int)(i % C);
    long long r = i / C;
    const int t = (int)(r % Ti);
    r /= Ti;
    const int fi = (int)(r % Fi);
    const long long b = r / Fi;
    const int offs[3] = {off0, off1, off2};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < ntaps; j++) {
        const int num = fi - offs[j];
        if (num < 0 || num % stride) continue;
        const int fo = num / stride;
        if (fo >= Fo) continue;
        const float *g = dY + ((b * Fo + fo) * Ti + t) * (long long)C + c;
        if (VEC == 4) {
            const float4 gv = *reinterpret_cast<const float4 *>(g);
            acc[0] = fmaf(__ldg(W + c * ntaps + j), gv.x, acc[0]);
            acc[1] = fmaf(__ldg(W + (c + 1) * ntaps + j), gv.y, acc[1]);
            acc[2] = fmaf(__ldg(W + (c + 2) * ntaps + j), gv.z, acc[2]);
            acc[3] = fmaf(__ldg(W + (c + 3) * ntaps + j), gv.w, acc[3]);
        } else {
            acc[0] = fmaf(__ldg(W + c * ntaps + j), *g, acc[0]);
        }
    }
    if (VEC == 4) *reinterpret_cast<float4 *>(dX + i) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else dX[i] = acc[0];
}

// CTA = (256 / CW row lanes) x (CW channels) over a slice of output positions: coalesced along the channels, partial
// sums in registers, row lanes folded through shared memory, then C * ntaps atomics per CTA
__global__ void __launch_bounds__(256) dw_bwd_w_kernel(const float *X, const float *dY, long long rows /*nb*Fo*To*/, int C, int CW,
                                                       int Fi, int Ti, int Fo, int To, int ntaps, int off0, int off1, int off2,
                                                       int stride, float *gW, float *gB) {
    __shared__ float sh[4][256];
    const int lane_c = threadIdx.x % CW, lane_r = threadIdx.x / CW, RL = 256 / CW;
    const int c = blockIdx.x * CW + lane_c;
    const int offs[3] = {off0, off1, off2};
    const long long per = (rows + gridDim.y - 1) / gridDim.y;
    const long long r0 = (long long)blockIdx.y * per, r1 = (r0 + per) < rows ? (r0 + per) : rows;
    float aw[3] = {0.f, 0.f, 0.f}, ab = 0.f;
    if (c < C) {
        for (long long m = r0 + lane_r; m < r1; m += RL) {
            const int to = (int)(m % To);
            const long long r = m / To;
            const int fo = (int)(r % Fo);
            const long long b = r / Fo;
            const float g = __ldg(dY + m * C + c);
            ab += g;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                if (j < ntaps) {
                    const int fi = stride * fo + offs[j];
                    if (fi >= 0 && fi < Fi) aw[j] = fmaf(g, __ldg(X + ((b * Fi + fi) * Ti + to) * (long long)C + c), aw[j]);
                }
            }
        }
    }
    sh[0][threadIdx.x] = aw[0]; sh[1][threadIdx.x] = aw[1]; sh[2][threadIdx.x] = aw[2]; sh[3][threadIdx.x] = ab;
    __syncthreads();
    if (lane_r == 0 && c < C) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int q = 0; q < RL; q++)
#pragma unroll
            for (int e = 0; e < 4; e++) t[e] += sh[e][q * CW + lane_c];
        for (int jj = 0; jj < ntaps; jj++) atomicAdd(gW + c * ntaps + jj, t[jj]);
        atomicAdd(gB + c, t[3]);
    }
}


// kernel layouts -> the reference's (PyTorch) element order, for device-side readers of the gradients
// mode 0: dense conv [(j, c)][o] -> [o][c][3];  1: depthwise [o][ntaps] -> [o][1][3];  2: LayerNorm [f][t][o] -> [o][f][t]
__global__ void grad_layout_kernel(const float *src, float *dst, long long total, int mode, int Ci, int Co, int ntaps, int k0,
                                   int k1, int k2, int Fo, int To) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    if (mode == 2) {
        const int t = (int)(i % To);
        const long long r = i / To;
        const int f = (int)(r % Fo), o = (int)(r / Fo);
        dst[i] = src[((long long)f * To + t) * Co + o];
        return;
    }
    const int k = (int)(i % 3);
    const long long r = i / 3;
    const int tk[3] = {k0, k1, k2};
    int j = -1;
    for (int q = 0; q < ntaps; q++)
        if (tk[q] == k) j = q;
    float v = 0.f;   // dead taps only ever multiply padding
    if (j >= 0) {
        if (mode == 1) v = src[r * ntaps + j];
        else {
            const int c = (int)(r % Ci), o = (int)(r / Ci);
            v = src[((long long)j * Ci + c) * Co + o];
        }
    }
    dst[i] = v;
}

// the reference's element order -> kernel layouts (inverse of grad_layout_kernel), for parameters that live on the device
__global__ void param_layout_kernel(const float *src, float *dst, long long total, int mode, int Ci, int Co, int ntaps, int k0,
                                    int k1, int k2, int Fo, int To) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over dst
    if (i >= total) return;
    const int tk[3] = {k0, k1, k2};
    if (mode == 2) {            // [f][t][o] <- [o][f][t]
        const int o = (int)(i % Co);
        const long long r = i / Co;
        dst[i] = src[(long long)o * Fo * To + r];
    } else if (mode == 1) {     // [o][ntaps] <- [o][1][3]
        const int j = (int)(i % ntaps);
        dst[i] = src[(i / ntaps) * 3 + tk[j]];
    } else {                    // [(j, c)][o] <- [o][c][3]
        const int o = (int)(i % Co);
        const long long r = i / Co;
        const int c = (int)(r % Ci), j = (int)(r / Ci);
        dst[i] = src[((long long)o * Ci + c) * 3 + tk[j]];
    }
}

// wt[(j, o)][c] <- w_kn[(j, c)][o]
__global__ void wt_kernel(const float *w_kn, float *wt, long long total, int Ci, int Co) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over wt
    if (i >= total) return;
    const int c = (int)(i % Ci);
    const long long r = i / Ci;
    const int o = (int)(r % Co), j = (int)(r / Co);
    wt[i] = w_kn[((long long)j * Ci + c) * Co + o];
}

TrainState *state(Model *m) {
    if (m->train_state == nullptr) m->train_state = new TrainState();
    return reinterpret_cast<TrainState *>(m->train_state);
}

int zero(Ctx *ctx, DevBuf &b, size_t bytes) {
    PF_TRY(b.ensure(bytes));
    PF_CUDA(cudaMemsetAsync(b.p, 0, bytes, ctx->stream));
    return PFANN_OK;
}

}  // namespace

namespace pfann {
void train_invalidate(Model *m) {
    TrainState *t = reinterpret_cast<TrainState *>(m->train_state);
    if (!t) return;
    for (int i = 0; i < 16; i++) t->wt_valid[i] = false;
    t->B = 0;
    t->have_grads = false;
}
void train_release(Model *m) {
    TrainState *t = reinterpret_cast<TrainState *>(m->train_state);
    if (!t) return;
    for (int i = 0; i < 16; i++) {
        t->Y[i].release(); t->X[i].release(); t->stats[i].release();
        t->gW[i].release(); t->gB[i].release(); t->gG[i].release(); t->gBe[i].release();
        cudaFree(t->wt[i]);
    }
    t->lnred.release(); t->dA.release(); t->dB.release(); t->hp.release(); t->hr.release(); t->hdr.release();
    t->gw1.release(); t->gb1.release(); t->gw2.release(); t->gb2.release();
    delete t;
    m->train_state = nullptr;
}
}  // namespace pfann

extern "C" {

int pfann_model_train_forward(pfann_model *hm, const float *mel, int64_t B, int norm, float *z) {
    PF_CHECK(hm && mel && z && B > 0 && B <= 65535, PFANN_ERR_ARG, "pfann_model_train_forward: bad argument");
    Model *m = reinterpret_cast<Model *>(hm);
    PF_CHECK(m->precision >= 0, PFANN_ERR_STATE, "pfann_model_train_forward: call pfann_model_finalize first");
    PF_CHECK(is_device_ptr(mel) && is_device_ptr(z), PFANN_ERR_ARG, "pfann_model_train_forward: device pointers only");
    PF_CHECK(m->h / m->d <= 16 && m->u <= 32 && m->d <= 1024, PFANN_ERR_UNSUPPORTED,
             "pfann_model_train_forward: head needs h/d <= 16 and u <= 32");
    PF_CUDA(cudaSetDevice(m->ctx->device));
    TrainState *t = state(m);
    t->B = (int)B;
    t->mel = mel;
    t->have_grads = false;
    const float *in = mel;
    for (int i = 0; i < 16; i++) {
        const ConvWeights &cw = m->conv[i];
        const size_t bytes = (size_t)B * cw.g.out_per_sample() * 4;
        PF_TRY(t->Y[i].ensure(bytes));
        PF_TRY(t->X[i].ensure(bytes));
        PF_TRY(t->stats[i].ensure((size_t)B * sizeof(float2)));
        m->prof_idx = i;
        PF_TRY(enc_conv_f32(m, cw, in, t->Y[i].as<float>(), (int)B));
        PF_TRY(enc_ln_stats_f32(m, cw, t->Y[i].as<float>(), (int)B, t->stats[i].as<float2>()));
        PF_TRY(enc_ln_apply_f32(m, cw, t->Y[i].as<float>(), t->stats[i].as<float2>(), t->X[i].as<float>(), (int)B));
        in = t->X[i].as<float>();
    }
    PF_TRY(t->hp.ensure((size_t)B * m->d * m->u * 4));
    PF_TRY(t->hr.ensure((size_t)B * m->d * 4));
    const int threads = ((m->d + 31) / 32) * 32;
    ProfScope ps(m->ctx, K_HEAD, 33);
    head_train_fwd_kernel<<<(unsigned)B, threads, 64 * sizeof(float), m->ctx->stream>>>(
        in, m->w1, m->b1, m->w2, m->b2, t->hp.as<float>(), t->hr.as<float>(), z, m->d, m->h, m->u, norm);
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

int pfann_model_train_backward(pfann_model *hm, const float *dz, int norm) {
    PF_CHECK(hm && dz, PFANN_ERR_ARG, "pfann_model_train_backward: bad argument");
    Model *m = reinterpret_cast<Model *>(hm);
    TrainState *t = reinterpret_cast<TrainState *>(m->train_state);
    PF_CHECK(t && t->B > 0, PFANN_ERR_STATE, "pfann_model_train_backward: no forward to differentiate");
    PF_CHECK(is_device_ptr(dz), PFANN_ERR_ARG, "pfann_model_train_backward: device pointers only");
    PF_CUDA(cudaSetDevice(m->ctx->device));
    Ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int B = t->B, d = m->d, h = m->h, u = m->u, v = h / d;
    // ---- head ----
    PF_TRY(t->hdr.ensure((size_t)B * d * 4));
    PF_TRY(zero(ctx, t->gw1, (size_t)d * u * v * 4));
    PF_TRY(zero(ctx, t->gb1, (size_t)d * u * 4));
    PF_TRY(zero(ctx, t->gw2, (size_t)d * u * 4));
    PF_TRY(zero(ctx, t->gb2, (size_t)d * 4));
    size_t max_act = 0;
    for (int i = 0; i < 16; i++)
        if ((size_t)m->conv[i].g.out_per_sample() > max_act) max_act = (size_t)m->conv[i].g.out_per_sample();
    PF_TRY(t->dA.ensure((size_t)B * max_act * 4));
    PF_TRY(t->dB.ensure((size_t)B * max_act * 4));
    const int threads = ((d + 31) / 32) * 32;
    {
        ProfScope ps(ctx, K_HEAD, 33);
        normalize_bwd_kernel<<<B, threads, 64 * sizeof(float), st>>>(dz, t->hr.as<float>(), t->hdr.as<float>(), d, norm);
        int hb = (4 * ctx->sm_count + d - 1) / d;
        if (hb > B) hb = B;
        head_bwd_kernel<<<dim3(d, hb), 32, 0, st>>>(t->X[15].as<float>(), t->hp.as<float>(), t->hdr.as<float>(), m->w1, m->w2, B, d, h, u,
                                          t->gw1.as<float>(), t->gb1.as<float>(), t->gw2.as<float>(), t->gb2.as<float>(),
                                          t->dA.as<float>());
        ctx->launches += 2;
        PF_CUDA(cudaGetLastError());
    }
    float *dcur = t->dA.as<float>(), *dnext = t->dB.as<float>();   // dcur = gradient w.r.t. X[i]
    const int lnmode = m->act | (m->act_first ? 4 : 0);
    PF_TRY(t->lnred.ensure((size_t)B * sizeof(float2)));
    for (int i = 15; i >= 0; i--) {
        const ConvWeights &cw = m->conv[i];
        const ConvGeom &g = cw.g;
        const long long E = g.out_per_sample();
        m->prof_idx = i;
        PF_TRY(zero(ctx, t->gG[i], (size_t)E * 4));
        PF_TRY(zero(ctx, t->gBe[i], (size_t)E * 4));
        const size_t wbytes = (size_t)(g.depthwise ? g.Co * g.ntaps : g.K() * g.Co) * 4;
        PF_TRY(zero(ctx, t->gW[i], wbytes));
        PF_TRY(zero(ctx, t->gB[i], (size_t)g.Co * 4));
        {   // LayerNorm + ReLU backward: dcur becomes dY in place
            ProfScope ps(ctx, K_LN, 16 + i);
            ln_bwd_reduce_kernel<<<B, 512, 0, st>>>(dcur, t->X[i].as<float>(), t->Y[i].as<float>(), t->stats[i].as<float2>(),
                                                    cw.gamma, E, t->lnred.as<float2>(), lnmode);
            int group = 64;   // samples per thread before its 8 atomics
            while (group > 1 && (long long)cdiv(E, 1024) * cdiv(B, group) < 4LL * ctx->sm_count) group >>= 1;
            dim3 grid(cdiv(E, 1024), cdiv(B, group));
            ln_bwd_apply_kernel<<<grid, 256, 0, st>>>(dcur, t->X[i].as<float>(), t->Y[i].as<float>(), t->stats[i].as<float2>(),
                                                      t->lnred.as<float2>(), cw.gamma, E, B, group, t->gG[i].as<float>(),
                                                      t->gBe[i].as<float>(), lnmode);
            ctx->launches += 2;
            PF_CUDA(cudaGetLastError());
        }
        const float *xin = i == 0 ? t->mel : t->X[i - 1].as<float>();
        const long long rows = (long long)B * g.rows_per_sample();
        ProfScope ps(ctx, K_CONV_CC, i);
        if (g.depthwise) {
            int CW = 32;
            while (CW < 256 && CW < g.Co) CW <<= 1;
            const int cx = (int)cdiv(g.Co, CW);
            long long ys = (8LL * ctx->sm_count + cx - 1) / cx;
            if (ys > (long long)cdiv(rows, 256 / CW)) ys = cdiv(rows, 256 / CW);
            if (ys > 65535) ys = 65535;
            dw_bwd_w_kernel<<<dim3(cx, (unsigned)ys), 256, 0, st>>>(xin, dcur, rows, g.Co, CW, g.Fi, g.Ti, g.Fo, g.To, g.ntaps,
                                                                   g.tap_off[0], g.tap_off[1], g.tap_off[2], g.stride,
                                                                   t->gW[i].as<float>(), t->gB[i].as<float>());
            const long long total = (long long)B * g.Fi * g.Ti * g.Ci;
            dw_bwd_data_kernel<<<cdiv(total, 256 * ((g.Co & 3) == 0 ? 4 : 1)), 256, 0, st>>>(dcur, cw.w_kn, dnext, total, g.Co, g.Fi, g.Ti, g.Fo, g.ntaps,
                                                                 g.tap_off[0], g.tap_off[1], g.tap_off[2], g.stride);
            ctx->launches += 2;
        } else {
            BwdWArgs a;
            a.X = xin; a.dY = dcur; a.dW = t->gW[i].as<float>(); a.db = t->gB[i].as<float>();
            a.M = rows;
            a.Ci = g.Ci; a.Co = g.Co; a.Fi = g.Fi; a.Ti = g.Ti; a.Fo = g.Fo; a.To = g.To; a.axis = g.axis; a.ntaps = g.ntaps;
            a.K = g.K(); a.stride = g.stride;
            for (int j = 0; j < 3; j++) a.off[j] = g.tap_off[j];
            const int kt = (int)cdiv(a.K, 64), nt = (int)cdiv(a.Co, 64);
            int splits = (4 * ctx->sm_count + kt * nt - 1) / (kt * nt);
            if (splits > (int)cdiv(rows, 64)) splits = (int)cdiv(rows, 64);
            if (splits < 1) splits = 1;
            a.m_per_split = ((rows + splits - 1) / splits + 15) / 16 * 16;
            splits = (int)cdiv(rows, a.m_per_split);
            conv_bwd_w_kernel<<<dim3(kt, nt, splits), 256, 0, st>>>(a);
            ctx->launches++;
            if (i > 0) {
                if (!t->wt_valid[i]) {   // [(tap, o)][c] from w_kn [(tap, c)][o], once per parameter update
                    const long long total = (long long)g.K() * g.Co;
                    if (t->wt[i] == nullptr) PF_CUDA(cudaMalloc(&t->wt[i], (size_t)total * 4));
                    wt_kernel<<<cdiv(total, 256), 256, 0, st>>>(cw.w_kn, t->wt[i], total, g.Ci, g.Co);
                    ctx->launches++;
                    t->wt_valid[i] = true;
                }
                PF_TRY(enc_conv_bwd_data_f32(m, g, dcur, t->wt[i], dnext, B));
            }
        }
        PF_CUDA(cudaGetLastError());
        float *tmp = dcur; dcur = dnext; dnext = tmp;
    }
    t->have_grads = true;
    return PFANN_OK;
}

// Parameter refresh for the training loop: `data` (DEVICE, the reference's element order) goes straight into the
// kernel layouts of a finalized fp32 model -- one small kernel per parameter on the context's stream, no host copies.
int pfann_model_train_load_param(pfann_model *hm, const char *name, const float *data, int64_t numel) {
    PF_CHECK(hm && name && data, PFANN_ERR_ARG, "pfann_model_train_load_param: bad argument");
    Model *m = reinterpret_cast<Model *>(hm);
    PF_CHECK(m->precision == PFANN_PRECISION_FP32, PFANN_ERR_STATE,
             "pfann_model_train_load_param: the model must be finalized with PFANN_PRECISION_FP32 first");
    PF_CHECK(is_device_ptr(data), PFANN_ERR_ARG, "pfann_model_train_load_param: device pointers only");
    PF_CUDA(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    TrainState *t = state(m);
    const std::string key(name);
    int l = -1;
    char cn[16] = {0}, pn[16] = {0};
    float *dst = nullptr;
    long long total = 0, expect = 0;
    int mode = -1;
    const ConvGeom *g = nullptr;
    if (sscanf(name, "f.convs.%d.%15[a-z0-9].%15s", &l, cn, pn) == 3 && l >= 0 && l < 8) {
        const std::string c(cn), p(pn);
        const int i = 2 * l + ((c == "conv2" || c == "ln2") ? 1 : 0);
        ConvWeights &cw = m->conv[i];
        g = &cw.g;
        if ((c == "conv1" || c == "conv2") && p == "weight") {
            dst = cw.w_kn; mode = g->depthwise ? 1 : 0;
            total = g->depthwise ? (long long)g->Co * g->ntaps : (long long)g->K() * g->Co;
            expect = (long long)g->Co * (g->depthwise ? 1 : g->Ci) * 3;
            t->wt_valid[i] = false;
        } else if ((c == "conv1" || c == "conv2") && p == "bias") {
            dst = cw.bias; total = expect = g->Co;
        } else if ((c == "ln1" || c == "ln2") && (p == "weight" || p == "bias")) {
            dst = p == "weight" ? cw.gamma : cw.beta; mode = 2; total = expect = g->out_per_sample();
        }
    } else if (key == "g.linear1.weight") { dst = m->w1; total = expect = (long long)m->d * m->u * (m->h / m->d);
    } else if (key == "g.linear1.bias") { dst = m->b1; total = expect = (long long)m->d * m->u;
    } else if (key == "g.linear2.weight") { dst = m->w2; total = expect = (long long)m->d * m->u;
    } else if (key == "g.linear2.bias") { dst = m->b2; total = expect = m->d; }
    PF_CHECK(dst != nullptr, PFANN_ERR_ARG, "pfann_model_train_load_param: unknown parameter '%s'", name);
    PF_CHECK(expect == numel, PFANN_ERR_ARG, "pfann_model_train_load_param: '%s' has %lld elements, got %lld", name, expect,
             (long long)numel);
    if (mode < 0) {
        PF_CUDA(cudaMemcpyAsync(dst, data, (size_t)total * 4, cudaMemcpyDeviceToDevice, st));
    } else {
        param_layout_kernel<<<cdiv(total, 256), 256, 0, st>>>(data, dst, total, mode, g->Ci, g->Co, g->ntaps, g->tap_k[0],
                                                              g->tap_k[1], g->tap_k[2], g->Fo, g->To);
        m->ctx->launches++;
        PF_CUDA(cudaGetLastError());
    }
    t->have_grads = false;
    return PFANN_OK;
}

// gradient of the parameter with the reference's state_dict key `name`, in the reference's (PyTorch) element order
int pfann_model_get_grad(pfann_model *hm, const char *name, float *out, int64_t numel) {
    PF_CHECK(hm && name && out, PFANN_ERR_ARG, "pfann_model_get_grad: bad argument");
    Model *m = reinterpret_cast<Model *>(hm);
    TrainState *t = reinterpret_cast<TrainState *>(m->train_state);
    PF_CHECK(t && t->have_grads, PFANN_ERR_STATE, "pfann_model_get_grad: call pfann_model_train_backward first");
    PF_CUDA(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    auto fetch = [&](const DevBuf &b, size_t n, std::vector<float> &h) -> int {
        h.resize(n);
        PF_CUDA(cudaMemcpyAsync(h.data(), b.p, n * 4, cudaMemcpyDeviceToHost, st));
        PF_CUDA(cudaStreamSynchronize(st));
        return PFANN_OK;
    };
    std::vector<float> hbuf, res;
    const std::string key(name);
    int l = -1;
    char cn[16] = {0}, pn[16] = {0};
    const bool conv_key = sscanf(name, "f.convs.%d.%15[a-z0-9].%15s", &l, cn, pn) == 3 && l >= 0 && l < 8;
    if (is_device_ptr(out)) {   // stays on the device: one layout kernel or one copy on the context's stream
        const DevBuf *src = nullptr;
        long long total = 0;
        int mode = -1;
        const ConvGeom *g = nullptr;
        if (conv_key) {
            const std::string c(cn), p(pn);
            const int i = 2 * l + ((c == "conv2" || c == "ln2") ? 1 : 0);
            g = &m->conv[i].g;
            if ((c == "conv1" || c == "conv2") && p == "weight") {
                src = &t->gW[i]; mode = g->depthwise ? 1 : 0; total = (long long)g->Co * (g->depthwise ? 1 : g->Ci) * 3;
            } else if ((c == "conv1" || c == "conv2") && p == "bias") {
                src = &t->gB[i]; total = g->Co;
            } else if ((c == "ln1" || c == "ln2") && (p == "weight" || p == "bias")) {
                src = p == "weight" ? &t->gG[i] : &t->gBe[i]; mode = 2; total = g->out_per_sample();
            }
        } else if (key == "g.linear1.weight") { src = &t->gw1; total = (long long)m->d * m->u * (m->h / m->d);
        } else if (key == "g.linear1.bias") { src = &t->gb1; total = (long long)m->d * m->u;
        } else if (key == "g.linear2.weight") { src = &t->gw2; total = (long long)m->d * m->u;
        } else if (key == "g.linear2.bias") { src = &t->gb2; total = m->d; }
        PF_CHECK(src != nullptr, PFANN_ERR_ARG, "pfann_model_get_grad: unknown parameter '%s'", name);
        PF_CHECK(total == numel, PFANN_ERR_ARG, "pfann_model_get_grad: '%s' has %lld elements, caller expects %lld", name, total,
                 (long long)numel);
        if (mode < 0) {
            PF_CUDA(cudaMemcpyAsync(out, src->p, (size_t)total * 4, cudaMemcpyDeviceToDevice, st));
        } else {
            grad_layout_kernel<<<cdiv(total, 256), 256, 0, st>>>(src->as<float>(), out, total, mode, g->Ci, g->Co, g->ntaps,
                                                                 g->tap_k[0], g->tap_k[1], g->tap_k[2], g->Fo, g->To);
            m->ctx->launches++;
            PF_CUDA(cudaGetLastError());
        }
        return PFANN_OK;
    }
    if (conv_key) {
        const std::string c(cn), p(pn);
        const int which = (c == "conv2" || c == "ln2") ? 1 : 0, i = 2 * l + which;
        const ConvGeom &g = m->conv[i].g;
        if ((c == "conv1" || c == "conv2") && p == "weight") {
            const int cin = g.depthwise ? 1 : g.Ci;
            res.assign((size_t)g.Co * cin * 3, 0.f);   // dead taps only ever multiply padding: gradient 0
            if (g.depthwise) {
                PF_TRY(fetch(t->gW[i], (size_t)g.Co * g.ntaps, hbuf));
                for (int o = 0; o < g.Co; o++)
                    for (int j = 0; j < g.ntaps; j++) res[(size_t)o * 3 + g.tap_k[j]] = hbuf[(size_t)o * g.ntaps + j];
            } else {
                PF_TRY(fetch(t->gW[i], (size_t)g.K() * g.Co, hbuf));
                for (int o = 0; o < g.Co; o++)
                    for (int j = 0; j < g.ntaps; j++)
                        for (int cc = 0; cc < g.Ci; cc++)
                            res[((size_t)o * g.Ci + cc) * 3 + g.tap_k[j]] = hbuf[((size_t)j * g.Ci + cc) * g.Co + o];
            }
        } else if ((c == "conv1" || c == "conv2") && p == "bias") {
            PF_TRY(fetch(t->gB[i], (size_t)g.Co, res));
        } else if ((c == "ln1" || c == "ln2") && (p == "weight" || p == "bias")) {
            PF_TRY(fetch(p == "weight" ? t->gG[i] : t->gBe[i], (size_t)g.out_per_sample(), hbuf));
            res.resize(hbuf.size());
            for (int o = 0; o < g.Co; o++)
                for (int f = 0; f < g.Fo; f++)
                    for (int tt = 0; tt < g.To; tt++)
                        res[((size_t)o * g.Fo + f) * g.To + tt] = hbuf[((size_t)f * g.To + tt) * g.Co + o];
        }
    } else if (key == "g.linear1.weight") {
        PF_TRY(fetch(t->gw1, (size_t)m->d * m->u * (m->h / m->d), res));
    } else if (key == "g.linear1.bias") {
        PF_TRY(fetch(t->gb1, (size_t)m->d * m->u, res));
    } else if (key == "g.linear2.weight") {
        PF_TRY(fetch(t->gw2, (size_t)m->d * m->u, res));
    } else if (key == "g.linear2.bias") {
        PF_TRY(fetch(t->gb2, (size_t)m->d, res));
    }
    PF_CHECK(!res.empty(), PFANN_ERR_ARG, "pfann_model_get_grad: unknown parameter '%s'", name);
    PF_CHECK((int64_t)res.size() == numel, PFANN_ERR_ARG, "pfann_model_get_grad: '%s' has %zu elements, caller expects %lld", name,
             res.size(), (long long)numel);
    memcpy(out, res.data(), res.size() * 4);
    return PFANN_OK;
}

}  // extern "C"
