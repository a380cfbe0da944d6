// encoder.cu -- stage 2: FpNetwork (model.py:14-153) layer pipeline, LayerNorm / head kernels and the
// fp32 CUDA-core convolution path (PFANN_PRECISION_FP32, validation grade).  The tensor-core path
// (PFANN_PRECISION_BF16) swaps the convolution GEMMs for encoder_tc.cu and keeps everything else.
//
// Per SeparableConv2d (model.py:54-73), activations channels-last X[b][f][t][c]:
//     conv1 (1x3 along T, stride 2)  -> Y (raw, fp32)  -> per-sample LayerNorm statistics over (C,F,T)
//     LN-apply: relu((Y - mean) * rstd * gamma[f][t][c] + beta[f][t][c])      (model.py:59-60)
//     conv2 (3x1 along F, stride 2; dense if fuller else depthwise)  -> Y -> stats -> LN-apply
// and after layer 7 the split head (model.py:122-130) fused with the last LN-apply and the L2 normalise.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "encoder.cuh"
#include "pfann_b200.h"

using namespace pfann;

namespace {

// ------------------------------------------------------------------------------------------------
// fp32 CUDA-core implicit-GEMM convolution: 64x64 tile, BK = 16, 4x4 register micro-tile.
// ------------------------------------------------------------------------------------------------
template <typename InT>
struct ConvArgs {
    const InT *X;       // [nb][Fi][Ti][Ci]
    const float *W;     // [K][Co]
    const float *bias;  // [Co]
    float *Y;           // [nb][Fo][To][Co]
    long long M;        // nb * Fo * To
    int Ci, Co, Fi, Ti, Fo, To, axis, ntaps, K, stride;
    int off[3];
    int trans;          // 1 = backward-data gather: output position p reads source (p - off[j]) / stride when that is integral
};

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float ld_act(const float *p) { return __ldg(p); }
__device__ __forceinline__ float ld_act(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

template <typename InT>
__global__ void __launch_bounds__(256) conv_gemm_fp32_kernel(const ConvArgs<InT> a) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A-load role: row ar (0..63), 4 consecutive k starting at ak
    const int ar = tid >> 2, ak = (tid & 3) * 4;
    const long long am = m0 + ar;
    const InT *rowp[3] = {nullptr, nullptr, nullptr};
    if (am < a.M) {
        const int to = (int)(am % a.To);
        const long long r = am / a.To;
        const int fo = (int)(r % a.Fo);
        const long long b = r / a.Fo;
        for (int j = 0; j < a.ntaps; j++) {
            int fi = fo, ti = to;
            bool ok = true;
            if (a.trans) {   // transposed convolution (training backward, encoder_train.cu): X is the gradient of the conv output
                const int num = (a.axis == 0 ? to : fo) - a.off[j];
                ok = num >= 0 && (num % a.stride) == 0;
                if (a.axis == 0) ti = num / a.stride; else fi = num / a.stride;
            } else {
                if (a.axis == 0) ti = a.stride * to + a.off[j]; else fi = a.stride * fo + a.off[j];
            }
            if (ok && fi >= 0 && fi < a.Fi && ti >= 0 && ti < a.Ti)
                rowp[j] = a.X + ((b * a.Fi + fi) * a.Ti + ti) * (long long)a.Ci;
        }
    }
    const bool vecA = (a.Ci & 3) == 0 && sizeof(InT) == 4;
    // B-load role: k row bk (0..15), 4 consecutive n starting at bn
    const int bk = tid >> 4, bn = (tid & 15) * 4;
    const bool vecB = (a.Co & 3) == 0;

    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.K; k0 += BK) {
        float av[4] = {0.f, 0.f, 0.f, 0.f};
        const int kk = k0 + ak;
        if (vecA) {
            if (kk < a.K) {
                const int tap = kk / a.Ci, c = kk - tap * a.Ci;
                const InT *p = rowp[tap];
                if (p) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(p + c));
                    av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = kk + e;
                if (k < a.K) {
                    const int tap = k / a.Ci, c = k - tap * a.Ci;
                    const InT *p = rowp[tap];
                    if (p) av[e] = ld_act(p + c);
                }
            }
        }
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        const int kb = k0 + bk;
        if (kb < a.K) {
            const float *p = a.W + (long long)kb * a.Co + n0 + bn;
            if (vecB && n0 + bn + 3 < a.Co) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
                bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (n0 + bn + e < a.Co) bv[e] = __ldg(p + e);
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; e++) As[ak + e][ar] = av[e];
#pragma unroll
        for (int e = 0; e < 4; e++) Bs[bk][bn + e] = bv[e];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; k++) {
            const float4 av4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 bv4 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float aa[4] = {av4.x, av4.y, av4.z, av4.w};
            const float bb[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const long long m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n < a.Co) a.Y[m * a.Co + n] = acc[i][j] + (a.bias ? __ldg(a.bias + n) : 0.f);
        }
    }
}

// Depthwise conv2 (fuller == false, model.py:29): a 3-tap FIR per channel along F.
template <typename InT>
__global__ void conv_dw_kernel(const InT *X, const float *W /*[Co][ntaps]*/, const float *bias, float *Y,
                               long long total, int C, int Fi, int Ti, int Fo, int To, int ntaps, int off0, int off1,
                               int off2, int stride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    long long r = i / C;
    const int to = (int)(r % To);
    r /= To;
    const int fo = (int)(r % Fo);
    const long long b = r / Fo;
    const int offs[3] = {off0, off1, off2};
    float acc = __ldg(bias + c);
    for (int j = 0; j < ntaps; j++) {
        const int fi = stride * fo + offs[j];
        if (fi < 0 || fi >= Fi) continue;
        acc = fmaf(__ldg(W + c * ntaps + j), (float)X[((b * Fi + fi) * Ti + to) * (long long)C + c], acc);
    }
    Y[i] = acc;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the whole (C,F,T) volume of one sample (model.py:21,30; eps 1e-5, biased variance)
// ------------------------------------------------------------------------------------------------
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

// conv_activation of model.py:7-12: 0 = ReLU, 1 = ELU(alpha = 1)
__device__ __forceinline__ float act_fn(float v, int act) {
    return act == 1 ? (v > 0.f ? v : expm1f(v)) : fmaxf(v, 0.f);
}
// the two orders of model.py:58-72.  mode bit 0-1 = activation, bit 2 = activation BEFORE the LayerNorm
// (relu_after_bn == False): out = LN(act(y)); otherwise out = act(LN(y)).
__device__ __forceinline__ float ln_act(float y, float2 st, float g, float b, int mode) {
    if (mode & 4) return fmaf((act_fn(y, mode & 3) - st.x) * st.y, g, b);
    return act_fn(fmaf((y - st.x) * st.y, g, b), mode & 3);
}

// two-pass statistics, one CTA per sample: stats[b] = (mean, 1/sqrt(var + eps)); with mode bit 2 the statistics are
// those of act(y)
__global__ void __launch_bounds__(512) ln_stats_kernel(const float *Y, long long E, float2 *stats, int mode) {
    __shared__ double red[16];
    const float *y = Y + (long long)blockIdx.x * E;
    const bool pre = (mode & 4) != 0;
    double s = 0.0;
    for (long long i = threadIdx.x; i < E; i += blockDim.x) s += (double)(pre ? act_fn(y[i], mode & 3) : y[i]);
    const double mean = block_sum_d(s, red) / (double)E;
    double q = 0.0;
    for (long long i = threadIdx.x; i < E; i += blockDim.x) {
        const double dlt = (double)(pre ? act_fn(y[i], mode & 3) : y[i]) - mean;
        q += dlt * dlt;
    }
    const double var = block_sum_d(q, red) / (double)E;
    if (threadIdx.x == 0) stats[blockIdx.x] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
}

__device__ __forceinline__ void store_out(float *p, float v) { *p = v; }
__device__ __forceinline__ void store_out(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// X[b][e] = relu((Y[b][e] - mean_b) * rstd_b * gamma[e] + beta[e]).  A thread owns 4 consecutive elements e and
// walks over the `group` samples of blockIdx.y, so the per-element affine (gamma, beta: 2.27 M parameters for
// default.json, L2-resident) is fetched once per group instead of once per sample.
__device__ __forceinline__ float4 ld_y4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ld_y4(const __nv_bfloat16 *p) {
    const uint2 u = *reinterpret_cast<const uint2 *>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&u.x), b = *reinterpret_cast<const __nv_bfloat162 *>(&u.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ float ld_y1(const float *p) { return *p; }
__device__ __forceinline__ float ld_y1(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

template <typename YT, typename OutT>
__global__ void __launch_bounds__(256) ln_apply_kernel(const YT *Y, const float2 *stats, const float *gamma,
                                                       const float *beta, OutT *X, long long E, int nb, int group,
                                                       int mode) {
    const long long e0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (e0 >= E) return;
    const int b0 = blockIdx.y * group;
    const int b1 = (b0 + group) < nb ? (b0 + group) : nb;
    if (e0 + 3 < E && (E & 3) == 0) {
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + e0));
        const float4 be = __ldg(reinterpret_cast<const float4 *>(beta + e0));
        for (int b = b0; b < b1; b++) {
            const float2 st = __ldg(stats + b);
            const float4 v = ld_y4(Y + (long long)b * E + e0);
            float o0, o1, o2, o3;
            if (mode == 0) {   // default option set: ReLU after the LayerNorm
                o0 = fmaxf(fmaf((v.x - st.x) * st.y, g.x, be.x), 0.f);
                o1 = fmaxf(fmaf((v.y - st.x) * st.y, g.y, be.y), 0.f);
                o2 = fmaxf(fmaf((v.z - st.x) * st.y, g.z, be.z), 0.f);
                o3 = fmaxf(fmaf((v.w - st.x) * st.y, g.w, be.w), 0.f);
            } else {
                o0 = ln_act(v.x, st, g.x, be.x, mode); o1 = ln_act(v.y, st, g.y, be.y, mode);
                o2 = ln_act(v.z, st, g.z, be.z, mode); o3 = ln_act(v.w, st, g.w, be.w, mode);
            }
            OutT *x = X + (long long)b * E + e0;
            if (sizeof(OutT) == 4) {
                *reinterpret_cast<float4 *>(x) = make_float4(o0, o1, o2, o3);
            } else {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t *>(&p0);
                pk.y = *reinterpret_cast<uint32_t *>(&p1);
                *reinterpret_cast<uint2 *>(x) = pk;
            }
        }
    } else {
        for (int b = b0; b < b1; b++) {
            const float2 st = __ldg(stats + b);
            for (int i = 0; i < 4 && e0 + i < E; i++)
                store_out(X + (long long)b * E + e0 + i,
                          ln_act(ld_y1(Y + (long long)b * E + e0 + i), st, __ldg(gamma + e0 + i), __ldg(beta + e0 + i), mode));
        }
    }
}

// bf16 -> bf16 form of ln_apply_kernel for the default option set (the tail of the tensor-core path): 8 elements
// (16 bytes) per thread and four samples in flight per thread, so that the loads of a group do not serialise
__global__ void __launch_bounds__(256) ln_apply_bf16x8_kernel(const __nv_bfloat16 *Y, const float2 *stats, const float *gamma,
                                                              const float *beta, __nv_bfloat16 *X, long long E, int nb,
                                                              int group) {
    const long long e0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (e0 >= E) return;
    const int b0 = blockIdx.y * group;
    const int b1 = (b0 + group) < nb ? (b0 + group) : nb;
    float g[8], be[8];
    {
        const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma + e0)), g1 = __ldg(reinterpret_cast<const float4 *>(gamma + e0 + 4));
        const float4 c0 = __ldg(reinterpret_cast<const float4 *>(beta + e0)), c1 = __ldg(reinterpret_cast<const float4 *>(beta + e0 + 4));
        g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
        be[0] = c0.x; be[1] = c0.y; be[2] = c0.z; be[3] = c0.w; be[4] = c1.x; be[5] = c1.y; be[6] = c1.z; be[7] = c1.w;
    }
    for (int b = b0; b < b1; b += 4) {
        uint4 v[4];
        float2 st[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int bb = b + q < b1 ? b + q : b1 - 1;
            st[q] = __ldg(stats + bb);
            v[q] = *reinterpret_cast<const uint4 *>(Y + (long long)bb * E + e0);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (b + q >= b1) break;
            const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
            uint32_t o[4];
            const float a = st[q].y, c = -st[q].x * st[q].y;   // (y - mean) * rstd = y * a + c
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w[i]));
                const float o0 = fmaxf(fmaf((f.x - st[q].x) * a, g[2 * i], be[2 * i]), 0.f);
                const float o1 = fmaxf(fmaf((f.y - st[q].x) * a, g[2 * i + 1], be[2 * i + 1]), 0.f);
                const __nv_bfloat162 pk = __floats2bfloat162_rn(o0, o1);
                o[i] = *reinterpret_cast<const uint32_t *>(&pk);
            }
            (void)c;
            *reinterpret_cast<uint4 *>(X + (long long)(b + q) * E + e0) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Layer 0, conv1 (C_in = 1, 1x3 along time) fused with ln1 + ReLU (model.py:56-60).
// The conv output (C x F x T/2 = 524 288 values per segment for default.json, the largest activation of
// the network) is never written: its LayerNorm statistics follow exactly from 9 moments of the mel tile,
//   sum  y   = P*sum_c b_c + sum_j A_j S_j                         S_j  = sum_p m_j(p)
//   sum  y^2 = sum_jj' G_jj' R_jj' + 2 sum_j H_j S_j + P*sum_c b_c^2   R_jj' = sum_p m_j(p) m_j'(p)
// (m_j(p) = mel value under tap j at output position p, 0 in the padding; A, G, H = sums over channels of
// w, w w', b w), so one kernel computes conv + normalise + affine + ReLU and stores the bf16 activation.
// ------------------------------------------------------------------------------------------------
struct L0Args {
    const float *mel;   // [nb][F][T]
    int F, T, To, ntaps;
    int off[3];
    int C;
};

__global__ void __launch_bounds__(256) l0_moments_kernel(const L0Args a, const Model::L0Consts k, float2 *stats) {
    const float *m = a.mel + (long long)blockIdx.x * a.F * a.T;
    const int P = a.F * a.To;
    // Same arithmetic as the moments block at the end of mel_kernel (mel.cu), so that the fused extract path and
    // mel -> model give identical statistics: fp32 within a thread and a warp, double across the warps.
    float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // S0 S1 S2 R00 R01 R02 R11 R12 R22
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        const int f = p / a.To, to = p - f * a.To;
        float v[3] = {0.f, 0.f, 0.f};
        for (int j = 0; j < a.ntaps; j++) {
            const int t = 2 * to + a.off[j];
            if (t >= 0 && t < a.T) v[j] = m[f * a.T + t];
        }
        acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2];
        acc[3] = fmaf(v[0], v[0], acc[3]); acc[4] = fmaf(v[0], v[1], acc[4]); acc[5] = fmaf(v[0], v[2], acc[5]);
        acc[6] = fmaf(v[1], v[1], acc[6]); acc[7] = fmaf(v[1], v[2], acc[7]); acc[8] = fmaf(v[2], v[2], acc[8]);
    }
    __shared__ double wred[8][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 9; i++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 9; i++) wred[warp][i] = (double)acc[i];
    }
    __syncthreads();
    double S[3], R[6];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 9; i++) {
            double t = 0.0;
            for (int w = 0; w < 8; w++) t += wred[w][i];
            if (i < 3) S[i] = t; else R[i - 3] = t;
        }
    }
    if (threadIdx.x == 0) {
        const double Rm[3][3] = {{R[0], R[1], R[2]}, {R[1], R[3], R[4]}, {R[2], R[4], R[5]}};
        double s1 = (double)P * k.Bsum, s2 = (double)P * k.B2;
        for (int j = 0; j < 3; j++) {
            s1 += k.A[j] * S[j];
            s2 += 2.0 * k.H[j] * S[j];
            for (int jj = 0; jj < 3; jj++) s2 += k.G[j][jj] * Rm[j][jj];
        }
        const double N = (double)P * a.C;
        const double mean = s1 / N;
        double var = s2 / N - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[blockIdx.x] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
    }
}

// Same statistics from moments the mel kernel already reduced while the tile was in its shared memory (fused
// extract path): one thread per sample.
__global__ void l0_stats_kernel(const double *moments /*[nb][9]*/, const Model::L0Consts k, double P, double C,
                                float2 *stats, int nb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const double *mo = moments + (long long)b * 9;
    const double S[3] = {mo[0], mo[1], mo[2]};
    const double Rm[3][3] = {{mo[3], mo[4], mo[5]}, {mo[4], mo[6], mo[7]}, {mo[5], mo[7], mo[8]}};
    double s1 = P * k.Bsum, s2 = P * k.B2;
    for (int j = 0; j < 3; j++) {
        s1 += k.A[j] * S[j];
        s2 += 2.0 * k.H[j] * S[j];
        for (int jj = 0; jj < 3; jj++) s2 += k.G[j][jj] * Rm[j][jj];
    }
    const double N = P * C;
    const double mean = s1 / N;
    double var = s2 / N - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[b] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
}

// thread = (PP consecutive output positions along time, group of 8 channels); it loops over the samples of its
// group so that the per-position LayerNorm affine (gamma, beta: 4 MB for default.json) and the conv weights stay
// in registers, and the 2*PP+1 mel values a sample contributes are fetched once for all 8*PP outputs.
template <typename OutT, int PP>
__global__ void __launch_bounds__(256) l0_conv_ln_kernel(const L0Args a, const float *__restrict__ w /*[C][ntaps]*/,
                                                         const float *__restrict__ bias, const float *__restrict__ gamma,
                                                         const float *__restrict__ beta, const float2 *__restrict__ stats,
                                                         OutT *__restrict__ X, int nb, int group) {
    const int cgroups = a.C >> 3;
    const int ppb = 256 / cgroups;  // position groups per block
    const int cg = threadIdx.x % cgroups;
    const int pg = blockIdx.x * ppb + threadIdx.x / cgroups;  // position group: PP consecutive `to` of one f
    const int P = a.F * a.To;
    const int p0 = pg * PP;
    if (p0 >= P) return;
    const int f = p0 / a.To, to0 = p0 - f * a.To;
    float wr[8][3], br[8], gr[PP][8], be[PP][8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int ch = cg * 8 + c;
#pragma unroll
        for (int j = 0; j < 3; j++) wr[c][j] = j < a.ntaps ? __ldg(w + ch * a.ntaps + j) : 0.f;
        br[c] = __ldg(bias + ch);
#pragma unroll
        for (int q = 0; q < PP; q++) {
            gr[q][c] = __ldg(gamma + (long long)(p0 + q) * a.C + ch);
            be[q][c] = __ldg(beta + (long long)(p0 + q) * a.C + ch);
        }
    }
    // mel columns needed: t = 2*(to0+q) + off[j]; with offsets {0,1,2} that is the run [2*to0, 2*to0 + 2*PP]
    constexpr int NM = 2 * PP + 1;
    const int tbase = 2 * to0 + a.off[0];
    const int b0 = blockIdx.y * group;
    const int b1 = (b0 + group) < nb ? (b0 + group) : nb;
    // software pipeline over samples: the next sample's mel run and statistics are in flight while this one is
    // being computed
    float nx[NM];
    float2 nst;
    {
        const float *m = a.mel + ((long long)b0 * a.F + f) * a.T;
#pragma unroll
        for (int i = 0; i < NM; i++) {
            const int t = tbase + i;
            nx[i] = (t >= 0 && t < a.T) ? __ldg(m + t) : 0.f;
        }
        nst = __ldg(stats + b0);
    }
    for (int b = b0; b < b1; b++) {
        float mv[NM];
#pragma unroll
        for (int i = 0; i < NM; i++) mv[i] = nx[i];
        const float2 st = nst;
        if (b + 1 < b1) {
            const float *m = a.mel + ((long long)(b + 1) * a.F + f) * a.T;
#pragma unroll
            for (int i = 0; i < NM; i++) {
                const int t = tbase + i;
                nx[i] = (t >= 0 && t < a.T) ? __ldg(m + t) : 0.f;
            }
            nst = __ldg(stats + b + 1);
        }
#pragma unroll
        for (int q = 0; q < PP; q++) {
            float o[8];
#pragma unroll
            for (int c = 0; c < 8; c++) {
                float v = br[c];
                v = fmaf(wr[c][0], mv[2 * q], v);
                v = fmaf(wr[c][1], mv[2 * q + 1], v);
                v = fmaf(wr[c][2], mv[2 * q + 2], v);
                o[c] = fmaxf(fmaf((v - st.x) * st.y, gr[q][c], be[q][c]), 0.f);
            }
            OutT *dst = X + ((long long)b * P + p0 + q) * a.C + cg * 8;
            if (sizeof(OutT) == 2) {
                __nv_bfloat162 q0 = __floats2bfloat162_rn(o[0], o[1]), q1 = __floats2bfloat162_rn(o[2], o[3]);
                __nv_bfloat162 q2 = __floats2bfloat162_rn(o[4], o[5]), q3 = __floats2bfloat162_rn(o[6], o[7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t *>(&q0); pk.y = *reinterpret_cast<uint32_t *>(&q1);
                pk.z = *reinterpret_cast<uint32_t *>(&q2); pk.w = *reinterpret_cast<uint32_t *>(&q3);
                *reinterpret_cast<uint4 *>(dst) = pk;
            } else {
                float *d32 = reinterpret_cast<float *>(dst);
                *reinterpret_cast<float4 *>(d32) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4 *>(d32 + 4) = make_float4(o[4], o[5], o[6], o[7]);
            }
        }
    }
}

// fp32 -> bf16 cast of the mel input for the tensor-core path is not needed (layer 0 conv1 has C_in = 1
// and runs on CUDA cores in both paths).

// channels-last [b][f][t][c] -> NCHW fp32 [b][c][f][t] (debug / parity taps only)
template <typename InT>
__global__ void to_nchw_kernel(const InT *X, float *out, long long total, int C, int F, int T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int t = (int)(i % T);
    long long r = i / T;
    const int f = (int)(r % F);
    r /= F;
    const int c = (int)(r % C);
    const long long b = r / C;
    out[i] = (float)X[((b * F + f) * T + t) * (long long)C + c];
}

// ------------------------------------------------------------------------------------------------
// Head: last LN-apply + ReLU, grouped h -> d*u, ELU, grouped d*u -> d, optional L2 normalise
// (model.py:122-130).  Persistent CTAs (one thread per output dimension g); the head weights (132 KB for
// default.json) are staged ONCE per CTA into shared memory in a lane-major layout ([j][i][g], conflict-free)
// and reused for every sample the CTA processes, instead of being re-read from L2 per sample.
// ------------------------------------------------------------------------------------------------
// HG groups of dp threads per CTA; every group handles HS samples per iteration so that one shared-memory read of a
// weight feeds HS FMAs, and the loads of the next iteration's inputs are independent of the current arithmetic.
constexpr int HEAD_HG = 4, HEAD_HS = 4;

// ELU(alpha = 1) with exp through MUFU.EX2: absolute error ~1e-7 (expm1f costs ~20 instructions per call and was
// two thirds of the head kernel's instruction stream; the head adds 32 such terms with |w2| < 1)
__device__ __forceinline__ float elu_fast(float x) { return x > 0.f ? x : exp2f(x * 1.4426950408889634f) - 1.f; }

template <int V>
__global__ void __launch_bounds__(128 * HEAD_HG) head_kernel(const float *Y /*[nb][h]*/, const float2 *stats, const float *gamma,
                                                   const float *beta, const float *w1, const float *b1, const float *w2,
                                                   const float *b2, float *z, int nb, int d, int h, int u, int norm) {
    extern __shared__ float hsm[];
    const int dp = blockDim.x / HEAD_HG;    // d rounded up to a warp multiple
    float *w1s = hsm;                       // [u][V][dp]
    float *b1s = w1s + (size_t)u * V * dp;  // [u][dp]
    float *w2s = b1s + (size_t)u * dp;      // [u][dp]
    float *red = w2s + (size_t)u * dp;      // [HG][HS][dp / 32]
    const int grp = threadIdx.x / dp, g = threadIdx.x - grp * dp;
    const int nw = dp >> 5;
    // staging: thread (grp, g) copies rows j = grp, grp + HG, ... of output dimension g: V contiguous floats each
    for (int j = grp; j < u; j += HEAD_HG) {
        float wv[V];
#pragma unroll
        for (int i = 0; i < V; i++) wv[i] = 0.f;
        if (g < d) {
            const float4 *src = reinterpret_cast<const float4 *>(w1 + ((size_t)g * u + j) * V);
#pragma unroll
            for (int i = 0; i < V / 4; i++) {
                const float4 t = __ldg(src + i);
                wv[4 * i] = t.x; wv[4 * i + 1] = t.y; wv[4 * i + 2] = t.z; wv[4 * i + 3] = t.w;
            }
        }
#pragma unroll
        for (int i = 0; i < V; i++) w1s[(j * V + i) * dp + g] = wv[i];
        b1s[j * dp + g] = g < d ? __ldg(b1 + g * u + j) : 0.f;
        w2s[j * dp + g] = g < d ? __ldg(w2 + g * u + j) : 0.f;
    }
    const float bias2 = g < d ? __ldg(b2 + g) : 0.f;
    float gam[V], bet[V];  // this thread's LayerNorm affine values
#pragma unroll
    for (int i = 0; i < V; i++) {
        gam[i] = g < d ? __ldg(gamma + g * V + i) : 0.f;
        bet[i] = g < d ? __ldg(beta + g * V + i) : 0.f;
    }
    __syncthreads();
    const long long step = (long long)gridDim.x * HEAD_HG * HEAD_HS;
    for (long long bb = ((long long)blockIdx.x * HEAD_HG + grp) * HEAD_HS; bb < nb; bb += step) {
        float x[HEAD_HS][V], out[HEAD_HS];
#pragma unroll
        for (int s = 0; s < HEAD_HS; s++) {
            const long long b = bb + s < nb ? bb + s : nb - 1;  // tail: recompute the last sample, store is guarded
            const float2 st = __ldg(stats + b);
            out[s] = bias2;
            if (g < d) {
                const float4 *src = reinterpret_cast<const float4 *>(Y + b * h + g * V);
#pragma unroll
                for (int i = 0; i < V / 4; i++) {
                    const float4 t = __ldg(src + i);
                    x[s][4 * i] = fmaxf(fmaf((t.x - st.x) * st.y, gam[4 * i], bet[4 * i]), 0.f);
                    x[s][4 * i + 1] = fmaxf(fmaf((t.y - st.x) * st.y, gam[4 * i + 1], bet[4 * i + 1]), 0.f);
                    x[s][4 * i + 2] = fmaxf(fmaf((t.z - st.x) * st.y, gam[4 * i + 2], bet[4 * i + 2]), 0.f);
                    x[s][4 * i + 3] = fmaxf(fmaf((t.w - st.x) * st.y, gam[4 * i + 3], bet[4 * i + 3]), 0.f);
                }
            } else {
#pragma unroll
                for (int i = 0; i < V; i++) x[s][i] = 0.f;
            }
        }
        for (int j = 0; j < u; j++) {
            float acc[HEAD_HS];
            const float bj = b1s[j * dp + g], wj = w2s[j * dp + g];
#pragma unroll
            for (int s = 0; s < HEAD_HS; s++) acc[s] = bj;
#pragma unroll
            for (int i = 0; i < V; i++) {
                const float w = w1s[(j * V + i) * dp + g];
#pragma unroll
                for (int s = 0; s < HEAD_HS; s++) acc[s] = fmaf(w, x[s][i], acc[s]);
            }
#pragma unroll
            for (int s = 0; s < HEAD_HS; s++) {
                const float e = elu_fast(acc[s]);
                out[s] = fmaf(wj, e, out[s]);
            }
        }
        if (norm) {
            float sq[HEAD_HS];
#pragma unroll
            for (int s = 0; s < HEAD_HS; s++) {
                sq[s] = g < d ? out[s] * out[s] : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq[s] += __shfl_xor_sync(0xffffffffu, sq[s], o);
            }
            float *rg = red + (size_t)grp * HEAD_HS * nw;
            if ((g & 31) == 0) {
#pragma unroll
                for (int s = 0; s < HEAD_HS; s++) rg[s * nw + (g >> 5)] = sq[s];
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(dp) : "memory");
#pragma unroll
            for (int s = 0; s < HEAD_HS; s++) {
                float t = 0.f;
                for (int i = 0; i < nw; i++) t += rg[s * nw + i];
                out[s] = out[s] / fmaxf(sqrtf(t), 1e-12f);  // F.normalize(p=2, eps=1e-12)
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(dp) : "memory");  // red[] free again
        }
        if (g < d) {
#pragma unroll
            for (int s = 0; s < HEAD_HS; s++)
                if (bb + s < nb) z[(bb + s) * d + g] = out[s];
        }
    }
}

// generic fallback (v > 16 or weights too large for shared memory): one CTA per sample, weights from L2
__global__ void head_kernel_generic(const float *Y, const float2 *stats, const float *gamma, const float *beta,
                                    const float *w1, const float *b1, const float *w2, const float *b2, float *z, int d,
                                    int h, int u, int norm, int mode) {
    extern __shared__ float hs[];  // [h] + [32]
    float *red = hs + h;
    const long long b = blockIdx.x;
    const float2 st = stats[b];
    for (int i = threadIdx.x; i < h; i += blockDim.x)
        hs[i] = ln_act(Y[b * h + i], st, __ldg(gamma + i), __ldg(beta + i), mode);
    __syncthreads();
    const int g = threadIdx.x, v = h / d;
    float out = 0.f;
    if (g < d) {
        out = __ldg(b2 + g);
        for (int j = 0; j < u; j++) {
            float acc = __ldg(b1 + g * u + j);
            const float *w = w1 + (long long)(g * u + j) * v;
            for (int i = 0; i < v; i++) acc = fmaf(__ldg(w + i), hs[g * v + i], acc);
            const float e = acc > 0.f ? acc : expm1f(acc);
            out = fmaf(__ldg(w2 + g * u + j), e, out);
        }
    }
    if (norm) {
        float s = g < d ? out * out : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[i];
        out = out / fmaxf(sqrtf(t), 1e-12f);
    }
    if (g < d) z[b * d + g] = out;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int upload(const std::vector<float> &v, float **dst) {
    PF_CUDA(cudaMalloc(dst, sizeof(float) * (v.size() ? v.size() : 1)));
    PF_CUDA(cudaMemcpy(*dst, v.data(), sizeof(float) * v.size(), cudaMemcpyHostToDevice));
    return PFANN_OK;
}

int upload_bf16(const std::vector<float> &v, __nv_bfloat16 **dst) {
    std::vector<__nv_bfloat16> t(v.size());
    for (size_t i = 0; i < v.size(); i++) t[i] = __float2bfloat16_rn(v[i]);
    PF_CUDA(cudaMalloc(dst, sizeof(__nv_bfloat16) * (v.size() ? v.size() : 1)));
    PF_CUDA(cudaMemcpy(*dst, t.data(), sizeof(__nv_bfloat16) * v.size(), cudaMemcpyHostToDevice));
    return PFANN_OK;
}

void free_conv(ConvWeights &c) {
    cudaFree(c.w_kn);
    cudaFree(c.w_nk);
    cudaFree(c.bias);
    cudaFree(c.gamma);
    cudaFree(c.beta);
    cudaFree(c.gb16);
    c = ConvWeights();
}

// live taps of a k=3, stride-s "same" convolution over an axis of length n (model.py:18-19)
void live_taps(int n, int s, int *ntaps, int *tap_k, int *tap_off) {
    const int k = 3;
    const int no = (n - 1) / s + 1;
    const int pad = (n - 1) / s * s + k - n, padl = pad / 2;
    *ntaps = 0;
    for (int j = 0; j < k; j++) {
        bool live = false;
        for (int o = 0; o < no && !live; o++) {
            const int i = o * s + j - padl;
            live = (i >= 0 && i < n);
        }
        if (live) {
            tap_k[*ntaps] = j;
            tap_off[*ntaps] = j - padl;
            (*ntaps)++;
        }
    }
}

const std::vector<float> *find_param(Model *m, const std::string &key, size_t numel) {
    auto it = m->host.find(key);
    if (it == m->host.end()) {
        set_error("pfann_model_finalize: parameter '%s' was never set", key.c_str());
        return nullptr;
    }
    if (it->second.size() != numel) {
        set_error("pfann_model_finalize: parameter '%s' has %zu elements, expected %zu", key.c_str(),
                  it->second.size(), numel);
        return nullptr;
    }
    return &it->second;
}

int finalize_conv(Model *m, int l, int which, const ConvGeom &g) {
    ConvWeights &cw = m->conv[2 * l + which];
    free_conv(cw);
    cw.g = g;
    char key[64];
    const char *cn = which == 0 ? "conv1" : "conv2", *ln = which == 0 ? "ln1" : "ln2";
    const int cin_w = g.depthwise ? 1 : g.Ci;
    snprintf(key, sizeof key, "f.convs.%d.%s.weight", l, cn);
    const std::vector<float> *w = find_param(m, key, (size_t)g.Co * cin_w * 3);
    snprintf(key, sizeof key, "f.convs.%d.%s.bias", l, cn);
    const std::vector<float> *b = find_param(m, key, (size_t)g.Co);
    snprintf(key, sizeof key, "f.convs.%d.%s.weight", l, ln);
    const std::vector<float> *ga = find_param(m, key, (size_t)g.Co * g.Fo * g.To);
    snprintf(key, sizeof key, "f.convs.%d.%s.bias", l, ln);
    const std::vector<float> *be = find_param(m, key, (size_t)g.Co * g.Fo * g.To);
    if (!w || !b || !ga || !be) return PFANN_ERR_STATE;
    // reference weight element order: [Co][Cin][kh][kw] with exactly one of kh/kw == 3 -> [Co][Cin][3]
    if (g.depthwise) {
        std::vector<float> t((size_t)g.Co * g.ntaps);
        for (int o = 0; o < g.Co; o++)
            for (int j = 0; j < g.ntaps; j++) t[(size_t)o * g.ntaps + j] = (*w)[(size_t)o * 3 + g.tap_k[j]];
        PF_TRY(upload(t, &cw.w_kn));
    } else {
        const int K = g.K();
        std::vector<float> kn((size_t)K * g.Co), nk((size_t)K * g.Co);
        for (int o = 0; o < g.Co; o++)
            for (int j = 0; j < g.ntaps; j++)
                for (int c = 0; c < g.Ci; c++) {
                    const float v = (*w)[((size_t)o * g.Ci + c) * 3 + g.tap_k[j]];
                    kn[(size_t)(j * g.Ci + c) * g.Co + o] = v;
                    nk[(size_t)o * K + j * g.Ci + c] = v;
                }
        PF_TRY(upload(kn, &cw.w_kn));
        if (m->precision == PFANN_PRECISION_BF16) PF_TRY(upload_bf16(nk, &cw.w_nk));
    }
    PF_TRY(upload(*b, &cw.bias));
    // LayerNorm affine [Co][Fo][To] -> channels-last [Fo][To][Co]
    std::vector<float> gp(ga->size()), bp(be->size());
    for (int o = 0; o < g.Co; o++)
        for (int f = 0; f < g.Fo; f++)
            for (int t = 0; t < g.To; t++) {
                const size_t src = ((size_t)o * g.Fo + f) * g.To + t, dst = ((size_t)f * g.To + t) * g.Co + o;
                gp[dst] = (*ga)[src];
                bp[dst] = (*be)[src];
            }
    PF_TRY(upload(gp, &cw.gamma));
    PF_TRY(upload(bp, &cw.beta));
    const LnGeom lg = ln_geom(g);
    if (m->precision == PFANN_PRECISION_BF16 && lg.ok) {
        // gamma/beta of every CTA position (rb, nh) of the fused conv+LayerNorm kernel, in the order its epilogue
        // reads shared memory: [32-column chunk][lane quarter][4 x 16 B gamma, 4 x 16 B beta][lane = row][8 channels].
        // Samples shorter than 128 rows repeat cyclically inside the 128-row tile.
        const int R = g.Fo * g.To;
        std::vector<float> gb((size_t)lg.P * 32768);
        for (int p = 0; p < lg.P; p++) {
            const int rb = p / lg.NT, nh = p % lg.NT;
            for (int cb = 0; cb < 4; cb++)
                for (int q = 0; q < 4; q++)
                    for (int jj = 0; jj < 8; jj++)
                        for (int r = 0; r < 32; r++)
                            for (int e = 0; e < 8; e++) {
                                const int row = (rb * 128 + q * 32 + r) % R;
                                const size_t src = (size_t)row * g.Co + nh * 128 + cb * 32 + (jj & 3) * 8 + e;
                                const size_t dst = (size_t)p * 32768 + ((((size_t)cb * 4 + q) * 8 + jj) * 32 + r) * 8 + e;
                                gb[dst] = jj < 4 ? gp[src] : bp[src];
                            }
        }
        PF_TRY(upload_bf16(gb, &cw.gb16));
    }
    return PFANN_OK;
}

template <typename InT>
int launch_conv_fp32(Model *m, const ConvWeights &cw, const InT *X, float *Y, int nb) {
    const ConvGeom &g = cw.g;
    cudaStream_t st = m->ctx->stream;
    ProfScope ps(m->ctx, K_CONV_CC, m->prof_idx);
    if (g.depthwise) {
        const long long total = (long long)nb * g.out_per_sample();
        conv_dw_kernel<InT><<<cdiv(total, 256), 256, 0, st>>>(X, cw.w_kn, cw.bias, Y, total, g.Co, g.Fi, g.Ti, g.Fo,
                                                                 g.To, g.ntaps, g.tap_off[0], g.tap_off[1],
                                                                 g.tap_off[2], g.stride);
    } else {
        ConvArgs<InT> a;
        a.X = X; a.W = cw.w_kn; a.bias = cw.bias; a.Y = Y;
        a.M = (long long)nb * g.rows_per_sample();
        a.Ci = g.Ci; a.Co = g.Co; a.Fi = g.Fi; a.Ti = g.Ti; a.Fo = g.Fo; a.To = g.To;
        a.axis = g.axis; a.ntaps = g.ntaps; a.K = g.K(); a.stride = g.stride; a.trans = 0;
        for (int j = 0; j < 3; j++) a.off[j] = g.tap_off[j];
        dim3 grid(cdiv(a.M, BM), cdiv(g.Co, BN));
        conv_gemm_fp32_kernel<InT><<<grid, 256, 0, st>>>(a);
    }
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

template <typename YT, typename OutT>
int launch_ln_apply(Model *m, const ConvWeights &cw, const YT *Y, OutT *X, int nb) {
    const long long E = cw.g.out_per_sample();
    // enough CTAs to fill the machine first, then amortise the affine over as many samples as possible
    int group = 16;
    while (group > 1 && (long long)cdiv(E, 1024) * cdiv(nb, group) < 2LL * m->ctx->sm_count) group >>= 1;
    dim3 grid(cdiv(E, 1024), cdiv(nb, group));
    ProfScope ps(m->ctx, K_LN, 16 + m->prof_idx);
    if (std::is_same<YT, __nv_bfloat16>::value && std::is_same<OutT, __nv_bfloat16>::value && m->act == 0 && !m->act_first &&
        (E & 7) == 0) {
        const int threads = E / 8 < 256 ? (int)((E / 8 + 31) / 32 * 32) : 256;
        int grp = 16;
        while (grp > 4 && (long long)cdiv(E, 8 * threads) * cdiv(nb, grp) < 4LL * m->ctx->sm_count) grp >>= 1;
        ln_apply_bf16x8_kernel<<<dim3(cdiv(E, 8 * threads), cdiv(nb, grp)), threads, 0, m->ctx->stream>>>(
            reinterpret_cast<const __nv_bfloat16 *>(Y), m->cur_stats, cw.gamma, cw.beta, reinterpret_cast<__nv_bfloat16 *>(X), E,
            nb, grp);
    } else
    ln_apply_kernel<YT, OutT><<<grid, 256, 0, m->ctx->stream>>>(Y, m->cur_stats, cw.gamma, cw.beta, X, E, nb,
                                                                group, m->act | (m->act_first ? 4 : 0));
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

int launch_stats(Model *m, const ConvWeights &cw, const float *Y, int nb) {
    ProfScope ps(m->ctx, K_LN, 16 + m->prof_idx);
    ln_stats_kernel<<<nb, 512, 0, m->ctx->stream>>>(Y, cw.g.out_per_sample(), m->cur_stats,
                                                    m->act | (m->act_first ? 4 : 0));
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

template <typename ActT>
int save_tap(Model *m, int l, const ActT *X, int nb, int total_nb = -1, int s0 = 0) {
    if (m->tap_layer != l) return PFANN_OK;
    const ConvGeom &g = m->conv[2 * l + 1].g;
    if (total_nb < 0) total_nb = nb;
    const long long total = (long long)nb * g.out_per_sample();
    PF_TRY(m->tapbuf.ensure((size_t)total_nb * g.out_per_sample() * 4));  // sized once for the whole call
    to_nchw_kernel<ActT><<<cdiv(total, 256), 256, 0, m->ctx->stream>>>(
        X, m->tapbuf.as<float>() + (long long)s0 * g.out_per_sample(), total, g.Co, g.Fo, g.To);
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    m->tap_numel = (long long)total_nb * g.out_per_sample();
    return PFANN_OK;
}

L0Args l0_args(Model *m, const float *mel) {
    const ConvGeom &g = m->conv[0].g;
    L0Args a;
    a.mel = mel; a.F = g.Fi; a.T = g.Ti; a.To = g.To; a.ntaps = g.ntaps; a.C = g.Co;
    for (int j = 0; j < 3; j++) a.off[j] = g.tap_off[j];
    return a;
}

// ln1 statistics of layer 0 from the 9 moments of the mel tile -> m->cur_stats (one launch)
int launch_l0_stats(Model *m, const float *mel, int nb) {
    const ConvGeom &g = m->conv[0].g;
    cudaStream_t st = m->ctx->stream;
    ProfScope ps(m->ctx, K_LN, 34);
    if (m->cur_moments != nullptr)
        l0_stats_kernel<<<cdiv(nb, 256), 256, 0, st>>>(m->cur_moments, m->l0c, (double)g.Fi * g.To, (double)g.Co,
                                                        m->cur_stats, nb);
    else
        l0_moments_kernel<<<nb, 256, 0, st>>>(l0_args(m, mel), m->l0c, m->cur_stats);
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

template <typename ActT>
int launch_l0_fused(Model *m, const float *mel, ActT *X, int nb) {
    const ConvWeights &cw = m->conv[0];
    const ConvGeom &g = cw.g;
    const L0Args a = l0_args(m, mel);
    cudaStream_t st = m->ctx->stream;
    PF_TRY(launch_l0_stats(m, mel, nb));
    if (sizeof(ActT) == 2 && tc_l0_supported(m))
        return tc_l0(m, mel, m->cur_stats, reinterpret_cast<__nv_bfloat16 *>(X), nb);
    const int cgroups = g.Co / 8, ppb = 256 / cgroups, P = g.Fi * g.To;
    const int group = 32;
    // the 4-positions-per-thread variant needs the three taps to be the contiguous run {o, o+1, o+2}
    const bool run3 = g.ntaps == 3 && g.tap_off[1] == g.tap_off[0] + 1 && g.tap_off[2] == g.tap_off[0] + 2;
    {
        ProfScope ps(m->ctx, K_CONV_CC, 0);
        static const int pp_env = getenv("PFANN_L0_PP") ? atoi(getenv("PFANN_L0_PP")) : 4;
        if (run3 && g.To % 4 == 0 && pp_env == 4) {
            dim3 grid(cdiv(P / 4, ppb), cdiv(nb, group));
            l0_conv_ln_kernel<ActT, 4><<<grid, 256, 0, st>>>(a, m->l0_w, cw.bias, cw.gamma, cw.beta,
                                                             m->cur_stats, X, nb, group);
        } else if (run3 && g.To % 2 == 0 && pp_env == 2) {
            dim3 grid(cdiv(P / 2, ppb), cdiv(nb, group));
            l0_conv_ln_kernel<ActT, 2><<<grid, 256, 0, st>>>(a, m->l0_w, cw.bias, cw.gamma, cw.beta,
                                                             m->cur_stats, X, nb, group);
        } else {
            // one position per thread: mel run = [2 to + off0, +2]; other tap sets go through wr = 0 columns
            dim3 grid(cdiv(P, ppb), cdiv(nb, group));
            l0_conv_ln_kernel<ActT, 1><<<grid, 256, 0, st>>>(a, m->l0_w, cw.bias, cw.gamma, cw.beta,
                                                             m->cur_stats, X, nb, group);
        }
    }
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

// one chunk of nb <= m->chunk samples through the 8 layers + head
template <typename ActT>
int forward_chunk(Model *m, const float *mel, int nb, int norm, float *z) {
    const bool tc = (sizeof(ActT) == 2);
    float *Y = m->ybuf.as<float>();
    ActT *xa = m->xa.as<ActT>(), *xb = m->xb.as<ActT>();
    const int l_begin = 0;
    m->cur_stats = m->stats.as<float2>();
    m->cur_partials = m->partials.as<float2>();
    for (int l = l_begin; l < 8; l++) {
        for (int which = 0; which < 2; which++) {
            const ConvWeights &cw = m->conv[2 * l + which];
            m->prof_idx = 2 * l + which;
            const bool first = (l == 0 && which == 0);
            const bool last = (l == 7 && which == 1);
            bool stats_done = false, ybf = false;
            if (first && tc && tc_front_supported(m)) {
                // the whole first SeparableConv2d in one kernel: log-mel in, X1 out, X0 never touches HBM
                PF_TRY(launch_l0_stats(m, mel, nb));
                m->prof_idx = 1;
                PF_TRY(tc_front(m, mel, m->cur_stats, reinterpret_cast<__nv_bfloat16 *>(xb), nb));
                PF_TRY(save_tap<ActT>(m, 0, xb, nb));
                break;
            }
            if (first && m->l0_fused) {
                // conv1 + ln1 + ReLU in one pass, statistics from the mel moments: nothing else to do
                PF_TRY(launch_l0_fused<ActT>(m, mel, xa, nb));
                continue;
            }
            if (first) {
                // layer 0 conv1: C_in = 1, K = 3 -- CUDA cores in both paths, reads the fp32 mel directly
                PF_TRY(launch_conv_fp32<float>(m, cw, mel, Y, nb));
            } else {
                const ActT *in = which == 0 ? xb : xa;
                if (tc && !last && tc_ln_supported(m, 2 * l + which)) {
                    // conv + LayerNorm + ReLU in one kernel (accumulators stay in TMEM, no raw output)
                    PF_TRY(tc_conv_ln(m, 2 * l + which, reinterpret_cast<const __nv_bfloat16 *>(in),
                                      reinterpret_cast<__nv_bfloat16 *>(which == 0 ? xa : xb), nb));
                    if (which == 1) PF_TRY(save_tap<ActT>(m, l, xb, nb));
                    continue;
                }
                if (tc && tc_supported(cw.g)) {
                    ybf = m->y_bf16 && !last;  // the head reads the last raw output in fp32
                    PF_TRY(tc_conv(m, 2 * l + which, reinterpret_cast<const __nv_bfloat16 *>(in), Y, ybf, nb));
                    stats_done = true;  // statistics come out of the GEMM epilogue (from the fp32 accumulators)
                } else {
                    // depthwise conv2 (fuller == false) and geometries without a tensor-core mapping
                    PF_TRY(launch_conv_fp32<ActT>(m, cw, in, Y, nb));
                }
            }
            if (!stats_done) PF_TRY(launch_stats(m, cw, Y, nb));
            if (last) break;  // the head applies the last LayerNorm itself
            if (ybf)
                PF_TRY((launch_ln_apply<__nv_bfloat16, ActT>(m, cw, reinterpret_cast<const __nv_bfloat16 *>(Y),
                                                             which == 0 ? xa : xb, nb)));
            else
                PF_TRY((launch_ln_apply<float, ActT>(m, cw, Y, which == 0 ? xa : xb, nb)));
            if (which == 1) PF_TRY(save_tap<ActT>(m, l, xb, nb));
        }
    }
    const ConvWeights &last = m->conv[15];
    if (m->tap_layer == 7) {
        // materialise the layer-7 output only when asked for
        PF_TRY((launch_ln_apply<float, ActT>(m, last, Y, xb, nb)));
        PF_TRY(save_tap<ActT>(m, 7, xb, nb));
    }
    PF_CUDA(cudaGetLastError());  // anything left over from the layer loop is not the head's
    const int threads = ((m->d + 31) / 32) * 32;
    ProfScope ps(m->ctx, K_HEAD, 33);
    const int v = m->h / m->d;
    const size_t smem_fast =
        ((size_t)m->u * v * threads + 2 * (size_t)m->u * threads + (size_t)HEAD_HG * HEAD_HS * (threads / 32)) * 4;
    if (!m->variant && (v == 8 || v == 16) && threads <= 128 && smem_fast <= 200 * 1024) {
        // the opt-in is per device/context and per function: set it for the instantiation being launched
        if (v == 8)
            PF_CUDA(cudaFuncSetAttribute(head_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast));
        else
            PF_CUDA(cudaFuncSetAttribute(head_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast));
        const int want = (nb + HEAD_HG * HEAD_HS - 1) / (HEAD_HG * HEAD_HS);
        const int grid = want < m->ctx->sm_count ? want : m->ctx->sm_count;
        if (v == 8)
            head_kernel<8><<<grid, threads * HEAD_HG, smem_fast, m->ctx->stream>>>(
                Y, m->cur_stats, last.gamma, last.beta, m->w1, m->b1, m->w2, m->b2, z, nb, m->d, m->h, m->u, norm);
        else
            head_kernel<16><<<grid, threads * HEAD_HG, smem_fast, m->ctx->stream>>>(
                Y, m->cur_stats, last.gamma, last.beta, m->w1, m->b1, m->w2, m->b2, z, nb, m->d, m->h, m->u, norm);
    } else {
        head_kernel_generic<<<nb, threads, (m->h + 32) * sizeof(float), m->ctx->stream>>>(
            Y, m->cur_stats, last.gamma, last.beta, m->w1, m->b1, m->w2, m->b2, z, m->d, m->h, m->u, norm,
            m->act | (m->act_first ? 4 : 0));
    }
    m->ctx->launches++;
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            cudaFuncAttributes fa8, fa16;
            cudaFuncGetAttributes(&fa8, head_kernel<8>);
            cudaFuncGetAttributes(&fa16, head_kernel<16>);
            set_error("head kernel launch failed: %s (d=%d h=%d u=%d v=%d threads=%d nb=%d smem=%zu; max dynamic smem <8> %d <16> %d)",
                      cudaGetErrorString(e), m->d, m->h, m->u, v, threads, nb, smem_fast, fa8.maxDynamicSharedSizeBytes,
                      fa16.maxDynamicSharedSizeBytes);
            return PFANN_ERR_CUDA;
        }
    }
    return PFANN_OK;
}

}  // namespace

namespace pfann {

int plan_workspace(Model *m) {
    const size_t act = m->precision == PFANN_PRECISION_BF16 ? 2 : 4;
    long long tY = 0, tA = 0, tB = 0;
    for (int i = 0; i < 16; i++) {
        const long long e = m->conv[i].g.out_per_sample();
        // fused layer 0 and the fused conv+LayerNorm kernels never store a raw convolution output
        const bool stores_y = !(i == 0 && m->l0_fused) && !(m->precision == PFANN_PRECISION_BF16 && tc_ln_supported(m, i));
        if (stores_y && e > tY) tY = e;
        if ((i & 1) == 0 && e > tA) tA = e;
        if ((i & 1) == 1 && e > tB) tB = e;
    }
    PF_TRY(m->ybuf.ensure((size_t)(tY ? tY : 1) * m->chunk * 4));
    PF_TRY(m->xa.ensure((size_t)(tA ? tA : 1) * m->chunk * act));
    PF_TRY(m->xb.ensure((size_t)(tB ? tB : 1) * m->chunk * act));
    PF_TRY(m->stats.ensure((size_t)m->chunk * sizeof(float2)));
    return PFANN_OK;
}

}  // namespace pfann

namespace {
}  // namespace

namespace pfann {

// device-pointer forward used by the C-ABI wrappers and by extract.cu
// `moments` (optional, [B][9] doubles from the mel kernel) replaces the layer-0 moments pass
int model_forward_dev(Model *m, const float *mel, int64_t B, int norm, float *z, const double *moments) {
    PF_CHECK(m->precision >= 0, PFANN_ERR_STATE, "pfann_model_forward: call pfann_model_finalize first");
    PF_TRY(plan_workspace(m));
    const long long mel_per = (long long)m->F * m->T;
    for (int64_t b0 = 0; b0 < B; b0 += m->chunk) {
        const int nb = (int)((B - b0) < m->chunk ? (B - b0) : m->chunk);
        m->cur_moments = (moments != nullptr && m->l0_fused) ? moments + b0 * 9 : nullptr;
        if (m->precision == PFANN_PRECISION_BF16)
            PF_TRY(forward_chunk<__nv_bfloat16>(m, mel + b0 * mel_per, nb, norm, z + b0 * m->d));
        else
            PF_TRY(forward_chunk<float>(m, mel + b0 * mel_per, nb, norm, z + b0 * m->d));
    }
    return PFANN_OK;
}

// ---- fp32 building blocks shared with the training path (encoder_train.cu) ----
int enc_conv_f32(Model *m, const ConvWeights &cw, const float *X, float *Y, int nb) {
    return launch_conv_fp32<float>(m, cw, X, Y, nb);
}
int enc_ln_stats_f32(Model *m, const ConvWeights &cw, const float *Y, int nb, float2 *stats) {
    float2 *keep = m->cur_stats;
    m->cur_stats = stats;
    const int rc = launch_stats(m, cw, Y, nb);
    m->cur_stats = keep;
    return rc;
}
int enc_ln_apply_f32(Model *m, const ConvWeights &cw, const float *Y, const float2 *stats, float *X, int nb) {
    float2 *keep = m->cur_stats;
    m->cur_stats = const_cast<float2 *>(stats);
    const int rc = launch_ln_apply<float, float>(m, cw, Y, X, nb);
    m->cur_stats = keep;
    return rc;
}
// dXin[nb][Fi][Ti][Ci] = transposed convolution of dY[nb][Fo][To][Co] with Wt[(tap, o)][c] (dense convs only)
int enc_conv_bwd_data_f32(Model *m, const ConvGeom &g, const float *dY, const float *Wt, float *dX, int nb) {
    ConvArgs<float> a;
    a.X = dY; a.W = Wt; a.bias = nullptr; a.Y = dX;
    a.M = (long long)nb * g.Fi * g.Ti;
    a.Ci = g.Co; a.Co = g.Ci; a.Fi = g.Fo; a.Ti = g.To; a.Fo = g.Fi; a.To = g.Ti;
    a.axis = g.axis; a.ntaps = g.ntaps; a.K = g.ntaps * g.Co; a.stride = g.stride; a.trans = 1;
    for (int j = 0; j < 3; j++) a.off[j] = g.tap_off[j];
    dim3 grid(cdiv(a.M, BM), cdiv(a.Co, BN));
    ProfScope ps(m->ctx, K_CONV_CC, m->prof_idx);
    conv_gemm_fp32_kernel<float><<<grid, 256, 0, m->ctx->stream>>>(a);
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

}  // namespace pfann

extern "C" {

int pfann_model_create(pfann_ctx *hctx, int d, int h, int u, int F, int T, int fuller, pfann_model **out) {
    return pfann_model_create_ex(hctx, d, h, u, F, T, fuller, PFANN_ACT_RELU, 1, nullptr, out);
}

int pfann_model_create_ex(pfann_ctx *hctx, int d, int h, int u, int F, int T, int fuller, int conv_activation,
                          int relu_after_bn, const int *strides, pfann_model **out) {
    PF_CHECK(hctx && out, PFANN_ERR_ARG, "pfann_model_create: NULL argument");
    PF_CHECK(d > 0 && h > 0 && u > 0 && F > 0 && T > 0, PFANN_ERR_ARG, "pfann_model_create: bad dimensions");
    PF_CHECK(h % d == 0, PFANN_ERR_ARG, "h must be divisible by d");  // model.py:112
    PF_CHECK(d <= 1024, PFANN_ERR_UNSUPPORTED, "pfann_model_create: d > 1024 unsupported");
    PF_CHECK(conv_activation == PFANN_ACT_RELU || conv_activation == PFANN_ACT_ELU, PFANN_ERR_ARG,
             "pfann_model_create: conv_activation must be PFANN_ACT_RELU or PFANN_ACT_ELU");  // model.py:7-12
    int f = F, t = T;
    bool custom = false;
    for (int i = 0; i < 8; i++) {
        const int st = strides ? strides[2 * i] : 2, sf = strides ? strides[2 * i + 1] : 2;
        PF_CHECK(st >= 1 && st <= 3 && sf >= 1 && sf <= 3, PFANN_ERR_ARG, "pfann_model_create: strides must be 1..3");
        custom = custom || st != 2 || sf != 2;
        f = (f - 1) / sf + 1;
        t = (t - 1) / st + 1;
    }
    PF_CHECK(f == 1 && t == 1, PFANN_ERR_ARG, "output must be 1x1");  // model.py:94
    Model *m = new Model();
    m->ctx = reinterpret_cast<Ctx *>(hctx);
    m->d = d; m->h = h; m->u = u; m->F = F; m->T = T;
    m->act = conv_activation;
    m->act_first = relu_after_bn == 0;
    for (int i = 0; i < 8; i++) {
        m->st_t[i] = strides ? strides[2 * i] : 2;
        m->st_f[i] = strides ? strides[2 * i + 1] : 2;
    }
    m->variant = custom || m->act != PFANN_ACT_RELU || m->act_first;
    m->fuller = fuller != 0;
    *out = reinterpret_cast<pfann_model *>(m);
    return PFANN_OK;
}

void pfann_model_destroy(pfann_model *hm) {
    Model *m = reinterpret_cast<Model *>(hm);
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    tc_release(m);
    train_release(m);
    for (int i = 0; i < 16; i++) free_conv(m->conv[i]);
    cudaFree(m->w1); cudaFree(m->b1); cudaFree(m->w2); cudaFree(m->b2); cudaFree(m->l0_w);
    cudaFree(m->l0_gb16); cudaFree(m->l0_btile);
    m->ybuf.release(); m->xa.release(); m->xb.release(); m->stats.release(); m->partials.release();
    m->tapbuf.release(); m->melbuf.release(); m->zbuf.release(); m->ln_part.release();
    if (m->ln_err_host) cudaFreeHost(m->ln_err_host);
    m->mombuf.release(); m->melbuf2.release(); m->mombuf2.release();
    delete m;
}

int pfann_model_set_param(pfann_model *hm, const char *name, const float *data, int64_t numel) {
    PF_CHECK(hm && name && data && numel >= 0, PFANN_ERR_ARG, "pfann_model_set_param: bad argument");
    Model *m = reinterpret_cast<Model *>(hm);
    std::vector<float> v((size_t)numel);
    if (is_device_ptr(data)) {
        PF_CUDA(cudaSetDevice(m->ctx->device));
        PF_CUDA(cudaMemcpy(v.data(), data, sizeof(float) * numel, cudaMemcpyDeviceToHost));
    } else {
        memcpy(v.data(), data, sizeof(float) * numel);
    }
    m->host[name] = std::move(v);
    m->precision = -1;  // needs (re)finalize
    return PFANN_OK;
}

int pfann_model_finalize(pfann_model *hm, int precision) {
    PF_CHECK(hm, PFANN_ERR_ARG, "pfann_model_finalize: NULL model");
    PF_CHECK(precision == PFANN_PRECISION_FP32 || precision == PFANN_PRECISION_BF16, PFANN_ERR_ARG,
             "pfann_model_finalize: unknown precision %d", precision);
    Model *m = reinterpret_cast<Model *>(hm);
    PF_CUDA(cudaSetDevice(m->ctx->device));
    tc_release(m);
    train_invalidate(m);   // saved activations and transposed weights belong to the previous parameters
    // option variants have no tensor-core kernels: they run on the CUDA-core fp32 path (still on the GPU)
    m->precision = m->variant ? PFANN_PRECISION_FP32 : precision;
    const int ch[9] = {1, m->d, m->d, 2 * m->d, 2 * m->d, 4 * m->d, 4 * m->d, m->h, m->h};
    int F = m->F, T = m->T;
    for (int l = 0; l < 8; l++) {
        const int F2 = (F - 1) / m->st_f[l] + 1, T2 = (T - 1) / m->st_t[l] + 1;
        ConvGeom g1 = {};
        g1.Ci = ch[l]; g1.Co = ch[l + 1]; g1.Fi = F; g1.Ti = T; g1.Fo = F; g1.To = T2; g1.axis = 0;
        g1.stride = m->st_t[l];
        g1.depthwise = false;
        live_taps(T, g1.stride, &g1.ntaps, g1.tap_k, g1.tap_off);
        ConvGeom g2 = {};
        g2.Ci = ch[l + 1]; g2.Co = ch[l + 1]; g2.Fi = F; g2.Ti = T2; g2.Fo = F2; g2.To = T2; g2.axis = 1;
        g2.stride = m->st_f[l];
        g2.depthwise = !m->fuller;
        live_taps(F, g2.stride, &g2.ntaps, g2.tap_k, g2.tap_off);
        int rc = finalize_conv(m, l, 0, g1);
        if (rc == PFANN_OK) rc = finalize_conv(m, l, 1, g2);
        if (rc != PFANN_OK) {
            m->precision = -1;
            return rc;
        }
        F = F2; T = T2;
    }
    {
        // layer-0 fusion constants (live taps of conv1, channel sums in double)
        const ConvGeom &g = m->conv[0].g;
        const std::vector<float> &w = m->host["f.convs.0.conv1.weight"], &b = m->host["f.convs.0.conv1.bias"];
        Model::L0Consts c = {};
        std::vector<float> wt((size_t)g.Co * g.ntaps);
        for (int o = 0; o < g.Co; o++) {
            c.Bsum += b[o];
            c.B2 += (double)b[o] * b[o];
            for (int j = 0; j < g.ntaps; j++) {
                const double wj = w[(size_t)o * 3 + g.tap_k[j]];
                wt[(size_t)o * g.ntaps + j] = (float)wj;
                c.A[j] += wj;
                c.H[j] += (double)b[o] * wj;
                for (int jj = 0; jj < g.ntaps; jj++) c.G[j][jj] += wj * (double)w[(size_t)o * 3 + g.tap_k[jj]];
            }
        }
        m->l0c = c;
        cudaFree(m->l0_w);
        m->l0_w = nullptr;
        PF_TRY(upload(wt, &m->l0_w));
        cudaFree(m->l0_gb16); cudaFree(m->l0_btile);
        m->l0_gb16 = nullptr; m->l0_btile = nullptr;
        if (m->precision == PFANN_PRECISION_BF16 && g.Co == 128 && (g.Fo * g.To) % 128 == 0) {
            // tensor-core layer-0 kernel: weight tile [128 channels][64 K] bf16 in the 128-byte-swizzled K-major
            // layout (only K columns 0-15 are read): [wh0 wh1 wh2 | wh0 wh1 wh2 | wl0 wl1 wl2 | bh bm bl 0 0 0 0]
            std::vector<__nv_bfloat16> bt(128 * 64, __float2bfloat16_rn(0.f));
            for (int n = 0; n < 128; n++) {
                float wv[3] = {0.f, 0.f, 0.f};
                for (int j = 0; j < g.ntaps; j++) wv[j] = wt[(size_t)n * g.ntaps + j];
                __nv_bfloat16 k[16];
                for (int i = 0; i < 16; i++) k[i] = __float2bfloat16_rn(0.f);
                for (int j = 0; j < 3; j++) {
                    const __nv_bfloat16 wh = __float2bfloat16_rn(wv[j]);
                    const __nv_bfloat16 wl = __float2bfloat16_rn(wv[j] - __bfloat162float(wh));
                    k[j] = wh; k[3 + j] = wh; k[6 + j] = wl;
                }
                const __nv_bfloat16 bh = __float2bfloat16_rn(b[n]);
                const __nv_bfloat16 bm = __float2bfloat16_rn(b[n] - __bfloat162float(bh));
                const __nv_bfloat16 bl = __float2bfloat16_rn(b[n] - __bfloat162float(bh) - __bfloat162float(bm));
                k[9] = bh; k[10] = bm; k[11] = bl;
                for (int c = 0; c < 2; c++)       // logical 16-byte chunk c of row n lives at chunk c ^ (n % 8)
                    for (int e = 0; e < 8; e++) bt[(size_t)n * 64 + ((c ^ (n & 7)) * 8) + e] = k[c * 8 + e];
            }
            PF_CUDA(cudaMalloc(&m->l0_btile, bt.size() * sizeof(__nv_bfloat16)));
            PF_CUDA(cudaMemcpy(m->l0_btile, bt.data(), bt.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
            // ln1 affine per block of 128 positions in the epilogue's register layout (see finalize_conv)
            const int PB = g.Fo * g.To / 128;
            const std::vector<float> &ga = m->host["f.convs.0.ln1.weight"], &be = m->host["f.convs.0.ln1.bias"];
            std::vector<float> gb((size_t)PB * 32768);
            for (int pb = 0; pb < PB; pb++)
                for (int cb = 0; cb < 4; cb++)
                    for (int q = 0; q < 4; q++)
                        for (int jj = 0; jj < 8; jj++)
                            for (int r = 0; r < 32; r++)
                                for (int e = 0; e < 8; e++) {
                                    const int pos = pb * 128 + q * 32 + r, ch = cb * 32 + (jj & 3) * 8 + e;
                                    const size_t src = (size_t)ch * g.Fo * g.To + pos;  // reference order [C][F][T]
                                    const size_t dst = (size_t)pb * 32768 + ((((size_t)cb * 4 + q) * 8 + jj) * 32 + r) * 8 + e;
                                    gb[dst] = jj < 4 ? ga[src] : be[src];
                                }
            PF_TRY(upload_bf16(gb, &m->l0_gb16));
        }
        m->y_bf16 = getenv("PFANN_B200_Y_FP32") == nullptr;
        bool taps_run = true;  // the fused kernel reads the mel run off[0], off[0]+1, off[0]+2
        for (int j = 1; j < g.ntaps; j++) taps_run = taps_run && g.tap_off[j] == g.tap_off[0] + j;
        m->l0_fused = !m->variant && taps_run && g.Ci == 1 && g.Co % 8 == 0 && (256 % (g.Co / 8)) == 0 &&
                      g.Co / 8 <= 256 && getenv("PFANN_B200_NO_L0_FUSION") == nullptr;
    }
    const int v = m->h / m->d;
    const std::vector<float> *w1 = find_param(m, "g.linear1.weight", (size_t)m->d * m->u * v);
    const std::vector<float> *b1 = find_param(m, "g.linear1.bias", (size_t)m->d * m->u);
    const std::vector<float> *w2 = find_param(m, "g.linear2.weight", (size_t)m->d * m->u);
    const std::vector<float> *b2 = find_param(m, "g.linear2.bias", (size_t)m->d);
    if (!w1 || !b1 || !w2 || !b2) {
        m->precision = -1;
        return PFANN_ERR_STATE;
    }
    cudaFree(m->w1); cudaFree(m->b1); cudaFree(m->w2); cudaFree(m->b2);
    PF_TRY(upload(*w1, &m->w1));
    PF_TRY(upload(*b1, &m->b1));
    PF_TRY(upload(*w2, &m->w2));
    PF_TRY(upload(*b2, &m->b2));
    if (m->precision == PFANN_PRECISION_BF16) {
        int rc = tc_prepare(m);
        if (rc != PFANN_OK) {
            m->precision = -1;
            return rc;
        }
    }
    return PFANN_OK;
}

int pfann_model_set_chunk(pfann_model *hm, int chunk) {
    PF_CHECK(hm && chunk > 0 && chunk <= 65535, PFANN_ERR_ARG, "pfann_model_set_chunk: chunk must be in 1..65535");
    Model *m = reinterpret_cast<Model *>(hm);
    m->chunk = chunk;
    if (m->precision >= 0) return pfann_model_finalize(hm, m->precision);  // workspaces, tensor maps
    return PFANN_OK;
}

int pfann_model_set_tap(pfann_model *hm, int layer) {
    PF_CHECK(hm && layer >= -1 && layer < 8, PFANN_ERR_ARG, "pfann_model_set_tap: layer must be -1..7");
    reinterpret_cast<Model *>(hm)->tap_layer = layer;
    return PFANN_OK;
}

int pfann_model_forward(pfann_model *hm, const float *mel, int64_t B, int norm, float *z) {
    PF_CHECK(hm && B >= 0 && (B == 0 || (mel && z)), PFANN_ERR_ARG, "pfann_model_forward: bad argument");
    Model *m = reinterpret_cast<Model *>(hm);
    if (B == 0) return PFANN_OK;
    PF_CUDA(cudaSetDevice(m->ctx->device));
    const size_t in_b = (size_t)B * m->F * m->T * 4, out_b = (size_t)B * m->d * 4;
    const void *xd;
    void *zd;
    PF_TRY(stage_input(m->ctx, 0, mel, in_b, &xd));
    PF_TRY(stage_output(m->ctx, 0, z, out_b, &zd));
    PF_TRY(model_forward_dev(m, (const float *)xd, B, norm, (float *)zd, nullptr));
    PF_TRY(finish_output(m->ctx, 0, z, out_b));
    return tc_ln_check(m, !is_device_ptr(z));
}

int pfann_model_get_activation(pfann_model *hm, int layer, float *out, int64_t numel) {
    PF_CHECK(hm && out, PFANN_ERR_ARG, "pfann_model_get_activation: NULL argument");
    Model *m = reinterpret_cast<Model *>(hm);
    PF_CHECK(layer == m->tap_layer && m->tap_numel > 0, PFANN_ERR_STATE,
             "pfann_model_get_activation: layer %d was not tapped (pfann_model_set_tap before forward)", layer);
    PF_CHECK(numel == m->tap_numel, PFANN_ERR_ARG, "pfann_model_get_activation: expected %lld elements, got %lld",
             m->tap_numel, (long long)numel);
    PF_CUDA(cudaSetDevice(m->ctx->device));
    PF_CUDA(cudaMemcpyAsync(out, m->tapbuf.p, (size_t)numel * 4, cudaMemcpyDefault, m->ctx->stream));
    PF_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return PFANN_OK;
}

}  // extern "C"
