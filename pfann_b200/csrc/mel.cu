// mel.cu -- stage 1: framed STFT power + HTK mel filter bank + log, one fused kernel.
//
// Replaces datautil/melspec.py:33-50 (MelSpec.forward -> torchaudio MelSpectrogram -> torch.stft + matmul,
// ~8 library launches and three HBM round trips of the [B,513,32] spectrum) and, on the PCM entry point,
// the segmenter tail datautil/musicdata.py:48,82-88.
//
// One CTA (8 warps) per 1-second segment:
//   * the segment (32 000 B fp32 or 16 000 B int16) is pulled into shared memory with ONE bulk async
//     copy (cp.async.bulk, TMA engine) completing on an mbarrier; mean removal (PCM path), the L2 norm
//     and the reflect padding are done in place in shared memory;
//   * each warp transforms whole frames: the 1024-point real FFT is a 512-point complex FFT held in
//     registers (16 complex points per lane): a 16-point in-lane DIF, a twiddle, then a 32-point radix-2
//     DIF ACROSS lanes made of shfl_xor butterflies; only the finished spectrum touches shared memory
//     (for the Z[k] / conj Z[512-k] real-FFT recombination);
//   * the mel projection uses the filter bank's sparsity (<= 7 taps per mel bin for the default config,
//     rows < 39 all zero) instead of the reference's dense [513 x 256] sgemm;
//   * log(. + 1e-8) and the [n_mels][T] tile is written once, coalesced.
// Algorithmic HBM bytes per segment: 32 000 in + 32 768 out (fp32 entry point).
// tools/fft_dataflow_check.py holds a numpy emulation of exactly this dataflow.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "pfann_b200.h"

using namespace pfann;

namespace {

constexpr int NFFT = 1024;
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr int ZSTRIDE = 17;                 // padded row of 16 complex values (bank-conflict free)
constexpr int ZBUF_F2 = 32 * ZSTRIDE;       // float2 per warp

struct MelPlan {
    Ctx *ctx;
    int sample_rate, n_fft, hop, n_mels, seg_len, T;
    int fb_stride;
    int naf_mode = 0, mel_log = 1, norm_max = 0;   // melspec.py:27-49 options (see pfann_mel_create_ex)
    int k_lo = 0;             // smallest spectrum bin with a non-zero mel weight (bins below are never formed)
    float2 *d_tw = nullptr;   // W_1024^k = exp(-2 pi i k / 1024), k in [0, 1024)
    float *d_win = nullptr;   // periodic Hann, 1024
    int *d_fb_start = nullptr, *d_fb_cnt = nullptr;
    float *d_fb_w = nullptr;  // [n_mels][fb_stride]
    // fast path (default options, <= 7 taps per filter): weights tap-major [FB_TAPS][n_mels] against a start index that
    // is clamped so that start + FB_TAPS never leaves the spectrum
    float *d_fb_wt = nullptr;
    int *d_fb_start2 = nullptr;
    int fb_blk_taps[8] = {};
    bool fast = false;
    size_t smem_fast = 0;
    DevBuf seg_start, seg_valid;
    size_t smem_bytes;
};

struct MelArgs {
    const float *x;          // fp32 rows [B][n]           (x != nullptr)
    const int16_t *pcm;      // or int16 PCM + descriptors  (pcm != nullptr)
    const float *wavf;       // or fp32 mono samples + the same descriptors (GPU ingest path, ingest.cu)
    int64_t pcm_len;
    const int64_t *seg_start;
    const int32_t *seg_valid;
    float *out;              // [B][n_mels][T]
    int n, hop, T, n_mels, fb_stride, k_lo;
    int naf_mode, mel_log, norm_max;
    const float2 *tw;
    const float *win;
    const int *fb_start, *fb_cnt;
    const float *fb_w;
    const float *fb_wt;      // fast path tables (see MelPlan)
    const int *fb_start2;
    int fb_blk_taps[8];      // taps actually used by mel bins [32 i, 32 i + 32): the projection stops there
    // optional (fused extract path): 9 moments of the log-mel tile under the taps of layer-0 conv1, so that the
    // encoder needs no second pass over the tile for its first LayerNorm (see encoder.cu: l0_stats_kernel)
    double *moments;         // [B][9]: S0 S1 S2 R00 R01 R02 R11 R12 R22, or nullptr
    int m_ntaps, m_off[3];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

__host__ __device__ constexpr int bitrev4(int i) {
    return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3);
}

// 16-point radix-2 DIF in registers; output a[i] = X[bitrev4(i)].
__device__ __forceinline__ void fft16_dif(float2 (&a)[16]) {
    // W_16^j, j = 0..7
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
    const float2 W[8] = {{1.f, 0.f}, {c1, -s1}, {r2, -r2}, {s1, -c1}, {0.f, -1.f}, {-s1, -c1}, {-r2, -r2}, {-c1, -s1}};
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int half = 8 >> s;
#pragma unroll
        for (int base = 0; base < 16; base += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; j++) {
                float2 u = a[base + j], v = a[base + j + half];
                a[base + j] = cadd(u, v);
                float2 d = csub(u, v);
                a[base + j + half] = (j == 0) ? d : cmul(d, W[j * (8 / half)]);
            }
        }
    }
}

__device__ __forceinline__ float block_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();  // protect red[] from a previous use
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NWARPS; i++) t += red[i];  // same order in every thread: deterministic
    return t;
}

__device__ __forceinline__ float block_max(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = red[0];
#pragma unroll
    for (int i = 1; i < NWARPS; i++) t = fmaxf(t, red[i]);
    return t;
}

template <bool PCM>
__global__ void __launch_bounds__(NTHREADS, 2) mel_kernel(const MelArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = a.n, pad = NFFT / 2;
    float *xs = reinterpret_cast<float *>(smem_raw);                        // [n + NFFT]
    float2 *zbuf = reinterpret_cast<float2 *>(xs + ((n + NFFT + 3) & ~3));  // [NWARPS][ZBUF_F2]
    float *tile = reinterpret_cast<float *>(zbuf + NWARPS * ZBUF_F2);       // [n_mels][T+1] (also int16 stage)
    __shared__ float red[NWARPS];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = blockIdx.x;
    float *x = xs + pad;

    if (tid == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();

    // ---- load the segment ------------------------------------------------------------------
    if (!PCM) {
        const float *src = a.x + b * (int64_t)n;
        const bool bulk = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((n & 3) == 0);
        if (bulk) {
            if (tid == 0) {
                ptx::mbar_expect_tx(&bar, (uint32_t)n * 4u);
                ptx::bulk_g2s(x, src, (uint32_t)n * 4u, &bar);
            }
            ptx::mbar_wait(&bar, 0);
        } else {
            for (int i = tid; i < n; i += NTHREADS) x[i] = __ldg(src + i);
            __syncthreads();
        }
    } else {
        const int64_t start = a.seg_start[b];
        const int valid = a.seg_valid[b];
        float part = 0.f;
        if (a.wavf != nullptr) {
            // already scaled, resampled and mixed to mono on the GPU (ingest.cu): musicdata.py:82-84 (zero-pad)
            const float *src = a.wavf + start;
            for (int i = tid; i < n; i += NTHREADS) {
                const float v = i < valid ? __ldg(src + i) : 0.f;
                x[i] = v;
                part += v;
            }
        } else {
            const int16_t *src = a.pcm + start;
            int16_t *stg = reinterpret_cast<int16_t *>(tile);
            const bool bulk = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((valid & 7) == 0) && valid > 0;
            if (bulk) {
                if (tid == 0) {
                    ptx::mbar_expect_tx(&bar, (uint32_t)valid * 2u);
                    ptx::bulk_g2s(stg, src, (uint32_t)valid * 2u, &bar);
                }
                ptx::mbar_wait(&bar, 0);
            } else {
                for (int i = tid; i < valid; i += NTHREADS) stg[i] = src[i];
                __syncthreads();
            }
            // musicdata.py:48 (x 1/32768 in fp32), :82-84 (zero-pad), :88 (mean removal)
            for (int i = tid; i < n; i += NTHREADS) {
                float v = i < valid ? (float)stg[i] * (1.0f / 32768.0f) : 0.f;
                x[i] = v;
                part += v;
            }
        }
        const float mean = block_sum(part, red) / (float)n;
        for (int i = tid; i < n; i += NTHREADS) x[i] -= mean;
        __syncthreads();
    }

    // ---- normalisation factor (melspec.py:35-36, F.normalize eps 1e-12): L2, or max |x| for spec_norm = 'max' ----
    float scale;
    if (a.norm_max) {
        float part = 0.f;
        for (int i = tid; i < n; i += NTHREADS) part = fmaxf(part, fabsf(x[i]));
        scale = 1.0f / fmaxf(block_max(part, red), 1e-12f);
    } else {
        float part = 0.f;
        for (int i = tid; i < n; i += NTHREADS) part = fmaf(x[i], x[i], part);
        const float ss = block_sum(part, red);
        scale = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    }

    // ---- padding of torch.stft(center=True): 'reflect', or zeros in naf_mode (melspec.py:28) ----
    for (int i = tid; i < pad; i += NTHREADS) {
        xs[i] = a.naf_mode ? 0.f : xs[2 * pad - i];                   // padded[i] = x[pad - i]
        xs[pad + n + i] = a.naf_mode ? 0.f : xs[pad + n - 2 - i];     // padded[pad+n+i] = x[n-2-i]
    }
    __syncthreads();

    // ---- per-lane constants -------------------------------------------------------------------
    float2 twl[16];  // W_512^(lane * k1) for the register holding k1 = bitrev4(i)
#pragma unroll
    for (int i = 0; i < 16; i++) twl[i] = __ldg(a.tw + ((2 * lane * bitrev4(i)) & (NFFT - 1)));
    // cross-lane stage twiddles W_(2 half)^(lane & (half-1)), half = 16,8,4,2,1 -- for the lanes that keep the
    // difference; the lanes that keep the sum multiply by 1, so one code path serves both (sgn = -1 / +1)
    float2 wst[5];
    float sgn[5];
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int half = 16 >> s;
        const bool upper = (lane & half) != 0;
        wst[s] = upper ? __ldg(a.tw + (lane & (half - 1)) * (NFFT / (2 * half))) : make_float2(1.f, 0.f);
        sgn[s] = upper ? -1.f : 1.f;
    }
    const int k2 = (int)(__brev((unsigned)lane) >> 27);
    float2 *zw = zbuf + warp * ZBUF_F2;
    float *pw = reinterpret_cast<float *>(zw);
    const int Tp = a.T + 1;

    for (int t = warp; t < a.T; t += NWARPS) {
        // windowed frame -> z[n'] = xw[2n'] + i xw[2n'+1], lane holds n' = 32 r + lane
        float2 v[16];
        const float *fr = xs + t * a.hop;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int j = 64 * r + 2 * lane;
            const float2 xv = *reinterpret_cast<const float2 *>(fr + j);
            const float2 wv = __ldg(reinterpret_cast<const float2 *>(a.win + j));
            v[r] = make_float2(xv.x * scale * wv.x, xv.y * scale * wv.y);
        }
        fft16_dif(v);
#pragma unroll
        for (int i = 1; i < 16; i++) v[i] = cmul(v[i], twl[i]);  // i = 0 -> k1 = 0 -> twiddle 1
#pragma unroll
        for (int s = 0; s < 5; s++) {
            const int half = 16 >> s;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                float2 p;
                p.x = __shfl_xor_sync(0xffffffffu, v[i].x, half);
                p.y = __shfl_xor_sync(0xffffffffu, v[i].y, half);
                // lower lane: (p + v) * 1, upper lane: (p - v) * w -- the same values as computing both and
                // selecting (x * 1 - y * 0 == x), at half the arithmetic
                const float2 t = make_float2(fmaf(sgn[s], v[i].x, p.x), fmaf(sgn[s], v[i].y, p.y));
                v[i] = cmul(t, wst[s]);
            }
        }
        // Z[k1 + 16 k2] with k1 = bitrev4(i), k2 = bitrev5(lane)
#pragma unroll
        for (int i = 0; i < 16; i++) zw[k2 * ZSTRIDE + bitrev4(i)] = v[i];
        __syncwarp();
        // real-FFT recombination: X[k] = E + W_1024^k O, power spectrum
        float P[17];
#pragma unroll
        for (int it = 0; it < 17; it++) {
            const int k = lane + 32 * it;
            P[it] = 0.f;
            if (32 * it + 31 < a.k_lo) continue;  // warp-uniform: no mel filter reaches these bins (f_min)
            if (k <= NFFT / 2) {
                const int ka = k & 511, kb = (512 - k) & 511;
                const float2 A = zw[(ka >> 4) * ZSTRIDE + (ka & 15)];
                float2 Bc = zw[(kb >> 4) * ZSTRIDE + (kb & 15)];
                Bc.y = -Bc.y;
                const float2 E = make_float2(0.5f * (A.x + Bc.x), 0.5f * (A.y + Bc.y));
                const float2 D = csub(A, Bc);
                const float2 O = make_float2(0.5f * D.y, -0.5f * D.x);
                const float2 X = cadd(E, cmul(__ldg(a.tw + k), O));
                P[it] = X.x * X.x + X.y * X.y;
                if (a.naf_mode) P[it] = sqrtf(P[it]);   // power = 1 (melspec.py:27)
            }
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 17; it++) {
            const int k = lane + 32 * it;
            if (k <= NFFT / 2) pw[k] = P[it];
        }
        __syncwarp();
        // sparse HTK mel projection + log (melspec.py:41,46)
        for (int m = lane; m < a.n_mels; m += 32) {
            const int s0 = __ldg(a.fb_start + m), c = __ldg(a.fb_cnt + m);
            const float *w = a.fb_w + (size_t)m * a.fb_stride;
            float acc = 0.f;
            for (int j = 0; j < c; j++) acc = fmaf(__ldg(w + j), pw[s0 + j], acc);
            const float v = acc + (a.naf_mode ? 0.06f : 1e-8f);              // melspec.py:39,41
            // MUFU.LG2 path: |error| ~1e-6 in the log domain
            tile[m * Tp + t] = a.mel_log == 1 ? __logf(v) : (a.mel_log == 2 ? __log10f(v) : v);
        }
        __syncwarp();
    }
    __syncthreads();
    if (a.norm_max) {   // melspec.py:48-49: x - amax(x) over the (mel, time) tile
        float part = -INFINITY;
        for (int i = tid; i < a.n_mels * a.T; i += NTHREADS) part = fmaxf(part, tile[(i / a.T) * Tp + (i % a.T)]);
        const float mx = block_max(part, red);
        for (int i = tid; i < a.n_mels * a.T; i += NTHREADS) tile[(i / a.T) * Tp + (i % a.T)] -= mx;
        __syncthreads();
    }
    if (a.moments != nullptr) {
        // per thread: fp32 partial sums over its positions (f, to), stride-2 taps along time; across threads: double
        const int To = (a.T + 1) / 2, P = a.n_mels * To;
        // fp32 within a thread and within a warp (512 products), double across the warps: fp64 runs at 1/64 rate
        float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // S0 S1 S2 R00 R01 R02 R11 R12 R22
        for (int p = tid; p < P; p += NTHREADS) {
            const int f = p / To, to = p - f * To;
            float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int t = 2 * to + a.m_off[j];
                if (j < a.m_ntaps && t >= 0 && t < a.T) v[j] = tile[f * Tp + t];
            }
            acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2];
            acc[3] = fmaf(v[0], v[0], acc[3]); acc[4] = fmaf(v[0], v[1], acc[4]); acc[5] = fmaf(v[0], v[2], acc[5]);
            acc[6] = fmaf(v[1], v[1], acc[6]); acc[7] = fmaf(v[1], v[2], acc[7]); acc[8] = fmaf(v[2], v[2], acc[8]);
        }
        double *wred = reinterpret_cast<double *>(zbuf);  // [NWARPS][9]: the FFT scratch is free by now
#pragma unroll
        for (int i = 0; i < 9; i++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 9; i++) wred[warp * 9 + i] = (double)acc[i];
        }
        __syncthreads();
        if (tid < 9) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < NWARPS; w++) t += wred[w * 9 + tid];  // fixed order: deterministic
            a.moments[b * 9 + tid] = t;
        }
    }
    float *o = a.out + b * (int64_t)a.n_mels * a.T;
    const int tot = a.n_mels * a.T;
    for (int i = tid; i < tot; i += NTHREADS) {
        const int m = i / a.T, t = i - m * a.T;
        o[i] = tile[m * Tp + t];
    }
}


// ------------------------------------------------------------------------------------------------
// Round-2 kernel for the default option set (power spectrum, natural log, L2 norm, n_mels = 256, <= 7 taps per mel
// filter).  Same results as mel_kernel within rounding, about half the instructions per frame:
//   * the 32-point cross-lane DFT is ONE shared-memory transpose + a second in-lane 16-point DFT + a single
//     shfl_xor(1) butterfly (was five shuffle stages: 160 SHFL and 480 arithmetic instructions per frame);
//   * the recombination forms bins k and 512-k from one (Z[k], Z[512-k]) pair and one twiddle;
//   * the mel projection is a fixed 7-tap dot product against a tap-major weight table staged once per CTA in shared
//     memory (was a divergent variable-length loop over global memory: a third of the frame's instructions);
//   * the 1/|x| factor of melspec.py:36 is applied once per mel bin (the spectrum is quadratic in the signal);
//   * the option branches are compiled out.
// tools/fft_dataflow_check.py (warp_fft_1024_real_v2) is the numpy emulation of this dataflow.
// ------------------------------------------------------------------------------------------------
constexpr int FB_TAPS = 7;
constexpr int FAST_MELS = 256;
constexpr int ZB2 = 552;   // float2 per warp: Z[k] at k + (k >> 4) + (k >> 8) * 8 (k < 512); transpose tile 16 x 34

__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // * (-i)

// 16-point radix-2 DIF, a[i] = X[bitrev4(i)]; the trivial twiddles (1, -i, W_8, W_8^3) are spelled out
__device__ __forceinline__ void fft16_fast(float2 (&a)[16]) {
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
    // stage 1: half = 8, twiddles W_16^j
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float2 u = a[j], v = a[j + 8];
        a[j] = cadd(u, v);
        a[j + 8] = csub(u, v);
    }
    a[9] = cmul(a[9], make_float2(c1, -s1));
    a[10] = make_float2((a[10].x + a[10].y) * r2, (a[10].y - a[10].x) * r2);
    a[11] = cmul(a[11], make_float2(s1, -c1));
    a[12] = mul_mi(a[12]);
    a[13] = cmul(a[13], make_float2(-s1, -c1));
    a[14] = make_float2((a[14].y - a[14].x) * r2, -(a[14].x + a[14].y) * r2);
    a[15] = cmul(a[15], make_float2(-c1, -s1));
    // stage 2: half = 4, twiddles W_8^j
#pragma unroll
    for (int base = 0; base < 16; base += 8) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 u = a[base + j], v = a[base + j + 4];
            a[base + j] = cadd(u, v);
            a[base + j + 4] = csub(u, v);
        }
        a[base + 5] = make_float2((a[base + 5].x + a[base + 5].y) * r2, (a[base + 5].y - a[base + 5].x) * r2);
        a[base + 6] = mul_mi(a[base + 6]);
        a[base + 7] = make_float2((a[base + 7].y - a[base + 7].x) * r2, -(a[base + 7].x + a[base + 7].y) * r2);
    }
    // stage 3: half = 2, twiddles 1, -i
#pragma unroll
    for (int base = 0; base < 16; base += 4) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const float2 u = a[base + j], v = a[base + j + 2];
            a[base + j] = cadd(u, v);
            a[base + j + 2] = csub(u, v);
        }
        a[base + 3] = mul_mi(a[base + 3]);
    }
    // stage 4: half = 1
#pragma unroll
    for (int base = 0; base < 16; base += 2) {
        const float2 u = a[base], v = a[base + 1];
        a[base] = cadd(u, v);
        a[base + 1] = csub(u, v);
    }
}

__host__ __device__ constexpr int zidx(int k) { return k + (k >> 4) + (k >> 8) * 8; }
// W_32^q = exp(-2 pi i q / 32), q = 0..15
__device__ constexpr float W32R[16] = {1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f, 6.123233996e-17f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f, -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f};
__device__ constexpr float W32I[16] = {-0.000000000e+00f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f, -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f, -1.000000000e+00f, -9.807852804e-01f, -9.238795325e-01f, -8.314696123e-01f, -7.071067812e-01f, -5.555702330e-01f, -3.826834324e-01f, -1.950903220e-01f};

template <bool PCM>
__global__ void __launch_bounds__(NTHREADS, 2) mel_fast_kernel(const MelArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = a.n, pad = NFFT / 2;
    float *xs = reinterpret_cast<float *>(smem_raw);                        // [n + NFFT]
    float2 *zbuf = reinterpret_cast<float2 *>(xs + ((n + NFFT + 3) & ~3));  // [NWARPS][ZB2]
    float *tile = reinterpret_cast<float *>(zbuf + NWARPS * ZB2);           // [256][T+1] (also int16 stage)
    float *fbw = tile + FAST_MELS * (a.T + 1);                              // [FB_TAPS][256]
    __shared__ float red[NWARPS];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = blockIdx.x;
    float *x = xs + pad;

    if (tid == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    for (int i = tid; i < FB_TAPS * FAST_MELS; i += NTHREADS) fbw[i] = __ldg(a.fb_wt + i);

    // ---- load the segment (as mel_kernel) ----
    if (!PCM) {
        const float *src = a.x + b * (int64_t)n;
        const bool bulk = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((n & 3) == 0);
        if (bulk) {
            if (tid == 0) {
                ptx::mbar_expect_tx(&bar, (uint32_t)n * 4u);
                ptx::bulk_g2s(x, src, (uint32_t)n * 4u, &bar);
            }
            ptx::mbar_wait(&bar, 0);
        } else {
            for (int i = tid; i < n; i += NTHREADS) x[i] = __ldg(src + i);
            __syncthreads();
        }
    } else {
        const int64_t start = a.seg_start[b];
        const int valid = a.seg_valid[b];
        float part = 0.f;
        if (a.wavf != nullptr) {
            const float *src = a.wavf + start;
            for (int i = tid; i < n; i += NTHREADS) {
                const float v = i < valid ? __ldg(src + i) : 0.f;
                x[i] = v;
                part += v;
            }
        } else {
            const int16_t *src = a.pcm + start;
            int16_t *stg = reinterpret_cast<int16_t *>(tile);
            const bool bulk = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((valid & 7) == 0) && valid > 0;
            if (bulk) {
                if (tid == 0) {
                    ptx::mbar_expect_tx(&bar, (uint32_t)valid * 2u);
                    ptx::bulk_g2s(stg, src, (uint32_t)valid * 2u, &bar);
                }
                ptx::mbar_wait(&bar, 0);
            } else {
                for (int i = tid; i < valid; i += NTHREADS) stg[i] = src[i];
                __syncthreads();
            }
            for (int i = tid; i < n; i += NTHREADS) {
                float v = i < valid ? (float)stg[i] * (1.0f / 32768.0f) : 0.f;
                x[i] = v;
                part += v;
            }
        }
        const float mean = block_sum(part, red) / (float)n;
        for (int i = tid; i < n; i += NTHREADS) x[i] -= mean;
        __syncthreads();
    }
    float sc2;   // (1 / max(|x|_2, 1e-12))^2, applied to the mel energies
    {
        float part = 0.f;
        for (int i = tid; i < n; i += NTHREADS) part = fmaf(x[i], x[i], part);
        const float ss = block_sum(part, red);
        const float scale = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        sc2 = scale * scale;
    }
    for (int i = tid; i < pad; i += NTHREADS) {
        xs[i] = xs[2 * pad - i];
        xs[pad + n + i] = xs[pad + n - 2 - i];
    }
    __syncthreads();

    // ---- per-lane constants ----
    float2 twl[16];   // W_512^(lane * k1), k1 = bitrev4(i)
#pragma unroll
    for (int i = 0; i < 16; i++) twl[i] = __ldg(a.tw + ((2 * lane * bitrev4(i)) & (NFFT - 1)));
    int s0[FAST_MELS / 32];
#pragma unroll
    for (int i = 0; i < FAST_MELS / 32; i++) s0[i] = __ldg(a.fb_start2 + lane + 32 * i);
    const bool odd = (lane & 1) != 0;
    const float sgn = odd ? -1.f : 1.f;
    float2 *zw = zbuf + warp * ZB2;
    float *pw = reinterpret_cast<float *>(zw);
    const int Tp = a.T + 1;
    const float2 *tr_rd = zw + (lane >> 1) * 34 + (lane & 1);
    const int kz = (lane >> 1) + (odd ? 256 : 0);   // Z index of this lane's results: kz + 16 q

    for (int t = warp; t < a.T; t += NWARPS) {
        float2 v[16];
        const float *fr = xs + t * a.hop;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int j = 64 * r + 2 * lane;
            const float2 xv = *reinterpret_cast<const float2 *>(fr + j);
            const float2 wv = __ldg(reinterpret_cast<const float2 *>(a.win + j));
            v[r] = make_float2(xv.x * wv.x, xv.y * wv.y);
        }
        fft16_fast(v);
#pragma unroll
        for (int i = 1; i < 16; i++) v[i] = cmul(v[i], twl[i]);
        // transpose: Y[k1][n2 = lane] -> lane L holds k1 = L >> 1, n2 = (L & 1) + 2 j
#pragma unroll
        for (int i = 0; i < 16; i++) zw[bitrev4(i) * 34 + lane] = v[i];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = tr_rd[2 * j];
        __syncwarp();
        fft16_fast(v);   // v[i] = E[q] (even lanes) / O[q] (odd lanes), q = bitrev4(i)
        // last butterfly of the 32-point DFT between lanes L and L ^ 1: X[q] = E + W_32^q O, X[q + 16] = E - W_32^q O
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int q = bitrev4(i);
            const float wr = W32R[q], wi = W32I[q];   // compile-time after unrolling
            float2 w = v[i];
            if (q != 0) {
                const float2 m = cmul(v[i], make_float2(wr, wi));
                w.x = odd ? m.x : v[i].x;
                w.y = odd ? m.y : v[i].y;
            }
            float2 p;
            p.x = __shfl_xor_sync(0xffffffffu, w.x, 1);
            p.y = __shfl_xor_sync(0xffffffffu, w.y, 1);
            // even lane: w + p; odd lane: p - w
            zw[zidx(kz + 16 * q) - 0] = make_float2(fmaf(sgn, w.x, p.x), fmaf(sgn, w.y, p.y));
        }
        __syncwarp();
        // recombination, bins k and 512 - k from one pair
        float P1[8], P2[8], Pm = 0.f;
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int k = lane + 32 * it, kb = (512 - k) & 511;
            const float2 A = zw[zidx(k)];
            float2 Bc = zw[zidx(kb)];
            Bc.y = -Bc.y;
            const float2 E = make_float2(0.5f * (A.x + Bc.x), 0.5f * (A.y + Bc.y));
            const float2 D = csub(A, Bc);
            const float2 O = make_float2(0.5f * D.y, -0.5f * D.x);
            const float2 T = cmul(__ldg(a.tw + k), O);
            const float2 X1 = cadd(E, T), X2 = csub(E, T);
            P1[it] = X1.x * X1.x + X1.y * X1.y;
            P2[it] = X2.x * X2.x + X2.y * X2.y;
        }
        if (lane == 0) {   // k = 256 pairs with itself: W_1024^256 = -i
            const float2 A = zw[zidx(256)];
            // E = (Re A, 0), O = (Im A, 0), X = E - i O
            Pm = A.x * A.x + A.y * A.y;
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int k = lane + 32 * it;
            pw[k] = P1[it];
            pw[512 - k] = P2[it];
        }
        if (lane == 0) pw[256] = Pm;
        __syncwarp();
        // mel projection (7 taps from shared memory) + log
#pragma unroll
        for (int i = 0; i < FAST_MELS / 32; i++) {
            const int m = lane + 32 * i;
            const float *pp = pw + s0[i];
            float acc = 0.f;
            const int nt = a.fb_blk_taps[i];   // warp-uniform: low mel filters are narrower than one FFT bin
#pragma unroll
            for (int j = 0; j < FB_TAPS; j++)
                if (j < nt) acc = fmaf(fbw[j * FAST_MELS + m], pp[j], acc);
            tile[m * Tp + t] = __logf(fmaf(acc, sc2, 1e-8f));
        }
        __syncwarp();
    }
    __syncthreads();
    if (a.moments != nullptr) {
        const int To = (a.T + 1) / 2, P = FAST_MELS * To;
        float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int p = tid; p < P; p += NTHREADS) {
            const int f = p / To, to = p - f * To;
            float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int t = 2 * to + a.m_off[j];
                if (j < a.m_ntaps && t >= 0 && t < a.T) v[j] = tile[f * Tp + t];
            }
            acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2];
            acc[3] = fmaf(v[0], v[0], acc[3]); acc[4] = fmaf(v[0], v[1], acc[4]); acc[5] = fmaf(v[0], v[2], acc[5]);
            acc[6] = fmaf(v[1], v[1], acc[6]); acc[7] = fmaf(v[1], v[2], acc[7]); acc[8] = fmaf(v[2], v[2], acc[8]);
        }
        double *wred = reinterpret_cast<double *>(zbuf);
#pragma unroll
        for (int i = 0; i < 9; i++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 9; i++) wred[warp * 9 + i] = (double)acc[i];
        }
        __syncthreads();
        if (tid < 9) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < NWARPS; w++) t += wred[w * 9 + tid];
            a.moments[b * 9 + tid] = t;
        }
    }
    float *o = a.out + b * (int64_t)FAST_MELS * a.T;
    const int tot = FAST_MELS * a.T;
    for (int i = tid; i < tot; i += NTHREADS) {
        const int m = i / a.T, t = i - m * a.T;
        o[i] = tile[m * Tp + t];
    }
}

size_t mel_smem_bytes_fast(int n, int T) {
    size_t xs = (size_t)((n + NFFT + 3) & ~3) * 4;
    size_t z = (size_t)NWARPS * ZB2 * 8;
    size_t tile = (size_t)FAST_MELS * (T + 1) * 4;
    // int16 staging aliases tile + weight table (the table is written before the staging copy completes only when the
    // staging fits in the tile alone, which is checked at plan time)
    return xs + z + tile + (size_t)FB_TAPS * FAST_MELS * 4;
}

size_t mel_smem_bytes(int n, int n_mels, int T) {
    size_t xs = (size_t)((n + NFFT + 3) & ~3) * 4;
    size_t z = (size_t)NWARPS * ZBUF_F2 * 8;
    size_t tile = (size_t)n_mels * (T + 1) * 4;
    if (tile < (size_t)n * 2 + 16) tile = (size_t)n * 2 + 16;  // int16 staging shares the tile region
    return xs + z + tile;
}

int launch_mel(MelPlan *p, const MelArgs &a, int64_t B, bool pcm) {
    if (B == 0) return PFANN_OK;
    PF_CHECK(B <= 0x7fffffffLL, PFANN_ERR_ARG, "mel: too many segments in one call (%lld)", (long long)B);
    // opt-in limit: per device/context, so it is set for the instantiation being launched on every call
    if (pcm)
        PF_CUDA(cudaFuncSetAttribute(mel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
    else
        PF_CUDA(cudaFuncSetAttribute(mel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
    ProfScope ps(p->ctx, K_MEL, 32);
    if (p->fast && getenv("PFANN_B200_MEL_V1") == nullptr) {   // read per call: tests compare the two kernels
        if (pcm) {
            PF_CUDA(cudaFuncSetAttribute(mel_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_fast));
            mel_fast_kernel<true><<<(unsigned)B, NTHREADS, p->smem_fast, p->ctx->stream>>>(a);
        } else {
            PF_CUDA(cudaFuncSetAttribute(mel_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_fast));
            mel_fast_kernel<false><<<(unsigned)B, NTHREADS, p->smem_fast, p->ctx->stream>>>(a);
        }
    } else if (pcm)
        mel_kernel<true><<<(unsigned)B, NTHREADS, p->smem_bytes, p->ctx->stream>>>(a);
    else
        mel_kernel<false><<<(unsigned)B, NTHREADS, p->smem_bytes, p->ctx->stream>>>(a);
    p->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

MelArgs base_args(MelPlan *p) {
    MelArgs a = {};
    a.n = p->seg_len;
    a.hop = p->hop;
    a.T = p->T;
    a.n_mels = p->n_mels;
    a.fb_stride = p->fb_stride;
    a.k_lo = p->k_lo;
    a.naf_mode = p->naf_mode; a.mel_log = p->mel_log; a.norm_max = p->norm_max;
    a.tw = p->d_tw;
    a.win = p->d_win;
    a.fb_start = p->d_fb_start;
    a.fb_cnt = p->d_fb_cnt;
    a.fb_w = p->d_fb_w;
    a.fb_wt = p->d_fb_wt;
    a.fb_start2 = p->d_fb_start2;
    for (int i = 0; i < 8; i++) a.fb_blk_taps[i] = p->fb_blk_taps[i];
    return a;
}

}  // namespace

namespace pfann {
// used by extract.cu: run stage 1 on device-resident inputs into a device buffer
int mel_forward_dev(pfann_mel *h, const float *x_dev, int64_t B, float *out_dev, double *moments, int m_ntaps,
                    const int *m_off) {
    MelPlan *p = reinterpret_cast<MelPlan *>(h);
    MelArgs a = base_args(p);
    a.x = x_dev;
    a.out = out_dev;
    a.moments = moments;
    a.m_ntaps = m_ntaps;
    for (int j = 0; j < 3; j++) a.m_off[j] = (moments && j < m_ntaps) ? m_off[j] : 0;
    return launch_mel(p, a, B, false);
}
int mel_forward_pcm_dev(pfann_mel *h, const int16_t *pcm_dev, int64_t n_samples, const int64_t *start_dev,
                        const int32_t *valid_dev, int64_t B, float *out_dev, double *moments, int m_ntaps,
                        const int *m_off, const float *wavf_dev) {
    MelPlan *p = reinterpret_cast<MelPlan *>(h);
    MelArgs a = base_args(p);
    a.pcm = pcm_dev;
    a.wavf = wavf_dev;
    a.pcm_len = n_samples;
    a.seg_start = start_dev;
    a.seg_valid = valid_dev;
    a.out = out_dev;
    a.moments = moments;
    a.m_ntaps = m_ntaps;
    for (int j = 0; j < 3; j++) a.m_off[j] = (moments && j < m_ntaps) ? m_off[j] : 0;
    return launch_mel(p, a, B, true);
}
Ctx *mel_ctx(pfann_mel *h) { return reinterpret_cast<MelPlan *>(h)->ctx; }
void mel_dims(pfann_mel *h, int *seg_len, int *n_mels, int *T) {
    MelPlan *p = reinterpret_cast<MelPlan *>(h);
    *seg_len = p->seg_len;
    *n_mels = p->n_mels;
    *T = p->T;
}
}  // namespace pfann

extern "C" {

int pfann_mel_create(pfann_ctx *hctx, int sample_rate, int n_fft, int hop, double f_min, double f_max,
                     int n_mels, int seg_len, pfann_mel **out) {
    return pfann_mel_create_ex(hctx, sample_rate, n_fft, hop, f_min, f_max, n_mels, seg_len, 0, PFANN_MEL_LOG_E, 0, out);
}

int pfann_mel_create_ex(pfann_ctx *hctx, int sample_rate, int n_fft, int hop, double f_min, double f_max,
                        int n_mels, int seg_len, int naf_mode, int mel_log, int spec_norm_max, pfann_mel **out) {
    PF_CHECK(hctx && out, PFANN_ERR_ARG, "pfann_mel_create: NULL argument");
    PF_CHECK(mel_log >= PFANN_MEL_LOG_NONE && mel_log <= PFANN_MEL_LOG_10, PFANN_ERR_ARG,
             "pfann_mel_create: mel_log must be 0 (none), 1 (log) or 2 (log10)");
    PF_CHECK(n_fft == NFFT, PFANN_ERR_UNSUPPORTED, "pfann_mel_create: n_fft=%d unsupported (kernel is built for 1024)",
             n_fft);
    PF_CHECK(hop > 0 && (hop % 2) == 0, PFANN_ERR_UNSUPPORTED, "pfann_mel_create: hop must be even (got %d)", hop);
    PF_CHECK(seg_len > n_fft / 2 && seg_len % 4 == 0, PFANN_ERR_UNSUPPORTED,
             "pfann_mel_create: seg_len=%d must be > n_fft/2 and a multiple of 4", seg_len);
    PF_CHECK(n_mels > 0 && n_mels <= 1024, PFANN_ERR_ARG, "pfann_mel_create: bad n_mels %d", n_mels);
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    PF_CUDA(cudaSetDevice(ctx->device));
    MelPlan *p = new MelPlan();
    p->ctx = ctx;
    p->sample_rate = sample_rate;
    p->n_fft = n_fft;
    p->hop = hop;
    p->n_mels = n_mels;
    p->seg_len = seg_len;
    p->naf_mode = naf_mode != 0; p->mel_log = mel_log; p->norm_max = spec_norm_max != 0;
    p->T = 1 + seg_len / hop;
    // the last frame must stay inside the padded signal
    PF_CHECK((p->T - 1) * hop + n_fft <= seg_len + n_fft, PFANN_ERR_ARG, "pfann_mel_create: bad framing");
    p->smem_bytes = mel_smem_bytes(seg_len, n_mels, p->T);
    PF_CHECK(p->smem_bytes <= 227 * 1024, PFANN_ERR_UNSUPPORTED,
             "pfann_mel_create: segment of %d samples needs %zu B of shared memory (> 227 KB)", seg_len,
             p->smem_bytes);

    // tables in double, rounded once to fp32
    std::vector<float2> tw(NFFT);
    std::vector<float> win(NFFT);
    const double PI = 3.14159265358979323846;
    for (int k = 0; k < NFFT; k++) {
        tw[k].x = (float)cos(2.0 * PI * k / NFFT);
        tw[k].y = (float)(-sin(2.0 * PI * k / NFFT));
        win[k] = (float)(0.5 - 0.5 * cos(2.0 * PI * k / NFFT));  // torch.hann_window(periodic=True)
    }
    // torchaudio.functional.melscale_fbanks (melspec.py:19-31), kept sparse: norm=None + mel_scale='htk', or the
    // slaney scale (linear below 1 kHz, logarithmic above) with slaney area normalisation in naf_mode
    const int n_freqs = n_fft / 2 + 1;
    std::vector<double> f_pts(n_mels + 2);
    const bool slaney = p->naf_mode != 0;
    const double f_sp = 200.0 / 3.0, min_log_mel = 1000.0 / f_sp, logstep = log(6.4) / 27.0;
    auto hz_to_mel = [&](double f) {
        if (!slaney) return 2595.0 * log10(1.0 + f / 700.0);
        return f >= 1000.0 ? min_log_mel + log(f / 1000.0) / logstep : f / f_sp;
    };
    auto mel_to_hz = [&](double m) {
        if (!slaney) return 700.0 * (pow(10.0, m / 2595.0) - 1.0);
        return m >= min_log_mel ? 1000.0 * exp(logstep * (m - min_log_mel)) : f_sp * m;
    };
    const double m_min = hz_to_mel(f_min), m_max = hz_to_mel(f_max);
    for (int i = 0; i < n_mels + 2; i++) f_pts[i] = mel_to_hz(m_min + (m_max - m_min) * i / (n_mels + 1));
    std::vector<int> start(n_mels), cnt(n_mels);
    std::vector<std::vector<float>> rows(n_mels);
    int stride = 1;
    const double nyq = (double)(sample_rate / 2);
    for (int m = 0; m < n_mels; m++) {
        int first = -1, last = -2;
        std::vector<float> w(n_freqs, 0.f);
        for (int k = 0; k < n_freqs; k++) {
            double f = nyq * k / (n_freqs - 1);
            double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
            double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
            double v = down < up ? down : up;
            if (v > 0.0) {
                if (slaney) v *= 2.0 / (f_pts[m + 2] - f_pts[m]);
                w[k] = (float)v;
                if (first < 0) first = k;
                last = k;
            }
        }
        start[m] = first < 0 ? 0 : first;
        cnt[m] = first < 0 ? 0 : last - first + 1;
        rows[m].assign(w.begin() + start[m], w.begin() + start[m] + cnt[m]);
        if (cnt[m] > stride) stride = cnt[m];
    }
    p->fb_stride = stride;
    p->k_lo = n_freqs;
    for (int m = 0; m < n_mels; m++)
        if (cnt[m] > 0 && start[m] < p->k_lo) p->k_lo = start[m];
    std::vector<float> fbw((size_t)n_mels * stride, 0.f);
    for (int m = 0; m < n_mels; m++)
        for (int j = 0; j < cnt[m]; j++) fbw[(size_t)m * stride + j] = rows[m][j];

    PF_CUDA(cudaMalloc(&p->d_tw, sizeof(float2) * NFFT));
    PF_CUDA(cudaMalloc(&p->d_win, sizeof(float) * NFFT));
    PF_CUDA(cudaMalloc(&p->d_fb_start, sizeof(int) * n_mels));
    PF_CUDA(cudaMalloc(&p->d_fb_cnt, sizeof(int) * n_mels));
    PF_CUDA(cudaMalloc(&p->d_fb_w, sizeof(float) * fbw.size()));
    PF_CUDA(cudaMemcpy(p->d_tw, tw.data(), sizeof(float2) * NFFT, cudaMemcpyHostToDevice));
    PF_CUDA(cudaMemcpy(p->d_win, win.data(), sizeof(float) * NFFT, cudaMemcpyHostToDevice));
    PF_CUDA(cudaMemcpy(p->d_fb_start, start.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    PF_CUDA(cudaMemcpy(p->d_fb_cnt, cnt.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    PF_CUDA(cudaMemcpy(p->d_fb_w, fbw.data(), sizeof(float) * fbw.size(), cudaMemcpyHostToDevice));
    // fast path tables
    p->smem_fast = mel_smem_bytes_fast(seg_len, p->T);
    p->fast = !p->naf_mode && p->mel_log == PFANN_MEL_LOG_E && !p->norm_max && n_mels == FAST_MELS && stride <= FB_TAPS &&
              p->smem_fast <= 227 * 1024 && (size_t)seg_len * 2 + 16 <= (size_t)FAST_MELS * (p->T + 1) * 4;
    if (p->fast) {
        std::vector<float> wt((size_t)FB_TAPS * n_mels, 0.f);
        std::vector<int> st2(n_mels);
        for (int m = 0; m < n_mels; m++) {
            st2[m] = start[m] < n_freqs - FB_TAPS ? start[m] : n_freqs - FB_TAPS;
            for (int j = 0; j < cnt[m]; j++) wt[(size_t)(start[m] - st2[m] + j) * n_mels + m] = rows[m][j];
            const int used = start[m] - st2[m] + cnt[m];
            if (used > p->fb_blk_taps[m / 32]) p->fb_blk_taps[m / 32] = used;
        }
        PF_CUDA(cudaMalloc(&p->d_fb_wt, sizeof(float) * wt.size()));
        PF_CUDA(cudaMalloc(&p->d_fb_start2, sizeof(int) * n_mels));
        PF_CUDA(cudaMemcpy(p->d_fb_wt, wt.data(), sizeof(float) * wt.size(), cudaMemcpyHostToDevice));
        PF_CUDA(cudaMemcpy(p->d_fb_start2, st2.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    }
    *out = reinterpret_cast<pfann_mel *>(p);
    return PFANN_OK;
}

void pfann_mel_destroy(pfann_mel *h) {
    MelPlan *p = reinterpret_cast<MelPlan *>(h);
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaFree(p->d_tw);
    cudaFree(p->d_win);
    cudaFree(p->d_fb_start);
    cudaFree(p->d_fb_cnt);
    cudaFree(p->d_fb_w);
    cudaFree(p->d_fb_wt);
    cudaFree(p->d_fb_start2);
    p->seg_start.release();
    p->seg_valid.release();
    delete p;
}

int pfann_mel_n_frames(pfann_mel *h) { return h ? reinterpret_cast<MelPlan *>(h)->T : 0; }

int pfann_mel_forward(pfann_mel *h, const float *x, int64_t B, float *out) {
    PF_CHECK(h && (B == 0 || (x && out)), PFANN_ERR_ARG, "pfann_mel_forward: NULL argument");
    PF_CHECK(B >= 0, PFANN_ERR_ARG, "pfann_mel_forward: negative batch");
    MelPlan *p = reinterpret_cast<MelPlan *>(h);
    if (B == 0) return PFANN_OK;
    PF_CUDA(cudaSetDevice(p->ctx->device));
    const size_t in_b = (size_t)B * p->seg_len * 4, out_b = (size_t)B * p->n_mels * p->T * 4;
    const void *xd;
    void *od;
    PF_TRY(stage_input(p->ctx, 0, x, in_b, &xd));
    PF_TRY(stage_output(p->ctx, 0, out, out_b, &od));
    PF_TRY(mel_forward_dev(h, (const float *)xd, B, (float *)od, nullptr, 0, nullptr));
    return finish_output(p->ctx, 0, out, out_b);
}

int pfann_mel_forward_pcm16(pfann_mel *h, const int16_t *pcm, int64_t n_samples, const int64_t *seg_start,
                            const int32_t *seg_valid, int64_t B, float *out) {
    PF_CHECK(h && (B == 0 || (pcm && seg_start && seg_valid && out)), PFANN_ERR_ARG,
             "pfann_mel_forward_pcm16: NULL argument");
    MelPlan *p = reinterpret_cast<MelPlan *>(h);
    if (B == 0) return PFANN_OK;
    PF_CUDA(cudaSetDevice(p->ctx->device));
    const size_t out_b = (size_t)B * p->n_mels * p->T * 4;
    const void *pd, *sd, *vd;
    void *od;
    PF_TRY(stage_input(p->ctx, 0, pcm, (size_t)n_samples * 2, &pd));
    PF_TRY(stage_input(p->ctx, 1, seg_start, (size_t)B * 8, &sd));
    PF_TRY(stage_input(p->ctx, 2, seg_valid, (size_t)B * 4, &vd));
    PF_TRY(stage_output(p->ctx, 0, out, out_b, &od));
    PF_TRY(mel_forward_pcm_dev(h, (const int16_t *)pd, n_samples, (const int64_t *)sd, (const int32_t *)vd, B,
                               (float *)od, nullptr, 0, nullptr, nullptr));
    return finish_output(p->ctx, 0, out, out_b);
}

}  // extern "C"
