// ingest.cu -- the decoder-to-segmenter part of datautil/musicdata.py:29-80 on the GPU: interleaved
// 16-bit PCM of any channel count and sample rate -> planar fp32 -> fractional resampling to the model rate ->
// mono mix with the reference's fake-stereo rule.  Framing, zero padding and mean removal (musicdata.py:82-88) happen
// inside the mel kernel (mel.cu, pfann_extract_f32).
//
// The resampler restates julius.ResampleFrac (julius is a third-party dependency of the reference, unpinned in
// readme.md:20 and absent from the build container: parity for this piece is "unpinned", checked against our own
// numpy restatement in oracle/pfann_oracle.py): windowed-sinc kernels, one per output phase,
//     gcd-reduced rates (old, new);  sr = min(old, new) * rolloff (0.945);  width = ceil(zeros * old / sr), zeros = 24
//     k_i[j] = sinc(t) cos^2(t / zeros / 2),  t = clamp((-i / new + (j - width) / old) * sr, +-zeros) * pi,  normalised
//     y[m * new + i] = sum_j k_i[j] * xpad[m * old + j],  xpad = replicate-pad(x, width, width + old)
//     output length int(new * n / old)
#include <math.h>

#include <numeric>
#include <vector>

#include "common.cuh"
#include "pfann_b200.h"

using namespace pfann;

namespace {

__global__ void pcm16_to_planar_kernel(const int16_t *pcm, long long n, int nch, float *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // frame
    if (i >= n) return;
    for (int c = 0; c < nch; c++) out[(long long)c * n + i] = (float)pcm[i * nch + c] * (1.0f / 32768.0f);  // musicdata.py:47-48
}

// thread = one output sample of one channel; the 2 width + old taps of its phase against the replicate-padded input
__global__ void __launch_bounds__(256) resample_kernel(const float *x, long long n, int nch, const float *kern, int K,
                                                       int width, int old_sr, int new_sr, float *y, long long n_out) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    const int c = blockIdx.y;
    const long long m = o / new_sr;
    const int i = (int)(o - m * new_sr);
    const float *xc = x + (long long)c * n;
    const float *k = kern + (size_t)i * K;
    const long long base = m * old_sr - width;
    float acc = 0.f;
    for (int j = 0; j < K; j++) {
        long long p = base + j;
        p = p < 0 ? 0 : (p >= n ? n - 1 : p);   // F.pad(..., mode='replicate')
        acc = fmaf(__ldg(k + j), __ldg(xc + p), acc);
    }
    y[(long long)c * n_out + o] = acc;
}

// musicdata.py:74-77: powers of the difference and the sum of the two channels (double accumulation, fixed order
// inside a CTA, one partial pair per CTA)
__global__ void __launch_bounds__(256) stereo_power_kernel(const float *x, long long n, double *part) {
    __shared__ double red[2][8];
    double p1 = 0.0, p2 = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float a = x[i], b = x[n + i];
        const float d = a - b, s = a + b;
        p1 += (double)d * d;
        p2 += (double)s * s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        p1 += __shfl_xor_sync(0xffffffffu, p1, o);
        p2 += __shfl_xor_sync(0xffffffffu, p2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = p1;
        red[1][threadIdx.x >> 5] = p2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; w++) {
            a += red[0][w];
            b += red[1][w];
        }
        part[2 * blockIdx.x] = a;
        part[2 * blockIdx.x + 1] = b;
    }
}

// wav.mean(dim=0) with channel 1 negated when the clip is fake stereo with opposite phase (musicdata.py:76-80)
__global__ void mix_kernel(const float *x, long long n, int nch, const double *part, int nparts, float *mono) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float sign1 = 1.f;
    if (nch == 2) {
        double p1 = 0.0, p2 = 0.0;
        for (int k = 0; k < nparts; k++) {   // every thread: same order, same decision
            p1 += part[2 * k];
            p2 += part[2 * k + 1];
        }
        if (p1 > p2 * 1000.0) sign1 = -1.f;
    }
    float s = x[i];
    for (int c = 1; c < nch; c++) s += (c == 1 ? sign1 : 1.f) * x[(long long)c * n + i];
    mono[i] = s / (float)nch;
}

}  // namespace

extern "C" {

int64_t pfann_resample_len(int64_t n, int old_sr, int new_sr) {
    if (n < 0 || old_sr <= 0 || new_sr <= 0) return -1;
    const int g = std::gcd(old_sr, new_sr);
    return (int64_t)(((__int128)(new_sr / g) * n) / (old_sr / g));   // int(new_sr * length / old_sr)
}

int pfann_pcm16_to_planar(pfann_ctx *hctx, const int16_t *pcm, int64_t n_frames, int nch, float *out) {
    PF_CHECK(hctx && n_frames >= 0 && nch >= 1 && nch <= 64 && (n_frames == 0 || (pcm && out)), PFANN_ERR_ARG,
             "pfann_pcm16_to_planar: bad argument");
    if (n_frames == 0) return PFANN_OK;
    PF_CHECK(is_device_ptr(pcm) && is_device_ptr(out), PFANN_ERR_ARG, "pfann_pcm16_to_planar: device pointers only");
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    PF_CUDA(cudaSetDevice(ctx->device));
    ProfScope ps(ctx, K_MISC);
    pcm16_to_planar_kernel<<<cdiv(n_frames, 256), 256, 0, ctx->stream>>>(pcm, n_frames, nch, out);
    ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

int pfann_resample_frac(pfann_ctx *hctx, const float *x, int nch, int64_t n, int old_sr, int new_sr, float *y) {
    PF_CHECK(hctx && nch >= 1 && n >= 0 && old_sr > 0 && new_sr > 0 && (n == 0 || (x && y)), PFANN_ERR_ARG,
             "pfann_resample_frac: bad argument");
    if (n == 0) return PFANN_OK;
    PF_CHECK(is_device_ptr(x) && is_device_ptr(y), PFANN_ERR_ARG, "pfann_resample_frac: device pointers only");
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    PF_CUDA(cudaSetDevice(ctx->device));
    const int g = std::gcd(old_sr, new_sr);
    const int o = old_sr / g, w = new_sr / g;
    const int64_t n_out = pfann_resample_len(n, old_sr, new_sr);
    if (o == w) {   // julius short-circuits equal rates: identity
        PF_CUDA(cudaMemcpyAsync(y, x, sizeof(float) * (size_t)nch * n, cudaMemcpyDeviceToDevice, ctx->stream));
        return PFANN_OK;
    }
    const double zeros = 24.0, rolloff = 0.945, PI = 3.14159265358979323846;
    const double sr = (double)(o < w ? o : w) * rolloff;
    const int width = (int)ceil(zeros * o / sr);
    const int K = 2 * width + o;
    PF_CHECK((size_t)w * K <= (64u << 20), PFANN_ERR_UNSUPPORTED, "pfann_resample_frac: %d -> %d needs %d x %d taps", old_sr,
             new_sr, w, K);
    std::vector<float> kern((size_t)w * K);
    for (int i = 0; i < w; i++) {
        double sum = 0.0;
        std::vector<double> row(K);
        for (int j = 0; j < K; j++) {
            double t = (-(double)i / w + (double)(j - width) / o) * sr;
            t = t < -zeros ? -zeros : (t > zeros ? zeros : t);
            t *= PI;
            const double win = cos(t / zeros / 2.0) * cos(t / zeros / 2.0);
            const double sinc = t == 0.0 ? 1.0 : sin(t) / t;
            row[j] = sinc * win;
            sum += row[j];
        }
        for (int j = 0; j < K; j++) kern[(size_t)i * K + j] = (float)(row[j] / sum);
    }
    PF_TRY(ctx->stage_in[3].ensure(kern.size() * sizeof(float)));
    PF_CUDA(cudaMemcpyAsync(ctx->stage_in[3].p, kern.data(), kern.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    PF_CUDA(cudaStreamSynchronize(ctx->stream));   // `kern` is a host temporary
    if (n_out > 0) {
        ProfScope ps(ctx, K_MISC);
        resample_kernel<<<dim3(cdiv(n_out, 256), (unsigned)nch), 256, 0, ctx->stream>>>(x, n, nch, ctx->stage_in[3].as<float>(), K,
                                                                                        width, o, w, y, n_out);
        ctx->launches++;
        PF_CUDA(cudaGetLastError());
    }
    return PFANN_OK;
}

int pfann_mix_mono(pfann_ctx *hctx, const float *x, int nch, int64_t n, float *mono) {
    PF_CHECK(hctx && nch >= 1 && n >= 0 && (n == 0 || (x && mono)), PFANN_ERR_ARG, "pfann_mix_mono: bad argument");
    if (n == 0) return PFANN_OK;
    PF_CHECK(is_device_ptr(x) && is_device_ptr(mono), PFANN_ERR_ARG, "pfann_mix_mono: device pointers only");
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    PF_CUDA(cudaSetDevice(ctx->device));
    const int nparts = 64;
    PF_TRY(ctx->stage_out[3].ensure(sizeof(double) * 2 * nparts));
    double *part = ctx->stage_out[3].as<double>();
    ProfScope ps(ctx, K_MISC);
    if (nch == 2) {
        stereo_power_kernel<<<nparts, 256, 0, ctx->stream>>>(x, n, part);
        ctx->launches++;
    }
    mix_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(x, n, nch, part, nparts, mono);
    ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

}  // extern "C"
