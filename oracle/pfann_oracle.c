/*
 * pfann_oracle.c -- CPU restatement of the pfann hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the *checker* for the CUDA product in pfann_b200/.  Nothing in the product
 * path may call it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs load it.  It restates, in plain C with double-precision
 * accumulation ("exact math" rather than a mirror of torch's fp32 rounding), the
 * algorithms of the reference (all citations relative to /root/reference):
 *
 *   stage 1  datautil/melspec.py:33-50  (+ torchaudio MelSpectrogram / melscale_fbanks)
 *   stage 2  model.py:54-73 (SeparableConv2d.forward), model.py:122-130 (MyG.forward)
 *   stage 3  database.py:121 (faiss IndexFlatIP.search -- third party, un-vendored, version
 *            unpinned: restated as exact fp32 inner product, descending, -1 padded)
 *            cpp/seqscore.cpp:33-136 (seq_score)
 *
 * Parity pinning: the reference ships no golden vectors.  The oracle
 * is pinned against outputs of the reference's own Python code run in the build container
 * (tools/gen_golden.py -> tests/golden/), and seq_score additionally against the
 * reference's own cpp/seqscore.cpp compiled into oracle/_ref/ (oracle/Makefile).
 * faiss itself is absent: for the kNN leg parity is "unpinned" beyond the IndexFlatIP
 * contract stated above.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------
 * Stage 1: log-mel front end
 * ---------------------------------------------------------------------------------------- */

/* torchaudio.functional.melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate,
 * norm=None, mel_scale='htk') as used by melspec.py:19-31.  fb is [n_freqs][n_mels]. */
/* torchaudio's two mel scales (functional._hz_to_mel / _mel_to_hz): 'htk' and 'slaney' (linear below 1 kHz,
 * logarithmic above; melspec.py:30 selects slaney in naf_mode) */
static double hz_to_mel(double f, int slaney)
{
    if (!slaney) return 2595.0 * log10(1.0 + f / 700.0);
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = 1000.0 / (200.0 / 3.0);
    const double logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m, int slaney)
{
    if (!slaney) return 700.0 * (pow(10.0, m / 2595.0) - 1.0);
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = 1000.0 / (200.0 / 3.0);
    const double logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

/* slaney != 0: mel_scale='slaney' AND norm='slaney' (area normalisation 2 / (f[m+2] - f[m])), the pair
 * melspec.py:29-30 selects in naf_mode */
int orc_mel_fbanks_ex(int n_freqs, double f_min, double f_max, int n_mels, int sample_rate, int slaney,
                      float *fb)
{
    double *f_pts = (double *)malloc(sizeof(double) * (n_mels + 2));
    if (!f_pts) return -1;
    double m_min = hz_to_mel(f_min, slaney);
    double m_max = hz_to_mel(f_max, slaney);
    for (int i = 0; i < n_mels + 2; i++) {
        double m = m_min + (m_max - m_min) * (double)i / (double)(n_mels + 1);
        f_pts[i] = mel_to_hz(m, slaney);
    }
    double nyq = (double)(sample_rate / 2);
    for (int k = 0; k < n_freqs; k++) {
        double f = nyq * (double)k / (double)(n_freqs - 1);
        for (int m = 0; m < n_mels; m++) {
            double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
            double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
            double v = down < up ? down : up;
            if (v < 0.0) v = 0.0;
            if (slaney) v *= 2.0 / (f_pts[m + 2] - f_pts[m]);
            fb[(size_t)k * n_mels + m] = (float)v;
        }
    }
    free(f_pts);
    return 0;
}

int orc_mel_fbanks(int n_freqs, double f_min, double f_max, int n_mels, int sample_rate,
                   float *fb)
{
    return orc_mel_fbanks_ex(n_freqs, f_min, f_max, n_mels, sample_rate, 0, fb);
}

/* in-place iterative radix-2 complex FFT, n a power of two, double precision */
static void fft_c2c(double *re, double *im, int n)
{
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        double ang = -2.0 * M_PI / (double)len;
        for (int i = 0; i < n; i += len) {
            for (int k = 0; k < len / 2; k++) {
                double wr = cos(ang * k), wi = sin(ang * k);
                int a = i + k, b = i + k + len / 2;
                double xr = re[b] * wr - im[b] * wi;
                double xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr; im[a] += xi;
            }
        }
    }
}

/* MelSpec.forward for the default option set (melspec.py:33-50 with naf_mode=False,
 * mel_log='log', spec_norm='l2'):  x[B][n] -> out[B][n_mels][T],  T = 1 + n / hop.
 *   L2-normalise row (eps 1e-12)            melspec.py:35-36
 *   reflect-pad n_fft/2, Hann(periodic), rFFT, |.|^2     melspec.py:19-31 (torch.stft)
 *   HTK triangular mel (norm=None)           melspec.py:28-30
 *   + 1e-8, natural log                      melspec.py:41,46                          */
int orc_melspec_ex(const float *x, int B, int n, int sample_rate, int n_fft, int hop,
                   double f_min, double f_max, int n_mels, int naf_mode, int mel_log, int norm_max, float *out);

int orc_melspec(const float *x, int B, int n, int sample_rate, int n_fft, int hop,
                double f_min, double f_max, int n_mels, float *out)
{
    return orc_melspec_ex(x, B, n, sample_rate, n_fft, hop, f_min, f_max, n_mels, 0, 1, 0, out);
}

/* All options of MelSpec (melspec.py:4-50):
 *   naf_mode   power = 1 (magnitude), zero padding instead of reflect, slaney mel scale + norm, + 0.06 instead of + 1e-8
 *   mel_log    0 = none, 1 = natural log, 2 = log10              (melspec.py:43-46)
 *   norm_max   spec_norm == 'max': normalise the waveform by max |x| (F.normalize p = inf) and subtract the
 *              maximum of the (log-)mel tile afterwards           (melspec.py:35,48-49)  */
int orc_melspec_ex(const float *x, int B, int n, int sample_rate, int n_fft, int hop,
                   double f_min, double f_max, int n_mels, int naf_mode, int mel_log, int norm_max, float *out)
{
    if (n_fft & (n_fft - 1)) return -2;
    if (n <= n_fft / 2) return -3; /* reflect pad needs pad < n */
    const int n_freqs = n_fft / 2 + 1;
    const int T = 1 + n / hop;
    const int pad = n_fft / 2;
    float *fb = (float *)malloc(sizeof(float) * (size_t)n_freqs * n_mels);
    double *win = (double *)malloc(sizeof(double) * n_fft);
    if (!fb || !win) return -1;
    orc_mel_fbanks_ex(n_freqs, f_min, f_max, n_mels, sample_rate, naf_mode, fb);
    for (int i = 0; i < n_fft; i++) win[i] = 0.5 - 0.5 * cos(2.0 * M_PI * i / n_fft);
    int rc = 0;
#pragma omp parallel
    {
        double *xp = (double *)malloc(sizeof(double) * (n + 2 * pad));
        double *re = (double *)malloc(sizeof(double) * n_fft);
        double *im = (double *)malloc(sizeof(double) * n_fft);
        double *pw = (double *)malloc(sizeof(double) * n_freqs);
#pragma omp for
        for (int b = 0; b < B; b++) {
            const float *xb = x + (size_t)b * n;
            double ss = 0.0, mx = 0.0;
            for (int i = 0; i < n; i++) {
                ss += (double)xb[i] * xb[i];
                if (fabs((double)xb[i]) > mx) mx = fabs((double)xb[i]);
            }
            double nrm = norm_max ? mx : sqrt(ss);
            if (nrm < 1e-12) nrm = 1e-12;
            for (int i = 0; i < n + 2 * pad; i++) {
                int j = i - pad;
                if (naf_mode) {                      /* pad_mode='constant' */
                    xp[i] = (j < 0 || j >= n) ? 0.0 : (double)xb[j] / nrm;
                    continue;
                }
                if (j < 0) j = -j;
                if (j >= n) j = 2 * (n - 1) - j;
                xp[i] = (double)xb[j] / nrm;
            }
            double tile_max = -1e300;
            for (int t = 0; t < T; t++) {
                for (int i = 0; i < n_fft; i++) {
                    re[i] = xp[(size_t)t * hop + i] * win[i];
                    im[i] = 0.0;
                }
                fft_c2c(re, im, n_fft);
                for (int k = 0; k < n_freqs; k++) {
                    pw[k] = re[k] * re[k] + im[k] * im[k];
                    if (naf_mode) pw[k] = sqrt(pw[k]);   /* power = 1 */
                }
                for (int m = 0; m < n_mels; m++) {
                    double acc = 0.0;
                    for (int k = 0; k < n_freqs; k++) {
                        float w = fb[(size_t)k * n_mels + m];
                        if (w != 0.0f) acc += pw[k] * (double)w;
                    }
                    double v = acc + (naf_mode ? 0.06 : 1e-8);
                    if (mel_log == 1) v = log(v);
                    else if (mel_log == 2) v = log10(v);
                    if (v > tile_max) tile_max = v;
                    out[((size_t)b * n_mels + m) * T + t] = (float)v;
                }
            }
            if (norm_max) {   /* the reference subtracts the fp32 maximum of the fp32 tile (melspec.py:49) */
                float fmx = -3.4e38f;
                for (size_t i = 0; i < (size_t)n_mels * T; i++)
                    if (out[(size_t)b * n_mels * T + i] > fmx) fmx = out[(size_t)b * n_mels * T + i];
                for (size_t i = 0; i < (size_t)n_mels * T; i++) out[(size_t)b * n_mels * T + i] -= fmx;
            }
        }
        free(xp); free(re); free(im); free(pw);
    }
    free(fb); free(win);
    return rc;
}

/* Segment framing of datautil/musicdata.py:82-88 on 16-bit PCM that is already mono at the
 * target rate:  wav = pcm * (1/32768) (musicdata.py:48); zero-pad to >= seg; unfold(seg, hop);
 * subtract the per-row mean.  Returns the number of rows written (n_seg). */
int64_t orc_frame_pcm16(const int16_t *pcm, int64_t n, int seg, int hop, float *rows, int64_t max_rows)
{
    int64_t len = n < seg ? seg : n;
    int64_t n_seg = (len - seg) / hop + 1;
    if (n_seg > max_rows) return -1;
    for (int64_t s = 0; s < n_seg; s++) {
        double mean = 0.0;
        for (int i = 0; i < seg; i++) {
            int64_t p = s * hop + i;
            float v = p < n ? (float)pcm[p] * (1.0f / 32768.0f) : 0.0f;
            rows[s * seg + i] = v;
            mean += v;
        }
        mean /= seg;
        for (int i = 0; i < seg; i++) rows[s * seg + i] = (float)((double)rows[s * seg + i] - mean);
    }
    return n_seg;
}

/* ------------------------------------------------------------------------------------------
 * Stage 2: encoder
 * ---------------------------------------------------------------------------------------- */

static double act_fn(double y, int act)
{
    if (act == 1) return y > 0.0 ? y : expm1(y);   /* ELU(alpha = 1), model.py:10-11 */
    return y > 0.0 ? y : 0.0;
}

/* the two orders of model.py:58-72: relu_after_bn ? act(LN(v)) : LN(act(v)) */
static void layer_norm_act(double *v, size_t n, const float *g, const float *be, int act, int relu_after_bn)
{
    if (!relu_after_bn)
        for (size_t i = 0; i < n; i++) v[i] = act_fn(v[i], act);
    double mean = 0.0;
    for (size_t i = 0; i < n; i++) mean += v[i];
    mean /= (double)n;
    double var = 0.0;
    for (size_t i = 0; i < n; i++) { double d = v[i] - mean; var += d * d; }
    var /= (double)n;
    double rstd = 1.0 / sqrt(var + 1e-5);
    for (size_t i = 0; i < n; i++) {
        double y = (v[i] - mean) * rstd * (double)g[i] + (double)be[i];
        v[i] = relu_after_bn ? act_fn(y, act) : y;
    }
}

static void layer_norm_relu(double *v, size_t n, const float *g, const float *be)
{
    /* torch.nn.LayerNorm over the whole (C,F,T) volume, eps=1e-5, biased variance,
     * per-element affine (model.py:21,30), followed by ReLU (model.py:22,31). */
    double mean = 0.0;
    for (size_t i = 0; i < n; i++) mean += v[i];
    mean /= (double)n;
    double var = 0.0;
    for (size_t i = 0; i < n; i++) { double d = v[i] - mean; var += d * d; }
    var /= (double)n;
    double rstd = 1.0 / sqrt(var + 1e-5);
    for (size_t i = 0; i < n; i++) {
        double y = (v[i] - mean) * rstd * (double)g[i] + (double)be[i];
        v[i] = y > 0.0 ? y : 0.0;
    }
}

/* SeparableConv2d.forward (model.py:54-73) for k=3, stride (2,2), ReLU, relu_after_bn=True.
 *   x   [B][Cin][F][T]      (NCHW like the reference)
 *   w1  [Cout][Cin][1][3]   conv1 along time, stride 2, TF-"same" pad (model.py:18-20)
 *   g1/be1 [Cout][F][T2]    ln1 affine (model.py:21)
 *   w2  fuller ? [Cout][Cout][3][1] : [Cout][1][3][1]  conv2 along freq (model.py:24-29)
 *   g2/be2 [Cout][F2][T2]
 *   y   [B][Cout][F2][T2];  mid (optional) [B][Cout][F][T2] = relu(ln1(conv1))          */
int orc_sepconv_ex(const float *x, int B, int Cin, int Cout, int F, int T,
                   const float *w1, const float *b1, const float *g1, const float *be1,
                   const float *w2, const float *b2, const float *g2, const float *be2,
                   int fuller, int s_t, int s_f, int act, int relu_after_bn, float *y, float *mid);

int orc_sepconv(const float *x, int B, int Cin, int Cout, int F, int T,
                const float *w1, const float *b1, const float *g1, const float *be1,
                const float *w2, const float *b2, const float *g2, const float *be2,
                int fuller, float *y, float *mid)
{
    return orc_sepconv_ex(x, B, Cin, Cout, F, T, w1, b1, g1, be1, w2, b2, g2, be2, fuller, 2, 2, 0, 1, y, mid);
}

/* the same with the options of model.py:15,58-72,84-85: time stride of conv1 / frequency stride of conv2
 * (params['strides']), activation (0 ReLU, 1 ELU), relu_after_bn */
int orc_sepconv_ex(const float *x, int B, int Cin, int Cout, int F, int T,
                   const float *w1, const float *b1, const float *g1, const float *be1,
                   const float *w2, const float *b2, const float *g2, const float *be2,
                   int fuller, int s_t, int s_f, int act, int relu_after_bn, float *y, float *mid)
{
    const int k = 3;
    const int T2 = (T - 1) / s_t + 1, F2 = (F - 1) / s_f + 1;
    const int padT = (T - 1) / s_t * s_t + k - T, padTl = padT / 2;
    const int padF = (F - 1) / s_f * s_f + k - F, padFl = padF / 2;
    const size_t n1 = (size_t)Cout * F * T2, n2 = (size_t)Cout * F2 * T2;
    int rc = 0;
#pragma omp parallel for
    for (int b = 0; b < B; b++) {
        double *a1 = (double *)malloc(sizeof(double) * n1);
        double *a2 = (double *)malloc(sizeof(double) * n2);
        if (!a1 || !a2) { rc = -1; free(a1); free(a2); continue; }
        const float *xb = x + (size_t)b * Cin * F * T;
        for (int o = 0; o < Cout; o++)
            for (int f = 0; f < F; f++)
                for (int t = 0; t < T2; t++) {
                    double acc = b1[o];
                    for (int i = 0; i < Cin; i++)
                        for (int j = 0; j < k; j++) {
                            int tt = t * s_t + j - padTl;
                            if (tt < 0 || tt >= T) continue;
                            acc += (double)w1[((size_t)o * Cin + i) * k + j] *
                                   (double)xb[((size_t)i * F + f) * T + tt];
                        }
                    a1[((size_t)o * F + f) * T2 + t] = acc;
                }
        layer_norm_act(a1, n1, g1, be1, act, relu_after_bn);
        if (mid) for (size_t i = 0; i < n1; i++) mid[(size_t)b * n1 + i] = (float)a1[i];
        for (int o = 0; o < Cout; o++)
            for (int f = 0; f < F2; f++)
                for (int t = 0; t < T2; t++) {
                    double acc = b2[o];
                    for (int j = 0; j < k; j++) {
                        int ff = f * s_f + j - padFl;
                        if (ff < 0 || ff >= F) continue;
                        if (fuller) {
                            for (int i = 0; i < Cout; i++)
                                acc += (double)w2[((size_t)o * Cout + i) * k + j] *
                                       a1[((size_t)i * F + ff) * T2 + t];
                        } else {
                            acc += (double)w2[(size_t)o * k + j] * a1[((size_t)o * F + ff) * T2 + t];
                        }
                    }
                    a2[((size_t)o * F2 + f) * T2 + t] = acc;
                }
        layer_norm_act(a2, n2, g2, be2, act, relu_after_bn);
        for (size_t i = 0; i < n2; i++) y[(size_t)b * n2 + i] = (float)a2[i];
        free(a1); free(a2);
    }
    return rc;
}

/* MyG.forward (model.py:122-130): grouped 1x1 Conv1d h -> d*u (groups=d), ELU, grouped
 * d*u -> d, optional L2 normalise (eps 1e-12).  x [B][h], w1 [d*u][v] (v=h/d), w2 [d][u]. */
int orc_head(const float *x, int B, int d, int h, int u,
             const float *w1, const float *b1, const float *w2, const float *b2,
             int norm, float *z)
{
    if (h % d) return -2;
    const int v = h / d;
    for (int b = 0; b < B; b++) {
        double ss = 0.0;
        double *zz = (double *)malloc(sizeof(double) * d);
        for (int g = 0; g < d; g++) {
            double acc2 = b2[g];
            for (int j = 0; j < u; j++) {
                double acc = b1[g * u + j];
                for (int i = 0; i < v; i++)
                    acc += (double)w1[((size_t)g * u + j) * v + i] * (double)x[(size_t)b * h + g * v + i];
                double e = acc > 0.0 ? acc : expm1(acc);
                acc2 += (double)w2[(size_t)g * u + j] * e;
            }
            zz[g] = acc2;
            ss += acc2 * acc2;
        }
        double nrm = sqrt(ss);
        if (nrm < 1e-12) nrm = 1e-12;
        for (int g = 0; g < d; g++) z[(size_t)b * d + g] = (float)(norm ? zz[g] / nrm : zz[g]);
        free(zz);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Stage 3: search + sequence score
 * ---------------------------------------------------------------------------------------- */

/* Canonical fp32 inner product used by BOTH the oracle and the CUDA product for the exact
 * rescoring of kNN candidates: one fused multiply-add per element, k = 0..d-1 in order. */
static float dot_fma_seq(const float *a, const float *b, int d)
{
    float acc = 0.0f;
    for (int k = 0; k < d; k++) acc = fmaf(a[k], b[k], acc);
    return acc;
}

/* faiss IndexFlatIP.search contract (database.py:121): for every query the top_k database
 * rows by inner product, descending; ties -> lower row id first (our documented choice,
 * faiss' heap order for exact ties is unspecified); when fewer than k rows exist the tail is
 * labels -1 / distances -FLT_MAX (faiss fills the heap with the neutral element). */
int orc_flat_ip_search(const float *db, int64_t N, int d, const float *q, int Q, int k,
                       float *dist, int64_t *labels)
{
#pragma omp parallel for
    for (int qi = 0; qi < Q; qi++) {
        float *bd = dist + (size_t)qi * k;
        int64_t *bl = labels + (size_t)qi * k;
        int cnt = 0;
        for (int64_t i = 0; i < N; i++) {
            float s = dot_fma_seq(q + (size_t)qi * d, db + (size_t)i * d, d);
            /* insertion into a sorted list; (score desc, id asc); ids arrive ascending so a
             * later equal score never displaces an earlier one */
            if (cnt == k && !(s > bd[k - 1])) continue;
            int p = cnt < k ? cnt : k - 1;
            while (p > 0 && s > bd[p - 1]) { bd[p] = bd[p - 1]; bl[p] = bl[p - 1]; p--; }
            bd[p] = s; bl[p] = i;
            if (cnt < k) cnt++;
        }
        for (int p = cnt; p < k; p++) { bd[p] = -FLT_MAX; bl[p] = -1; }
    }
    return 0;
}

typedef struct { int song, off, shift; } cand_t;

static int cand_cmp(const void *a, const void *b)
{
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (x->song != y->song) return x->song < y->song ? -1 : 1;
    if (x->off != y->off) return x->off < y->off ? -1 : 1;
    if (x->shift != y->shift) return x->shift < y->shift ? -1 : 1;
    return 0;
}

/* seq_score (cpp/seqscore.cpp:33-136) with faiss::Index::reconstruct replaced by a direct
 * row pointer into a flat fp32 database `db` [ntotal][d].
 *   candidates = sorted unique (song, label - song_pos[song] - t/fsm, t%fsm)   :49-60
 *   per candidate: mean over its sub-query of <db row, query row>, rows outside the song
 *   skipped (:85-111); inner product accumulated k-sequentially as separate fp32 multiply
 *   and add (:99-102; the reference build line :4 has no -mfma so no contraction).
 *   best song: highest score, ties -> lower song id (:115-124).
 *   song_scores[2s], [2s+1] raised in candidate order with strict > (:126-132).
 * Returns best song id or -1. */
int orc_seq_score(const float *db, int d, const int64_t *song_pos, int n_songs,
                  const float *query, int query_len, const int64_t *labels, int top_k,
                  float *song_scores, int frame_shift_mul, float score_alpha)
{
    size_t cap = (size_t)query_len * top_k, nc = 0;
    cand_t *c = (cand_t *)malloc(sizeof(cand_t) * (cap ? cap : 1));
    for (int t = 0; t < query_len; t++) {
        int tim = t / frame_shift_mul, shift = t % frame_shift_mul;
        for (int i = 0; i < top_k; i++) {
            int64_t lab = labels[(size_t)t * top_k + i];
            if (lab < 0) continue;
            /* idx_to_song_id (seqscore.cpp:23-25): upper_bound over song_pos[0..n_songs) - 1 */
            int lo = 0, hi = n_songs;
            while (lo < hi) { int m = (lo + hi) / 2; if (song_pos[m] <= lab) lo = m + 1; else hi = m; }
            int song = lo - 1;
            c[nc].song = song; c[nc].off = (int)(lab - song_pos[song] - tim); c[nc].shift = shift;
            nc++;
        }
    }
    qsort(c, nc, sizeof(cand_t), cand_cmp);
    size_t nu = 0;
    for (size_t i = 0; i < nc; i++)
        if (nu == 0 || cand_cmp(&c[nu - 1], &c[i]) != 0) c[nu++] = c[i];
    float best = -INFINITY;
    int best_song = -1;
    for (size_t i = 0; i < nu; i++) {
        int song = c[i].song;
        if (song >= n_songs || song < 0) continue;
        int song_len = (int)(song_pos[song + 1] - song_pos[song]);
        int64_t song_start = song_pos[song];
        int t = c[i].off, shift = c[i].shift;
        float sco = 0.0f;
        int my_len = (query_len - shift + frame_shift_mul - 1) / frame_shift_mul;
        for (int j = 0; j < my_len; j++) {
            int qi = j * frame_shift_mul + shift;
            if (t + j < 0 || t + j >= song_len) continue;
            const float *vec = db + (size_t)(song_start + t + j) * d;
            volatile float ip = 0.0f; /* volatile: forbid contraction/reassociation */
            for (int k = 0; k < d; k++) { volatile float p = vec[k] * query[(size_t)qi * d + k]; ip = ip + p; }
            float l2 = 1.0f - 1.0f * ip;
            if (score_alpha == 0.0f) sco += ip;
            else if (score_alpha > 0.0f) sco += expf(-score_alpha * l2 * l2);
        }
        sco /= (float)(my_len > 1 ? my_len : 1);
        float tt = (float)(t * frame_shift_mul - shift);
        if (sco > best || (sco == best && song < best_song)) { best = sco; best_song = song; }
        if (sco > song_scores[2 * (size_t)song]) {
            song_scores[2 * (size_t)song] = sco;
            song_scores[2 * (size_t)song + 1] = tt;
        }
    }
    free(c);
    return best_song;
}

long long orc_version(void) { return 20261017001LL; }
