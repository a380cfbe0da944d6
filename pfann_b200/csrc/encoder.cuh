// encoder.cuh -- stage 2 data structures shared by encoder.cu (pipeline + CUDA-core kernels) and
// encoder_tc.cu (tcgen05 tensor-core convolution GEMMs).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace pfann {

// Activations are channels-last: X[b][f][t][c].  A convolution of the reference's SeparableConv2d
// (model.py:54-73) is then a GEMM  Y[m][n] = sum_kk A[m][kk] W[kk][n] + bias[n]  with
//   m  = (b, fo, to)  flattened output position,   n = output channel,
//   kk = (live tap j, input channel c);  A[m][(j,c)] = X[b][fi][ti][c]  (zero outside the input)
//   axis 0 (conv1, 1x3 along time, stride 2):  fi = fo,           ti = 2 to + off[j]
//   axis 1 (conv2, 3x1 along freq, stride 2):  fi = 2 fo + off[j], ti = to
// off[j] = k_j - pad_left for the TF-"same" padding of model.py:18-19,24-25.  Taps that only ever read
// padding (e.g. two of three when the axis has length 1) are dropped at finalize time.
struct ConvGeom {
    int Ci, Co;          // input / output channels
    int Fi, Ti, Fo, To;  // input / output spatial extent
    int axis;            // 0 = along T, 1 = along F
    int stride = 2;      // along the convolved axis (model.py:84-85; 2 unless params['strides'] says otherwise)
    int ntaps;           // live taps
    int tap_k[3];        // original kernel index of live tap j
    int tap_off[3];      // input offset of live tap j
    bool depthwise;      // conv2 with fuller == false (groups = C, model.py:29)
    long long rows_per_sample() const { return (long long)Fo * To; }
    long long out_per_sample() const { return (long long)Fo * To * Co; }
    int K() const { return ntaps * Ci; }
};

struct ConvWeights {
    ConvGeom g;
    float *w_kn = nullptr;           // fp32 [K][Co]   (CUDA-core path); depthwise: [Co][ntaps]
    __nv_bfloat16 *w_nk = nullptr;   // bf16 [Co][K]   (tensor-core path, K-major)
    float *bias = nullptr;           // [Co]
    float *gamma = nullptr;          // LayerNorm affine permuted to channels-last [Fo][To][Co]
    float *beta = nullptr;
    __nv_bfloat16 *gb16 = nullptr;   // gamma/beta per CTA position of the fused conv+LayerNorm kernel (bf16)
};

struct Model {
    Ctx *ctx;
    int d, h, u, F, T;
    bool fuller;
    // option variants (model.py:58-72,84-85): served by the CUDA-core kernels, whatever precision was asked for
    int act = 0;              // 0 ReLU, 1 ELU
    bool act_first = false;   // relu_after_bn == False: activation BEFORE the LayerNorm
    int st_t[8], st_f[8];     // time stride of conv1 / frequency stride of conv2 per layer
    bool variant = false;
    int precision = -1;
    int chunk = 256;
    std::map<std::string, std::vector<float>> host;  // reference-keyed fp32 parameters (as given)
    ConvWeights conv[16];                            // conv[2l] = conv1 of layer l, conv[2l+1] = conv2
    float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;  // head (model.py:118-120)
    // layer-0 conv1 fusion: sums over output channels that turn 9 moments of the mel tile into the exact
    // LayerNorm statistics of the (never materialised) conv output.  Live taps only.
    struct L0Consts {
        double A[3], G[3][3], H[3], Bsum, B2;
    } l0c;
    float *l0_w = nullptr;  // [Co][ntaps] fp32 (tap-major per channel)
    __nv_bfloat16 *l0_gb16 = nullptr;  // layer-0 ln1 affine per block of 128 positions (tensor-core layer-0 kernel)
    void *l0_btile = nullptr;          // its 16 KB weight tile (bf16 hi/lo split, shared-memory layout)
    bool l0_fused = false;
    bool y_bf16 = true;     // tensor-core convs write their raw output in bf16 (statistics stay fp32)
    DevBuf ybuf, xa, xb, stats, partials, tapbuf, melbuf, zbuf;   // chunk-sized workspace ("tail" phase)
    DevBuf ln_part;          // fused conv+LayerNorm: statistics exchange table
    int *ln_err_host = nullptr, *ln_err_dev = nullptr;  // its time-out flag: pinned, mapped (readable without a sync)
    float2 *cur_stats = nullptr, *cur_partials = nullptr;  // statistics buffers of the phase being executed
    const double *cur_moments = nullptr;  // layer-0 moments of the chunk being executed when the mel kernel made them
    DevBuf mombuf;                        // [chunk][9] doubles (fused extract path)
    DevBuf melbuf2, mombuf2;              // second set: mel of chunk k + 1 runs while chunk k is encoded
    int prof_idx = 0;      // convolution being executed (detail slot of the optional event profile)
    int tap_layer = -1;
    long long tap_numel = 0;
    void *tc_state = nullptr;  // tensor maps etc., owned by encoder_tc.cu
    void *train_state = nullptr;  // saved activations and gradients, owned by encoder_train.cu
};

// encoder_tc.cu
int tc_prepare(Model *m);                      // build per-layer state after weights are on the device
void tc_release(Model *m);
bool tc_supported(const ConvGeom &g);
// Y[m][n] (fp32 or bf16) = conv GEMM of X (bf16, channels-last) for `nb` samples; also writes per-sample
// LayerNorm partial sums into m->partials and reduces them into m->stats (mean, rstd).
int tc_conv(Model *m, int idx, const __nv_bfloat16 *X, void *Y, bool y_bf16, int nb);
// Fused conv + LayerNorm + ReLU (TMEM-resident accumulators, no raw output): Xout[m][n] bf16.  Returns
// PFANN_ERR_UNSUPPORTED (without touching anything) when the geometry has no fused mapping.
struct LnGeom {
    bool ok = false;
    int NT = 0, RB = 0, P = 0;  // 128-channel slices, 128-row blocks per sample, CTA positions per sample group
    int sg_shift = 0;           // log2(samples per 128-row tile)
};
LnGeom ln_geom(const ConvGeom &g);
bool tc_ln_supported(Model *m, int idx);
// Reports (and clears) a timed-out statistics exchange.  sync = true waits for the stream first (calls whose results
// are complete on return); sync = false only looks at the flag as it is now, i.e. reports a time-out of EARLIER
// stream-ordered work (device-pointer calls must not block).
int tc_ln_check(Model *m, bool sync = true);
// layer-0 conv1 + ln1 + ReLU as one K = 16 MMA per 128 positions (statistics from the moments kernel)
bool tc_l0_supported(Model *m);
int tc_l0(Model *m, const float *mel, const float2 *stats, __nv_bfloat16 *X, int nb);
int tc_conv_ln(Model *m, int idx, const __nv_bfloat16 *X, __nv_bfloat16 *Xout, int nb);
int tc_ln_err_ptr(Model *m, int **dev_flag);   // device address of the (lazily created) time-out flag
// front_tc.cu: layer-0 conv1 + ln1 + ReLU + conv2 + ln2 + ReLU in one kernel (log-mel in, X1 out; X0 never stored)
bool tc_front_supported(Model *m);
int tc_front(Model *m, const float *mel, const float2 *stats0, __nv_bfloat16 *Xout, int nb);
// encoder_train.cu
void train_release(Model *m);
void train_invalidate(Model *m);   // parameters changed: drop what was derived from them, keep the buffers
// encoder.cu: size the chunk workspace (needs the conv geometries)
int plan_workspace(Model *m);

}  // namespace pfann
