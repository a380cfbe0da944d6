// extract.cu -- stages 1+2 back to back on the device: the builder / matcher inner loop
// (builder.py:88-99, matcher.py:110-127) and the segmenter tail that feeds it (musicdata.py:82-88).
// The log-mel tile of a chunk never leaves HBM-resident scratch; only PCM goes in and z comes out.
#include <vector>

#include "encoder.cuh"
#include "pfann_b200.h"

namespace pfann {
int mel_forward_dev(pfann_mel *h, const float *x_dev, int64_t B, float *out_dev);
int mel_forward_pcm_dev(pfann_mel *h, const int16_t *pcm_dev, int64_t n_samples, const int64_t *start_dev,
                        const int32_t *valid_dev, int64_t B, float *out_dev);
Ctx *mel_ctx(pfann_mel *h);
void mel_dims(pfann_mel *h, int *seg_len, int *n_mels, int *T);
int model_forward_dev(Model *m, const float *mel, int64_t B, int norm, float *z);
}  // namespace pfann

using namespace pfann;

namespace {
int check_pair(pfann_mel *mel, Model *m, int *seg_len) {
    int n_mels, T;
    mel_dims(mel, seg_len, &n_mels, &T);
    PF_CHECK(mel_ctx(mel) == m->ctx, PFANN_ERR_ARG, "extract: mel plan and model belong to different contexts");
    PF_CHECK(n_mels == m->F && T == m->T, PFANN_ERR_ARG, "extract: mel plan gives [%d,%d] but the model expects [%d,%d]",
             n_mels, T, m->F, m->T);
    return PFANN_OK;
}
}  // namespace

extern "C" {

int64_t pfann_count_segments(const int64_t *clip_off, int n_clips, int seg_len, int hop) {
    int64_t tot = 0;
    for (int c = 0; c < n_clips; c++) {
        int64_t len = clip_off[c + 1] - clip_off[c];
        if (len < seg_len) len = seg_len;  // musicdata.py:82-84
        tot += (len - seg_len) / hop + 1;  // musicdata.py:87 (unfold)
    }
    return tot;
}

int pfann_extract_segments(pfann_mel *mel, pfann_model *hm, const float *x, int64_t B, int norm, float *z) {
    PF_CHECK(mel && hm && B >= 0 && (B == 0 || (x && z)), PFANN_ERR_ARG, "pfann_extract_segments: bad argument");
    Model *m = reinterpret_cast<Model *>(hm);
    int seg_len;
    PF_TRY(check_pair(mel, m, &seg_len));
    if (B == 0) return PFANN_OK;
    PF_CUDA(cudaSetDevice(m->ctx->device));
    const size_t out_b = (size_t)B * m->d * 4;
    const void *xd;
    void *zd;
    PF_TRY(stage_input(m->ctx, 0, x, (size_t)B * seg_len * 4, &xd));
    PF_TRY(stage_output(m->ctx, 0, z, out_b, &zd));
    const size_t mel_per = (size_t)m->F * m->T;
    PF_TRY(m->melbuf.ensure(mel_per * 4 * (size_t)m->chunk));
    for (int64_t b0 = 0; b0 < B; b0 += m->chunk) {
        const int64_t nb = (B - b0) < m->chunk ? (B - b0) : m->chunk;
        PF_TRY(mel_forward_dev(mel, (const float *)xd + b0 * seg_len, nb, m->melbuf.as<float>()));
        PF_TRY(model_forward_dev(m, m->melbuf.as<float>(), nb, norm, (float *)zd + b0 * m->d));
    }
    PF_TRY(finish_output(m->ctx, 0, z, out_b));
    return is_device_ptr(z) ? PFANN_OK : tc_ln_check(m);
}

int pfann_extract_pcm16(pfann_mel *mel, pfann_model *hm, const int16_t *pcm, const int64_t *clip_off, int n_clips,
                        int hop, int norm, float *z, int32_t *seg_counts) {
    PF_CHECK(mel && hm && clip_off && n_clips >= 0 && hop > 0, PFANN_ERR_ARG, "pfann_extract_pcm16: bad argument");
    Model *m = reinterpret_cast<Model *>(hm);
    int seg_len;
    PF_TRY(check_pair(mel, m, &seg_len));
    PF_CUDA(cudaSetDevice(m->ctx->device));
    const int64_t n_samples = clip_off[n_clips] - clip_off[0];
    std::vector<int64_t> start;
    std::vector<int32_t> valid;
    for (int c = 0; c < n_clips; c++) {
        const int64_t len = clip_off[c + 1] - clip_off[c];
        PF_CHECK(len >= 0, PFANN_ERR_ARG, "pfann_extract_pcm16: clip_off must be non-decreasing");
        const int64_t padded = len < seg_len ? seg_len : len;
        const int64_t ns = (padded - seg_len) / hop + 1;
        for (int64_t s = 0; s < ns; s++) {
            start.push_back(clip_off[c] - clip_off[0] + s * hop);
            const int64_t rest = len - s * hop;
            valid.push_back((int32_t)(rest < seg_len ? rest : seg_len));
        }
        if (seg_counts) seg_counts[c] = (int32_t)ns;  // builder.py:101 landmarkKey entry
    }
    const int64_t B = (int64_t)start.size();
    if (B == 0) return PFANN_OK;
    PF_CHECK(pcm && z, PFANN_ERR_ARG, "pfann_extract_pcm16: NULL pcm or z");
    const size_t out_b = (size_t)B * m->d * 4;
    const void *pd, *sd, *vd;
    void *zd;
    PF_TRY(stage_input(m->ctx, 0, pcm + clip_off[0], (size_t)n_samples * 2, &pd));
    PF_TRY(stage_input(m->ctx, 1, start.data(), (size_t)B * 8, &sd));
    PF_TRY(stage_input(m->ctx, 2, valid.data(), (size_t)B * 4, &vd));
    PF_TRY(stage_output(m->ctx, 0, z, out_b, &zd));
    const size_t mel_per = (size_t)m->F * m->T;
    PF_TRY(m->melbuf.ensure(mel_per * 4 * (size_t)m->chunk));
    for (int64_t b0 = 0; b0 < B; b0 += m->chunk) {
        const int64_t nb = (B - b0) < m->chunk ? (B - b0) : m->chunk;
        PF_TRY(mel_forward_pcm_dev(mel, (const int16_t *)pd, n_samples, (const int64_t *)sd + b0,
                                   (const int32_t *)vd + b0, nb, m->melbuf.as<float>()));
        PF_TRY(model_forward_dev(m, m->melbuf.as<float>(), nb, norm, (float *)zd + b0 * m->d));
    }
    // the descriptor vectors are host temporaries: make sure their H2D copies are done before they die
    PF_CUDA(cudaStreamSynchronize(m->ctx->stream));
    PF_TRY(finish_output(m->ctx, 0, z, out_b));
    return tc_ln_check(m);
}

}  // extern "C"
