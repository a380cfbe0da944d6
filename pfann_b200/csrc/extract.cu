// extract.cu -- stages 1+2 back to back on the device: the builder / matcher inner loop
// (builder.py:88-99, matcher.py:110-127) and the segmenter tail that feeds it (musicdata.py:82-88).
// The log-mel tile of a chunk never leaves HBM-resident scratch; only PCM goes in and z comes out.
#include <stdlib.h>

#include <vector>

#include "encoder.cuh"
#include "pfann_b200.h"

namespace pfann {
int mel_forward_dev(pfann_mel *h, const float *x_dev, int64_t B, float *out_dev, double *moments, int m_ntaps,
                    const int *m_off);
int mel_forward_pcm_dev(pfann_mel *h, const int16_t *pcm_dev, int64_t n_samples, const int64_t *start_dev,
                        const int32_t *valid_dev, int64_t B, float *out_dev, double *moments, int m_ntaps,
                        const int *m_off, const float *wavf_dev);
Ctx *mel_ctx(pfann_mel *h);
void mel_dims(pfann_mel *h, int *seg_len, int *n_mels, int *T);
int model_forward_dev(Model *m, const float *mel, int64_t B, int norm, float *z, const double *moments);
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
    // the mel kernel also reduces the 9 moments the first LayerNorm needs (taps of layer-0 conv1) while the tile is
    // in its shared memory
    const ConvGeom &g0 = m->conv[0].g;
    double *mom = nullptr;
    if (m->l0_fused && getenv("PFANN_B200_NO_MEL_MOMENTS") == nullptr) {
        PF_TRY(m->mombuf.ensure(sizeof(double) * 9 * (size_t)m->chunk));
        mom = m->mombuf.as<double>();
    }
    for (int64_t b0 = 0; b0 < B; b0 += m->chunk) {
        const int64_t nb = (B - b0) < m->chunk ? (B - b0) : m->chunk;
        PF_TRY(mel_forward_dev(mel, (const float *)xd + b0 * seg_len, nb, m->melbuf.as<float>(), mom, g0.ntaps, g0.tap_off));
        PF_TRY(model_forward_dev(m, m->melbuf.as<float>(), nb, norm, (float *)zd + b0 * m->d, mom));
    }
    PF_TRY(finish_output(m->ctx, 0, z, out_b));
    return tc_ln_check(m, !is_device_ptr(z));
}

}  // extern "C"

namespace {
// clips stored back to back as mono samples (int16 PCM: esz = 2, or fp32 from the GPU ingest: esz = 4)
int extract_clips(pfann_mel *mel, pfann_model *hm, const void *pcm_v, int esz, const int64_t *clip_off, int n_clips,
                  int hop, int norm, float *z, int32_t *seg_counts) {
    const unsigned char *pcm = reinterpret_cast<const unsigned char *>(pcm_v);
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
    Ctx *ctx = m->ctx;
    // Host PCM: the copy is pipelined with the compute -- chunk k + 1's samples travel on a copy stream while chunk
    // k is being fingerprinted, and finished fingerprints go back on a third stream (PCIe is full duplex).
    const bool pipe_in = !is_device_ptr(pcm), pipe_out = !is_device_ptr(z);
    const int64_t n_chunks = (B + m->chunk - 1) / m->chunk;
    if (pipe_in) {
        PF_TRY(ctx->stage_in[0].ensure((size_t)n_samples * esz));
        pd = ctx->stage_in[0].p;
    } else {
        pd = pcm + clip_off[0] * esz;
    }
    PF_TRY(stage_input(ctx, 1, start.data(), (size_t)B * 8, &sd));
    PF_TRY(stage_input(ctx, 2, valid.data(), (size_t)B * 4, &vd));
    PF_TRY(stage_output(ctx, 0, z, out_b, &zd));
    if ((pipe_in || pipe_out) && ctx->copy_in == nullptr) {
        PF_CUDA(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        PF_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    }
    std::vector<cudaEvent_t> ev_in(pipe_in ? n_chunks : 0), ev_done(pipe_out ? n_chunks : 0);
    for (auto &e : ev_in) PF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : ev_done) PF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int64_t copied = 0;  // samples already enqueued for upload
    auto upload_for_chunk = [&](int64_t k) -> int {
        const int64_t last = ((k + 1) * m->chunk < B ? (k + 1) * m->chunk : B) - 1;
        int64_t need = start[last] + seg_len;
        if (need > n_samples) need = n_samples;
        if (need > copied) {
            PF_CUDA(cudaMemcpyAsync((unsigned char *)ctx->stage_in[0].p + copied * esz, pcm + (clip_off[0] + copied) * esz,
                                    (size_t)(need - copied) * esz, cudaMemcpyHostToDevice, ctx->copy_in));
            copied = need;
        }
        PF_CUDA(cudaEventRecord(ev_in[k], ctx->copy_in));
        return PFANN_OK;
    };
    if (pipe_in) {
        // the staging buffer may still be read by work enqueued earlier on the compute stream
        cudaEvent_t ev0;
        PF_CUDA(cudaEventCreateWithFlags(&ev0, cudaEventDisableTiming));
        PF_CUDA(cudaEventRecord(ev0, ctx->stream));
        PF_CUDA(cudaStreamWaitEvent(ctx->copy_in, ev0, 0));
        PF_CUDA(cudaEventDestroy(ev0));
        PF_TRY(upload_for_chunk(0));
    }
    const size_t mel_per = (size_t)m->F * m->T;
    PF_TRY(m->melbuf.ensure(mel_per * 4 * (size_t)m->chunk));
    // the mel kernel also reduces the 9 moments the first LayerNorm needs (taps of layer-0 conv1) while the tile is
    // in its shared memory
    const ConvGeom &g0 = m->conv[0].g;
    const bool want_mom = m->l0_fused && getenv("PFANN_B200_NO_MEL_MOMENTS") == nullptr;
    if (want_mom) PF_TRY(m->mombuf.ensure(sizeof(double) * 9 * (size_t)m->chunk));
    // Overlap: the mel kernel of chunk k + 1 runs on a low-priority stream while chunk k is encoded on a high-priority
    // one.  The cooperative encoder kernels leave SMs idle (the fused layer-0 kernel uses 128 of 148) and have gaps
    // between launches; the short-lived mel CTAs fill them, and because the encoder's CTAs are scheduled first whenever
    // an SM frees up they never wait long for co-residency.  Two mel / moment buffers alternate.
    const bool overlap = n_chunks > 1 && getenv("PFANN_B200_NO_OVERLAP") == nullptr;
    float *melb[2] = {m->melbuf.as<float>(), m->melbuf.as<float>()};
    double *momb[2] = {want_mom ? m->mombuf.as<double>() : nullptr, want_mom ? m->mombuf.as<double>() : nullptr};
    cudaStream_t user = ctx->stream;
    std::vector<cudaEvent_t> ev_mel(overlap ? n_chunks : 0), ev_enc(overlap ? n_chunks : 0);
    if (overlap) {
        PF_TRY(m->melbuf2.ensure(mel_per * 4 * (size_t)m->chunk));
        melb[1] = m->melbuf2.as<float>();
        if (want_mom) {
            PF_TRY(m->mombuf2.ensure(sizeof(double) * 9 * (size_t)m->chunk));
            momb[1] = m->mombuf2.as<double>();
        }
        if (ctx->enc_stream == nullptr) {
            int lo = 0, hi = 0;
            PF_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = least, hi = greatest priority
            PF_CUDA(cudaStreamCreateWithPriority(&ctx->enc_stream, cudaStreamNonBlocking, hi));
            PF_CUDA(cudaStreamCreateWithPriority(&ctx->mel_stream, cudaStreamNonBlocking, lo));
        }
        for (auto &e : ev_mel) PF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : ev_enc) PF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        cudaEvent_t ev0;   // everything enqueued on the caller's stream so far (descriptors, earlier calls) comes first
        PF_CUDA(cudaEventCreateWithFlags(&ev0, cudaEventDisableTiming));
        PF_CUDA(cudaEventRecord(ev0, user));
        PF_CUDA(cudaStreamWaitEvent(ctx->enc_stream, ev0, 0));
        PF_CUDA(cudaStreamWaitEvent(ctx->mel_stream, ev0, 0));
        PF_CUDA(cudaEventDestroy(ev0));
    }
    auto run_mel = [&](int64_t k) -> int {
        const int64_t b0 = k * m->chunk;
        const int64_t nb = (B - b0) < m->chunk ? (B - b0) : m->chunk;
        return mel_forward_pcm_dev(mel, esz == 2 ? (const int16_t *)pd : nullptr, n_samples, (const int64_t *)sd + b0,
                                   (const int32_t *)vd + b0, nb, melb[k & 1], momb[k & 1], g0.ntaps, g0.tap_off,
                                   esz == 4 ? (const float *)pd : nullptr);
    };
    int rc = PFANN_OK;
    if (overlap) {   // mel of chunk 0
        ctx->stream = ctx->mel_stream;
        if (pipe_in) rc = cudaStreamWaitEvent(ctx->mel_stream, ev_in[0], 0) == cudaSuccess ? PFANN_OK : PFANN_ERR_CUDA;
        if (rc == PFANN_OK) rc = run_mel(0);
        if (rc == PFANN_OK && cudaEventRecord(ev_mel[0], ctx->mel_stream) != cudaSuccess) rc = PFANN_ERR_CUDA;
    }
    for (int64_t b0 = 0, k = 0; b0 < B && rc == PFANN_OK; b0 += m->chunk, k++) {
        const int64_t nb = (B - b0) < m->chunk ? (B - b0) : m->chunk;
        if (pipe_in && k + 1 < n_chunks) rc = upload_for_chunk(k + 1);
        if (rc != PFANN_OK) break;
        if (overlap) {
            if (k + 1 < n_chunks) {   // mel of the NEXT chunk first, so that it is in flight while this one is encoded
                ctx->stream = ctx->mel_stream;
                if (k >= 1) cudaStreamWaitEvent(ctx->mel_stream, ev_enc[k - 1], 0);   // its buffers were chunk k - 1's
                if (pipe_in) cudaStreamWaitEvent(ctx->mel_stream, ev_in[k + 1], 0);
                rc = run_mel(k + 1);
                if (rc != PFANN_OK) break;
                cudaEventRecord(ev_mel[k + 1], ctx->mel_stream);
            }
            ctx->stream = ctx->enc_stream;
            cudaStreamWaitEvent(ctx->enc_stream, ev_mel[k], 0);
            rc = model_forward_dev(m, melb[k & 1], nb, norm, (float *)zd + b0 * m->d, momb[k & 1]);
            if (rc != PFANN_OK) break;
            cudaEventRecord(ev_enc[k], ctx->enc_stream);
            if (pipe_out) {
                cudaStreamWaitEvent(ctx->copy_out, ev_enc[k], 0);
                cudaMemcpyAsync(z + b0 * m->d, (float *)zd + b0 * m->d, (size_t)nb * m->d * 4, cudaMemcpyDeviceToHost,
                                ctx->copy_out);
            }
        } else {
            if (pipe_in) cudaStreamWaitEvent(ctx->stream, ev_in[k], 0);
            rc = run_mel(k);
            if (rc == PFANN_OK) rc = model_forward_dev(m, melb[k & 1], nb, norm, (float *)zd + b0 * m->d, momb[k & 1]);
            if (rc == PFANN_OK && pipe_out) {
                cudaEventRecord(ev_done[k], ctx->stream);
                cudaStreamWaitEvent(ctx->copy_out, ev_done[k], 0);
                cudaMemcpyAsync(z + b0 * m->d, (float *)zd + b0 * m->d, (size_t)nb * m->d * 4, cudaMemcpyDeviceToHost,
                                ctx->copy_out);
            }
        }
    }
    ctx->stream = user;
    if (overlap) {
        // the caller's stream continues after the last encoder chunk (and after the mel stream, which is idle by then)
        if (rc == PFANN_OK) {
            cudaStreamWaitEvent(user, ev_enc[n_chunks - 1], 0);
        } else {
            cudaStreamSynchronize(ctx->enc_stream);
            cudaStreamSynchronize(ctx->mel_stream);
        }
    }
    if (cudaGetLastError() != cudaSuccess && rc == PFANN_OK) {
        set_error("pfann_extract: stream / event plumbing failed");
        rc = PFANN_ERR_CUDA;
    }
    // the descriptor vectors are host temporaries: make sure their H2D copies are done before they die
    PF_CUDA(cudaStreamSynchronize(ctx->stream));
    if (pipe_in) PF_CUDA(cudaStreamSynchronize(ctx->copy_in));
    if (pipe_out) PF_CUDA(cudaStreamSynchronize(ctx->copy_out));
    for (auto &e : ev_in) cudaEventDestroy(e);
    for (auto &e : ev_done) cudaEventDestroy(e);
    for (auto &e : ev_mel) cudaEventDestroy(e);
    for (auto &e : ev_enc) cudaEventDestroy(e);
    if (rc != PFANN_OK) return rc;
    return tc_ln_check(m);
}
}  // namespace

extern "C" {

int pfann_extract_pcm16(pfann_mel *mel, pfann_model *hm, const int16_t *pcm, const int64_t *clip_off, int n_clips,
                        int hop, int norm, float *z, int32_t *seg_counts) {
    return extract_clips(mel, hm, pcm, 2, clip_off, n_clips, hop, norm, z, seg_counts);
}

int pfann_extract_f32(pfann_mel *mel, pfann_model *hm, const float *wav, const int64_t *clip_off, int n_clips,
                      int hop, int norm, float *z, int32_t *seg_counts) {
    return extract_clips(mel, hm, wav, 4, clip_off, n_clips, hop, norm, z, seg_counts);
}

}  // extern "C"
