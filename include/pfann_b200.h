/*
 * pfann_b200.h -- C-ABI of libpfann_b200.so: the B200 (sm_100a) implementation of pfann's hot path
 *
 *     8 kHz segments -> log-mel -> conv encoder + split head + L2 -> inner-product kNN -> sequence score
 *
 * This is the drop-in boundary.  Plain C: opaque handles, caller-owned buffers,
 * plain pointers and sizes, int return codes (0 = ok, <0 = error, text via pfann_last_error()); no
 * exception crosses the boundary and there is NO CPU fallback: without a Blackwell GPU
 * pfann_ctx_create() fails.  Every data pointer may be a host pointer (staged through pinned/device
 * scratch inside the call, results complete on return) or a device pointer of the context's device
 * (work is enqueued on the context's stream, stream-ordered like any CUDA library).
 *
 * Each entry point cites the reference interface it replaces (paths relative to stdio2016/pfann).
 * Handles are thread-compatible, not thread-safe (the reference calls from one Python thread,
 * database.py:178).
 */
#ifndef PFANN_B200_H
#define PFANN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFANN_B200_VERSION 20261017001LL

typedef struct pfann_ctx pfann_ctx;     /* one per (process, GPU): device ordinal, stream, scratch   */
typedef struct pfann_mel pfann_mel;     /* stage 1 plan: windows, twiddles, sparse mel filter bank    */
typedef struct pfann_model pfann_model; /* stage 2 weights in kernel layout                           */
typedef struct pfann_db pfann_db;       /* stage 3 database shard resident in HBM                     */

/* ---- library / context ---------------------------------------------------------------------- */

/* ABI handshake, same idea as `version()` of cpp/seqscore.cpp:27-30 checked at database.py:29-32. */
long long pfann_version(void);
/* Message of the last failing call on this thread. */
const char *pfann_last_error(void);

int pfann_ctx_create(int device, pfann_ctx **out);
void pfann_ctx_destroy(pfann_ctx *ctx);
/* Enqueue all later work on `cuda_stream` (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream). */
int pfann_ctx_set_stream(pfann_ctx *ctx, void *cuda_stream);
int pfann_ctx_sync(pfann_ctx *ctx);
/* Number of kernels of this library launched through the context so far (bench.py: gpu_launches). */
long long pfann_ctx_launches(pfann_ctx *ctx);
int pfann_ctx_sm_count(pfann_ctx *ctx);
/* Optional per-kernel-class timing: when enabled every kernel launch is bracketed by a CUDA-event pair on the
 * launching stream.  pfann_ctx_profile_read synchronises, returns the summed milliseconds and launch counts per
 * class (0 mel, 1 conv tcgen05, 2 conv CUDA-core, 3 LayerNorm, 4 head, 5 kNN scan, 6 kNN select, 7 rerank,
 * 8 misc) and clears the record. */
#define PFANN_N_KERNEL_CLASSES 9
int pfann_ctx_profile(pfann_ctx *ctx, int enable);
int pfann_ctx_profile_read(pfann_ctx *ctx, double *ms, long long *count, int n_classes);
/* Finer breakdown of the record consumed by the LAST pfann_ctx_profile_read: slot i < 16 = convolution i
 * (2*layer + {0: conv1, 1: conv2}; fused conv+LayerNorm kernels included), 16 + i = LayerNorm kernels that follow
 * convolution i, 32 = mel, 33 = head, 34 = layer-0 moments, 35 = kNN sample scan, 36 = kNN filtered scan,
 * 37 = k-th selection of the sample, 38 = kNN select, 39 = rerank. */
#define PFANN_N_PROFILE_DETAIL 48
int pfann_ctx_profile_detail(pfann_ctx *ctx, double *ms, long long *count, int n);

/* ---- stage 1: log-mel front end -------------------------------------------------------------- */

/* Replaces datautil/melspec.py:52-63 build_mel_spec_layer(params) for the default option set
 * (naf_mode=False, mel_log='log', spec_norm='l2'); pfann_mel_create_ex takes the other options.  n_fft must be
 * 1024 (every shipped config). */
int pfann_mel_create(pfann_ctx *ctx, int sample_rate, int n_fft, int hop, double f_min, double f_max,
                     int n_mels, int seg_len, pfann_mel **out);
/* All options build_mel_spec_layer reads from the config (melspec.py:52-63):
 *   naf_mode       != 0: magnitude instead of power, zero instead of reflect padding, slaney mel scale and
 *                  normalisation, + 0.06 instead of + 1e-8                     (melspec.py:27-30,38-41)
 *   mel_log        PFANN_MEL_LOG_NONE / _E / _10                               (melspec.py:43-46)
 *   spec_norm_max  != 0: spec_norm == 'max' -- normalise the waveform by max |x| and subtract the maximum of
 *                  the tile afterwards                                         (melspec.py:35,48-49)        */
#define PFANN_MEL_LOG_NONE 0
#define PFANN_MEL_LOG_E 1
#define PFANN_MEL_LOG_10 2
int pfann_mel_create_ex(pfann_ctx *ctx, int sample_rate, int n_fft, int hop, double f_min, double f_max,
                        int n_mels, int seg_len, int naf_mode, int mel_log, int spec_norm_max, pfann_mel **out);
void pfann_mel_destroy(pfann_mel *mel);
/* Replaces MelSpec.forward (datautil/melspec.py:33-50): x[B][seg_len] fp32 -> out[B][n_mels][T] fp32,
 * T = 1 + seg_len / hop.  L2-normalise, reflect-pad STFT power, HTK mel, log(. + 1e-8). */
int pfann_mel_forward(pfann_mel *mel, const float *x, int64_t B, float *out);
/* Same, but takes the segmenter's job too (datautil/musicdata.py:48,82-88): mono int16 PCM at the model
 * rate; segment b covers pcm[seg_start[b] .. +seg_valid[b]) zero-padded to seg_len, scaled by 1/32768 and
 * mean-removed before the mel.  seg_start/seg_valid are host or device arrays of length B. */
int pfann_mel_forward_pcm16(pfann_mel *mel, const int16_t *pcm, int64_t n_samples, const int64_t *seg_start,
                            const int32_t *seg_valid, int64_t B, float *out);
int pfann_mel_n_frames(pfann_mel *mel);

/* ---- stage 2: fingerprint network ------------------------------------------------------------- */

#define PFANN_PRECISION_FP32 0 /* CUDA-core fp32 path (validation grade)                          */
#define PFANN_PRECISION_BF16 1 /* tcgen05 tensor-core path: bf16 operands, fp32 accumulate + LN   */

/* Replaces FpNetwork(d, h, u, F, T, params) (model.py:132-146).  Supported option set: k=3, stride 2,
 * ReLU, relu_after_bn=True; `fuller` as in params['model'] (model.py:26-29). */
int pfann_model_create(pfann_ctx *ctx, int d, int h, int u, int F, int T, int fuller, pfann_model **out);
/* All options of FpNetwork's params (model.py:132-146): conv_activation (model.py:7-12), relu_after_bn
 * (model.py:58-72: 0 = activation BEFORE each LayerNorm), strides (model.py:82-85: int[8][2] = time stride of
 * conv1, frequency stride of conv2 per layer; NULL = all 2).  Anything but (ReLU, 1, all 2) runs on the CUDA-core
 * fp32 kernels whatever precision pfann_model_finalize is given: the tcgen05 path serves the default option set. */
#define PFANN_ACT_RELU 0
#define PFANN_ACT_ELU 1
int pfann_model_create_ex(pfann_ctx *ctx, int d, int h, int u, int F, int T, int fuller, int conv_activation,
                          int relu_after_bn, const int *strides, pfann_model **out);
void pfann_model_destroy(pfann_model *m);
/* Replaces load_state_dict (builder.py:56): `name` is the reference state_dict key
 * ("f.convs.3.conv1.weight", "f.convs.0.ln2.bias", "g.linear1.weight", ...), `data` fp32 in the
 * reference's own (PyTorch NCHW) element order, host or device. */
int pfann_model_set_param(pfann_model *m, const char *name, const float *data, int64_t numel);
/* Re-layout weights for the kernels; fails if any parameter is missing. */
int pfann_model_finalize(pfann_model *m, int precision);
/* Segments processed per internal pass (workspace is sized for it). */
int pfann_model_set_chunk(pfann_model *m, int chunk);
/* Replaces FpNetwork.forward(x, norm) (model.py:148-153): mel[B][F][T] fp32 -> z[B][d] fp32. */
int pfann_model_forward(pfann_model *m, const float *mel, int64_t B, int norm, float *z);
/* Debug/parity taps: pfann_model_set_tap(layer) before a forward of at most one chunk keeps the activation
 * after SeparableConv2d `layer` (post ln2+ReLU, NCHW fp32, exactly the reference module's output,
 * model.py:73); pfann_model_get_activation copies it out.  layer = -1 disables the tap. */
int pfann_model_set_tap(pfann_model *m, int layer);
int pfann_model_get_activation(pfann_model *m, int layer, float *out, int64_t numel);

/* ---- stages 1+2 fused: the builder / matcher inner loop ---------------------------------------- */

/* Replaces builder.py:88-99 / matcher.py:110-127 for already-framed rows: x[B][seg_len] -> z[B][d]. */
int pfann_extract_segments(pfann_mel *mel, pfann_model *m, const float *x, int64_t B, int norm, float *z);
/* Replaces musicdata.py:82-88 + builder.py:88-99 for a batch of clips stored back to back as mono int16
 * PCM: clip c is pcm[clip_off[c] .. clip_off[c+1]); it yields max(len,seg)-seg)/hop+1 segments, written
 * consecutively to z (song order, like the `embeddings` file builder.py:99).  seg_counts[n_clips] (host,
 * may be NULL) receives the landmarkKey entries (builder.py:101).  clip_off is a HOST array [n_clips+1].
 * z must hold pfann_count_segments(...) rows. */
int pfann_extract_pcm16(pfann_mel *mel, pfann_model *m, const int16_t *pcm, const int64_t *clip_off,
                        int n_clips, int hop_samples, int norm, float *z, int32_t *seg_counts);
int64_t pfann_count_segments(const int64_t *clip_off, int n_clips, int seg_len, int hop_samples);
/* The same for clips that went through the GPU ingest below (fp32 mono samples at the model rate). */
int pfann_extract_f32(pfann_mel *mel, pfann_model *m, const float *wav, const int64_t *clip_off, int n_clips,
                      int hop_samples, int norm, float *z, int32_t *seg_counts);

/* ---- ingest: decoded PCM of any channel count / rate -> mono fp32 at the model rate (DEVICE pointers) ------- */

/* Replaces musicdata.py:44-48: interleaved int16 frames -> planar fp32 [nch][n_frames], scaled by 1/32768. */
int pfann_pcm16_to_planar(pfann_ctx *ctx, const int16_t *pcm, int64_t n_frames, int nch, float *out);
/* Replaces julius.ResampleFrac(old_sr, new_sr) as called at musicdata.py:29,58,66 (third-party, unpinned, absent from
 * the build container: restated from its published algorithm -- windowed sinc, zeros = 24, rolloff = 0.945, replicate
 * padding -- and checked against oracle/pfann_oracle.py only).  x [nch][n] -> y [nch][pfann_resample_len(n, ..)].
 * The reference resamples minute by minute and discards half a second at every seam (musicdata.py:50-66); away from
 * the ends that equals resampling the whole clip at once, which is what this does. */
int64_t pfann_resample_len(int64_t n, int old_sr, int new_sr);
int pfann_resample_frac(pfann_ctx *ctx, const float *x, int nch, int64_t n, int old_sr, int new_sr, float *y);
/* Replaces musicdata.py:72-80: mean over the channels; for two channels whose difference carries more than 1000x the
 * power of their sum (fake stereo with opposite phase) channel 1 is negated first.  x [nch][n] -> mono [n]. */
int pfann_mix_mono(pfann_ctx *ctx, const float *x, int nch, int64_t n, float *mono);

/* ---- training-step pieces (train.py, datautil/specaug.py, datautil/noise.py; DEVICE pointers) -------------- */

/* Replaces similarity_loss(y, tau) of train.py:41-52 (NT-Xent; rows 2i, 2i+1 are positive pairs) together with its
 * backward: loss (one float) and, if dy != NULL, dL/dy [N][d].  fp32 throughout: logits are divided by tau = 0.05,
 * bf16 products would be off by percents after exp(). */
int pfann_ntxent(pfann_ctx *ctx, const float *y, int N, int d, float tau, float *loss, float *dy);
/* Replaces SpecAugment.augment (datautil/specaug.py:40-42) for a batch: x[B][F][T] *= 1 - mask_b, mask_b = cutout
 * rectangle + frequency band + time band given as rects[b] = {f0, f1, t0, t1, fb0, fb1, tb0, tb1} (half-open); the
 * random draws stay on the host in the reference's order (specaug.py:13-38) so that seeds reproduce. */
int pfann_specaug_apply(pfann_ctx *ctx, float *x, const int32_t *rects, int64_t B, int F, int T);
/* Replaces NoiseData.add_noises arithmetic (datautil/noise.py:96-109): out[b] = x[b] + ratio_b * noise[b],
 * ratio_b = rms(x[b]) / rms(noise[b]) * 10^(-snr_db[b] / 20), rms clamped at sqrt(1e-12). */
int pfann_snr_mix(pfann_ctx *ctx, const float *x, const float *noise, const float *snr_db, int64_t B, int n, float *out);
/* Replaces the impulse-response augmentation of MusicSegmentDataset.__getitem__ (datautil/dataset_v2.py:157-163:
 * irfft(rfft(x, n) * H_room * H_mic, n)[pad_start:segment_size], n >= len(x) + len(h), i.e. a causal LINEAR convolution)
 * by the direct form: out[b][i] = sum_{k<L} h[b][k] x[b][out_start + i - k], x = 0 outside [0, n).  x [B][n], h [B][L]
 * (the room and microphone responses of row b, already convolved with each other or applied in two calls),
 * out [B][out_len].  fp32; differs from the fp32 FFT route by rounding only (test tolerance 2e-5 of the row's peak). */
int pfann_ir_conv(pfann_ctx *ctx, const float *x, int64_t B, int n, const float *h, int L, float *out, int out_start,
                  int out_len);

/* Replaces the forward/backward of `y = model(mel(x))` ... `loss.backward()` in the training loop (train.py:96-103,
 * FpNetwork.forward model.py:148-153 under torch autograd), fp32.  train_forward computes z[B][d] from mel[B][F][T]
 * like pfann_model_forward (norm: 0 skips the final L2 normalisation) and keeps every convolution's raw output,
 * LayerNorm statistics and activations; train_backward takes dz = dL/dz [B][d] of the same batch and leaves the
 * parameter gradients in the model; get_grad copies the gradient of the parameter with state_dict key `name`, in the
 * reference's element order (host or device `out`).  Gradients are sums over the batch like torch's; they are
 * overwritten, not accumulated, by the next backward.  Every option set of pfann_model_create_ex is supported. */
int pfann_model_train_forward(pfann_model *m, const float *mel, int64_t B, int norm, float *z);
int pfann_model_train_backward(pfann_model *m, const float *dz, int norm);
int pfann_model_get_grad(pfann_model *m, const char *name, float *out, int64_t numel);
/* optimizer.step() (train.py:103) changed the parameters: refresh one of them from DEVICE memory (the reference's
 * element order) without the host round trip of pfann_model_set_param + pfann_model_finalize.  The model must have
 * been finalized with PFANN_PRECISION_FP32 once; it then serves the training entry points and the fp32 forward. */
int pfann_model_train_load_param(pfann_model *m, const char *name, const float *data, int64_t numel);

/* ---- stage 3: database search + sequence score ------------------------------------------------- */

/* Replaces Database.__init__ (database.py:74-99) for a Flat inner-product index: `emb` is the
 * `embeddings` file (builder.py:99,122: fp32 [n][d], C order), `landmark_key` the `landmarkKey` file
 * (int32 [n_songs], builder.py:138-139).  The rows are copied to HBM (fp32 for exact rescoring + bf16 for
 * the tensor-core scan).  `id_base`/`song_base` are the global row / song index of this shard's first
 * row / song when the database is row-sharded across GPUs at song boundaries (0 for a single GPU). */
int pfann_db_open(pfann_ctx *ctx, const float *emb, int64_t n, int d, const int32_t *landmark_key,
                  int n_songs, int64_t id_base, int64_t song_base, pfann_db **out);
void pfann_db_close(pfann_db *db);
int64_t pfann_db_ntotal(pfann_db *db);
/* Testing hook: candidate slots per query (power of two <= 4096), rows of the threshold pre-pass, and
 * whether the scan runs on tensor cores (1) or fp32 CUDA cores (0); <= 0 / < 0 keeps the current value. */
int pfann_db_set_tuning(pfann_db *db, int cand_cap, int sample_rows, int use_tc);

/* Replaces faiss IndexFlatIP.search as called at database.py:121,172: q[Q][d] fp32 -> dist[Q][k] fp32
 * descending, labels[Q][k] int64 (global row ids), padded with -FLT_MAX / -1 when fewer than k rows.
 * Scores are exact fp32 inner products (k-sequential fused multiply-add); ties -> lower id first. */
int pfann_db_search(pfann_db *db, const float *q, int64_t Q, int k, float *dist, int64_t *labels);

/* Replaces seq_score() of cpp/seqscore.cpp:32-43 with the faiss::Index* swapped for our handle; the
 * remaining arguments, buffer ownership (song_scores[n_songs][2] in/out, caller zero-initialised,
 * database.py:176), and the return value (best song id or -1) are identical.  song_pos is GLOBAL
 * (int64 [n_songs+1]) and labels are global row ids; only candidates whose song lives in this shard are
 * scored. */
int pfann_db_seq_score(pfann_db *db, const int64_t *song_pos, int n_songs, const float *query,
                       int query_len, const int64_t *labels, int top_k, float *song_scores,
                       int frame_shift_mul, float score_alpha);

/* Batched form of Database.query_embeddings (database.py:111-115,168-195) for many query files per
 * call (the matchemb.py split, matchemb.py:59-80): queries[sum len][d] fp32, query_index[nq][2] int64
 * (start, len) as in the `query_index` file (extractemb.py:85).  Outputs per query file:
 * best_score[nq] fp32, best_song[nq] int32 (global id, -1 if none), best_time[nq] fp32 in FRAMES
 * (t*fsm - shift, seqscore.cpp:113; seconds = frames*hop_size/fsm, database.py:191).
 * If song_scores != NULL it is [nq][n_songs_total][2] fp32, zero-filled then raised like seq_score. */
int pfann_db_query(pfann_db *db, const float *queries, const int64_t *query_index, int nq, int top_k,
                   int frame_shift_mul, float score_alpha, float *best_score, int32_t *best_song,
                   float *best_time, float *song_scores, int64_t n_songs_total);

/* Multi-GPU pieces of the same path: local top-k (global ids) for an all-gather, then rerank given merged
 * labels.  pfann_topk_merge merges G gathered lists [G][Q][k] into [Q][k] (score desc, id asc). */
int pfann_topk_merge(pfann_ctx *ctx, const float *dist_g, const int64_t *labels_g, int G, int64_t Q, int k,
                     float *dist, int64_t *labels);
int pfann_db_rerank(pfann_db *db, const float *queries, const int64_t *query_index, int nq,
                    const int64_t *labels, int top_k, int frame_shift_mul, float score_alpha,
                    float *best_score, int32_t *best_song, float *best_time);

/* Sharded search, tighter variant of the threshold phase: besides this shard's own threshold, the k best SAMPLED scan
 * scores per query as sortable 32-bit keys (0 = none), topk [Q][k].  The ranks all-gather these lists (Q * k * 4 bytes
 * each) and call pfann_db_thresholds_from_topk on the gathered [world][Q][k] array: thr[q] = the k-th best of the UNION
 * of the samples minus the scan's error bound -- a lower bound of the global k-th best score that is several times
 * tighter than the maximum of the per-shard k-th bests, so fewer rows survive the filtered scan. */
int pfann_db_search_sample_topk(pfann_db *db, const float *q, int64_t Q, int k, float *thr, uint32_t *topk);
int pfann_db_thresholds_from_topk(pfann_db *db, const uint32_t *gathered, int world, const float *q, int64_t Q, int k,
                                  float *thr);
/* The same exchange for many GPUs without a host round trip per batch (pfann_b200/dist.py).  All data pointers
 * are DEVICE pointers and every call is stream-ordered; nothing is read back.
 *   1. pfann_db_search_thresholds: thr[Q] = a lower bound of this shard's k-th best score per query (sample
 *      pre-pass).  The caller takes the element-wise MAXIMUM over the shards (one all-reduce of Q floats): still a
 *      lower bound of the GLOBAL k-th best, provided all shards use one error bound (pfann_db_set_max_norm with the
 *      maximum of pfann_db_max_norm over the shards).
 *   2. pfann_db_search_filtered: exact top-k of the rows that reach thr, as sortable 64-bit keys
 *      (order-preserving fp32 score bits << 32 | 0xFFFFFFFF - global row id; 0 = empty) -> ONE all-gather of
 *      [Q][k] keys per batch, as BASELINE.json's north_star asks.  A candidate list that overflows its 4096
 *      slots is redone with tighter thresholds: immediately (one read-back of a flag per call) or, with
 *      defer_overflow_check, counted for pfann_db_take_overflow so that a stream of batches never waits for the host.
 *   3. pfann_topk_merge_keys: [G][Q][k] gathered keys -> global top-k labels (and distances) per query.
 *   4. pfann_db_rerank_packed: sequence score over the candidates whose songs this shard owns -> [nq][4] fp32
 *      (score, song id bits, time in frames, 0); one all-gather of these tiny records, then
 *   5. pfann_best_combine: [G][nq][4] -> the winner per query file (score desc, lower song id, zero floor of
 *      database.py:176,190), same record layout. */
int pfann_db_search_thresholds(pfann_db *db, const float *q, int64_t Q, int k, float *thr);
int pfann_db_search_filtered(pfann_db *db, const float *q, int64_t Q, int k, float *thr, uint64_t *keys,
                             int defer_overflow_check);
/* Candidate lists that overflowed in calls made with defer_overflow_check = 1 since the last call of this function
 * (read after synchronising the stream); non-zero: repeat those calls with defer_overflow_check = 0. */
int pfann_db_take_overflow(pfann_db *db);
int pfann_topk_merge_keys(pfann_ctx *ctx, const uint64_t *keys_g, int G, int64_t Q, int k, float *dist,
                          int64_t *labels);
int pfann_db_rerank_packed(pfann_db *db, const float *queries, const int64_t *query_index, int nq, int max_len,
                           const int64_t *labels, int top_k, int frame_shift_mul, float score_alpha, float *packed);
int pfann_best_combine(pfann_ctx *ctx, const float *packed_g, int G, int nq, float *packed_out);
float pfann_db_max_norm(pfann_db *db);
/* Rows of the threshold pre-pass = max(floor, scale * k * n / 256): with G shards whose thresholds are max-reduced the
 * union of the per-shard samples does the work, so each shard can sample less (dist.py uses 0.5 from 4 shards on). */
int pfann_db_set_sample_scale(pfann_db *db, float scale);
int pfann_db_set_max_norm(pfann_db *db, float max_norm);

/* ---- reference-compatible symbols ------------------------------------------------------------- */

/* Bit-compatible with cpp/seqscore.cpp:27-43 so that database.py:15-32 can load this library in place of
 * cpp/seqscore: version() returns 20220625002; seq_score()'s first argument is a pfann_db* instead of a
 * faiss::Index*. */
long long version(void);
int seq_score(void *index, const int64_t *song_pos, int n_songs, const float *query, int query_len,
              const int64_t *labels, int top_k, float *song_scores, int frame_shift_mul, float score_alpha);

#ifdef __cplusplus
}
#endif
#endif /* PFANN_B200_H */
