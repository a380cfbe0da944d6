// seqscore.cu -- stage 3b: the diagonal sequence-score vote on the GPU, one CTA per query file.
//
// Replaces cpp/seqscore.cpp:33-136 (and with it the Python loop database.py:129-163 that is 60-80 % of the
// reference's query time, thesis Table A.3).  Same algorithm, same arithmetic:
//   candidates = sorted unique (song, label - song_pos[song] - t/fsm, t%fsm)      seqscore.cpp:49-60
//   score      = mean over the sub-query of <db row, query row>, rows outside the song skipped; each inner
//                product accumulated k-sequentially with separate fp32 multiply and add (:99-102), the
//                per-row products summed j-sequentially (:107), then one fp32 division (:111)
//   per song   = first strict maximum in candidate order (:126-132); best song = highest score, ties ->
//                lower song id (:115-124)
// so (song id, offset) are bit-exact against the reference build (oracle/_ref) and so are the scores for
// score_alpha == 0 (expf differs in the last ulp between libm and CUDA for score_alpha > 0).
// Sharding: labels and song ids are GLOBAL; a shard scores only candidates whose song it owns.
#include <float.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "db.cuh"
#include "pfann_b200.h"

using namespace pfann;

namespace pfann {
int db_search_dev(Db *db, const float *q, int64_t Q, int k, float *dist, int64_t *labels);
}

namespace {

constexpr unsigned long long SENTINEL = ~0ull;
constexpr int OFF_BIAS = 1 << 27;
constexpr int RR_THREADS = 256;
constexpr int SMEM_CAND_MAX = 8192;  // candidates sorted in shared memory; larger lists use global scratch
constexpr int RR_DH = 64;            // columns staged per step of the diagonal dot products

__device__ __forceinline__ unsigned long long pack_key(int64_t song, int off, int shift) {
    return ((unsigned long long)song << 36) | ((unsigned long long)(unsigned)(off + OFF_BIAS) << 8) |
           (unsigned long long)shift;
}
__device__ __forceinline__ int key_song(unsigned long long k) { return (int)(k >> 36); }
__device__ __forceinline__ int key_off(unsigned long long k) { return (int)((k >> 8) & 0xFFFFFFFull) - OFF_BIAS; }
__device__ __forceinline__ int key_shift(unsigned long long k) { return (int)(k & 0xFF); }

struct RerankArgs {
    const float *emb;         // [n][d] fp32
    const int64_t *song_pos;  // [n_songs+1] global start rows of local songs
    int n_songs;
    int64_t id_base, song_base;
    int d;
    const float *queries;        // [sum len][d]
    const int64_t *query_index;  // [nq][2] (start, len)
    const int64_t *labels;       // [sum len][k]
    int k, fsm;
    float alpha;
    int capK;                    // power of two >= max len * k
    unsigned long long *gkeys;   // [nq][capK] or nullptr when shared memory is used
    float *gscores;              // [nq][capK] or nullptr
    unsigned stage_off;          // byte offset of the row staging tiles in dynamic shared memory, 0 = none
    // few query files per launch (matcher.py:136 calls once per file): gridDim.y CTAs share one file's candidates, the
    // last one to finish gathers the scores and runs the per-song reduction
    float *xscores;              // [nq][capK] exchange of candidate scores, or nullptr (one CTA per file)
    unsigned *done_cnt;          // [nq] CTAs finished, zero on entry
    // outputs
    float *best_score;           // [nq] raw best candidate score (-inf if none)
    int *best_song;              // [nq] global song id or -1
    float *best_time;            // [nq] t*fsm - shift of the best song's winning candidate
    int *sp_song;                // optional sparse per-song results [nq][capK]: song id or -1
    float *sp_score, *sp_time;
};

__global__ void __launch_bounds__(RR_THREADS) rerank_kernel(const RerankArgs a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ int n_valid_s, last_s;
    __shared__ float red_s[RR_THREADS / 32];
    __shared__ int red_song[RR_THREADS / 32];
    __shared__ float red_t[RR_THREADS / 32];
    const int qi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t start = a.query_index[2 * qi];
    const int L = (int)a.query_index[2 * qi + 1];
    const int nlab = L * a.k;
    int P = 1;
    while (P < nlab) P <<= 1;
    unsigned long long *keys;
    float *scores;
    if (a.gkeys) {
        keys = a.gkeys + (int64_t)qi * a.capK;
        scores = a.gscores + (int64_t)qi * a.capK;
    } else {
        keys = reinterpret_cast<unsigned long long *>(smraw);
        scores = reinterpret_cast<float *>(keys + a.capK);
    }
    // per-warp staging tile [32 rows][RR_DH + 1] behind the candidate arrays (d a multiple of RR_DH only)
    float *stage = nullptr;
    if (a.stage_off != 0) stage = reinterpret_cast<float *>(smraw + a.stage_off) + warp * (32 * (RR_DH + 1));
    if (tid == 0) n_valid_s = 0;
    __syncthreads();
    // (1) candidate keys (seqscore.cpp:49-58)
    const int64_t lo_row = a.song_pos[0], hi_row = a.song_pos[a.n_songs];
    int mine = 0;
    for (int i = tid; i < P; i += RR_THREADS) {
        unsigned long long key = SENTINEL;
        if (i < nlab) {
            const int t = i / a.k;
            const int64_t lab = a.labels[(start + t) * a.k + (i - t * a.k)];
            if (lab >= 0 && lab >= lo_row && lab < hi_row) {
                // idx_to_song_id (seqscore.cpp:23-25): upper_bound over the song starts, minus one
                int lo = 0, hi = a.n_songs;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (a.song_pos[mid] <= lab) lo = mid + 1; else hi = mid;
                }
                const int song = lo - 1;
                const int off = (int)(lab - a.song_pos[song]) - t / a.fsm;
                key = pack_key(a.song_base + song, off, t % a.fsm);
                mine++;
            }
        }
        keys[i] = key;
    }
    if (mine) atomicAdd(&n_valid_s, mine);
    // (2) sort ascending; duplicates become adjacent (seqscore.cpp:59-60)
    bitonic_sort<unsigned long long, false>(keys, P);
    const int n_valid = n_valid_s;
    // (3) score every unique candidate: one warp per candidate, one lane per sub-query row
    const int nsplit = (int)gridDim.y, split = (int)blockIdx.y;
    float *xs = a.xscores ? a.xscores + (int64_t)qi * a.capK : nullptr;
    for (int i = warp + (RR_THREADS / 32) * split; i < n_valid; i += (RR_THREADS / 32) * nsplit) {
        const unsigned long long key = keys[i];
        const bool head = (i == 0) || (keys[i - 1] != key);
        if (!head) {
            if (lane == 0) {
                scores[i] = -INFINITY;
                if (xs) xs[i] = -INFINITY;
            }
            continue;
        }
        const int song = key_song(key) - (int)a.song_base, off = key_off(key), shift = key_shift(key);
        const int64_t song_start = a.song_pos[song];
        const int song_len = (int)(a.song_pos[song + 1] - song_start);
        const int my_len = (L - shift + a.fsm - 1) / a.fsm;
        float sco = 0.f;
        for (int jb = 0; jb < my_len; jb += 32) {
            const int j = jb + lane;
            const bool ok = (j < my_len) && (off + j >= 0) && (off + j < song_len);
            float ip = 0.f;
            const float *vec = a.emb + (song_start - a.id_base + off + j) * a.d;
            const float *qv = a.queries + (start + (int64_t)j * a.fsm + shift) * a.d;
            if (stage != nullptr) {
                // The rows of a diagonal are consecutive database rows: the warp fetches them with coalesced 16-byte
                // loads (16 independent requests in flight per lane) into a padded shared-memory tile, 64 columns
                // at a time, and every lane then runs the reference's k-sequential multiply/add chain on its own
                // row from shared memory.  (One scalar global load per multiply left the kernel waiting on DRAM.)
                const unsigned okmask = __ballot_sync(0xffffffffu, ok);
                const float *base = a.emb + (song_start - a.id_base + off + jb) * a.d;  // row jb of this block
                for (int h0 = 0; h0 < a.d; h0 += RR_DH) {
                    float4 ld[16];
#pragma unroll
                    for (int r2 = 0; r2 < 16; r2++) {
                        const int row = 2 * r2 + (lane >> 4);
                        ld[r2] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if ((okmask >> row) & 1u)
                            ld[r2] = __ldg(reinterpret_cast<const float4 *>(base + (int64_t)row * a.d + h0) + (lane & 15));
                    }
#pragma unroll
                    for (int r2 = 0; r2 < 16; r2++) {
                        float *dst = stage + (2 * r2 + (lane >> 4)) * (RR_DH + 1) + (lane & 15) * 4;
                        dst[0] = ld[r2].x; dst[1] = ld[r2].y; dst[2] = ld[r2].z; dst[3] = ld[r2].w;
                    }
                    __syncwarp();
                    if (ok) {
                        const float *sv = stage + lane * (RR_DH + 1);
#pragma unroll
                        for (int kk0 = 0; kk0 < RR_DH; kk0 += 8) {
                            const float4 qa = __ldg(reinterpret_cast<const float4 *>(qv + h0 + kk0));
                            const float4 qb = __ldg(reinterpret_cast<const float4 *>(qv + h0 + kk0 + 4));
                            ip = __fadd_rn(ip, __fmul_rn(sv[kk0], qa.x));
                            ip = __fadd_rn(ip, __fmul_rn(sv[kk0 + 1], qa.y));
                            ip = __fadd_rn(ip, __fmul_rn(sv[kk0 + 2], qa.z));
                            ip = __fadd_rn(ip, __fmul_rn(sv[kk0 + 3], qa.w));
                            ip = __fadd_rn(ip, __fmul_rn(sv[kk0 + 4], qb.x));
                            ip = __fadd_rn(ip, __fmul_rn(sv[kk0 + 5], qb.y));
                            ip = __fadd_rn(ip, __fmul_rn(sv[kk0 + 6], qb.z));
                            ip = __fadd_rn(ip, __fmul_rn(sv[kk0 + 7], qb.w));
                        }
                    }
                    __syncwarp();
                }
            } else if (ok) {
                for (int kk = 0; kk < a.d; kk++) ip = __fadd_rn(ip, __fmul_rn(vec[kk], qv[kk]));
            }
            for (int l = 0; l < 32; l++) {
                const float v = __shfl_sync(0xffffffffu, ip, l);
                const int okl = __shfl_sync(0xffffffffu, (int)ok, l);
                if (okl) {
                    if (a.alpha == 0.0f) {
                        sco = __fadd_rn(sco, v);
                    } else if (a.alpha > 0.0f) {
                        const float l2 = 1.0f - 1.0f * v;
                        sco = __fadd_rn(sco, expf(-a.alpha * l2 * l2));
                    }
                }
            }
        }
        sco = __fdiv_rn(sco, (float)(my_len > 1 ? my_len : 1));
        if (lane == 0) {
            scores[i] = sco;
            if (xs) xs[i] = sco;
        }
    }
    __syncthreads();
    if (xs != nullptr) {
        // the last of this file's CTAs to arrive continues with everybody's scores; the others are done
        if (tid == 0) {
            __threadfence();
            const unsigned prev = atomicAdd(a.done_cnt + qi, 1u);
            last_s = (prev == (unsigned)nsplit - 1u);
            if (last_s) a.done_cnt[qi] = 0u;   // ready for the next launch
        }
        __syncthreads();
        if (!last_s) return;
        __threadfence();
        for (int i = tid; i < n_valid; i += RR_THREADS) scores[i] = __ldcg(xs + i);
        __syncthreads();
    }
    // (4) per-song first strict maximum, then the best song (ties -> lower song id)
    float my_best = -INFINITY, my_t = 0.f;
    int my_song = -1;
    for (int i = tid; i < P; i += RR_THREADS) {
        int out_song = -1;
        float out_score = 0.f, out_t = 0.f;
        if (i < n_valid) {
            const unsigned long long key = keys[i];
            const int song = key_song(key);
            if (i == 0 || key_song(keys[i - 1]) != song) {  // first candidate of this song
                float cur = -INFINITY, cur_t = 0.f;
                bool any = false;
                for (int j = i; j < n_valid && key_song(keys[j]) == song; j++) {
                    if (j > i && keys[j] == keys[j - 1]) continue;  // duplicate
                    const float s = scores[j];
                    if (!any || s > cur) {
                        cur = s;
                        cur_t = (float)(key_off(keys[j]) * a.fsm - key_shift(keys[j]));
                        any = true;
                    }
                }
                out_song = song; out_score = cur; out_t = cur_t;
                if (cur > my_best || (cur == my_best && song < my_song) || my_song < 0) {
                    my_best = cur; my_song = song; my_t = cur_t;
                }
            }
        }
        if (a.sp_song && i < a.capK) {
            a.sp_song[(int64_t)qi * a.capK + i] = out_song;
            a.sp_score[(int64_t)qi * a.capK + i] = out_score;
            a.sp_time[(int64_t)qi * a.capK + i] = out_t;
        }
    }
    if (a.sp_song)
        for (int i = P + tid; i < a.capK; i += RR_THREADS) a.sp_song[(int64_t)qi * a.capK + i] = -1;
    // block arg-max: (score desc, song asc); song -1 = nothing
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, my_best, o);
        const int og = __shfl_xor_sync(0xffffffffu, my_song, o);
        const float ot = __shfl_xor_sync(0xffffffffu, my_t, o);
        if (og >= 0 && (my_song < 0 || os > my_best || (os == my_best && og < my_song))) {
            my_best = os; my_song = og; my_t = ot;
        }
    }
    if (lane == 0) { red_s[warp] = my_best; red_song[warp] = my_song; red_t[warp] = my_t; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < RR_THREADS / 32; w++) {
            const float os = red_s[w];
            const int og = red_song[w];
            if (og >= 0 && (my_song < 0 || os > my_best || (os == my_best && og < my_song))) {
                my_best = os; my_song = og; my_t = red_t[w];
            }
        }
        a.best_score[qi] = my_song >= 0 ? my_best : -INFINITY;
        a.best_song[qi] = my_song;
        a.best_time[qi] = my_song >= 0 ? my_t : 0.f;
    }
}

// all-device rerank; max_len = longest query (host knows it from query_index)
int rerank_dev(Db *db, const float *queries, const int64_t *query_index, int nq, int max_len, const int64_t *labels,
               int top_k, int fsm, float alpha, float *best_score, int *best_song, float *best_time, int *sp_song,
               float *sp_score, float *sp_time, int *capK_out) {
    int capK = 1;
    while (capK < max_len * top_k) capK <<= 1;
    if (capK_out) *capK_out = capK;
    if (nq == 0) return PFANN_OK;
    RerankArgs a = {};
    a.emb = db->emb32; a.song_pos = db->song_pos; a.n_songs = db->n_songs;
    a.id_base = db->id_base; a.song_base = db->song_base; a.d = db->d;
    a.queries = queries; a.query_index = query_index; a.labels = labels;
    a.k = top_k; a.fsm = fsm; a.alpha = alpha; a.capK = capK;
    a.best_score = best_score; a.best_song = best_song; a.best_time = best_time;
    a.sp_song = sp_song; a.sp_score = sp_score; a.sp_time = sp_time;
    size_t smem = 0;
    if (capK <= SMEM_CAND_MAX) {
        smem = (size_t)capK * 12;
    } else {
        PF_TRY(db->rr_keys.ensure((size_t)nq * capK * 8));
        PF_TRY(db->rr_scores.ensure((size_t)nq * capK * 4));
        a.gkeys = db->rr_keys.as<unsigned long long>();
        a.gscores = db->rr_scores.as<float>();
    }
    const size_t stage_bytes = (size_t)(RR_THREADS / 32) * 32 * (RR_DH + 1) * sizeof(float);
    if (db->d % RR_DH == 0) {  // staged, coalesced diagonal fetch; dynamic shared memory starts 16-byte aligned
        a.stage_off = (unsigned)((smem + 15) & ~(size_t)15);
        if (a.stage_off == 0) a.stage_off = 16;
        smem = a.stage_off + stage_bytes;
    }
    PF_CUDA(cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(SMEM_CAND_MAX * 12 + 16 + stage_bytes)));
    int nsplit = 1;
    if (a.gkeys == nullptr && nq * 4 <= db->ctx->sm_count) {   // a handful of files: spread each over several CTAs
        nsplit = db->ctx->sm_count / nq;
        if (nsplit > 16) nsplit = 16;
        const bool fresh = db->rr_done.cap < (size_t)nq * 4;
        PF_TRY(db->rr_xscores.ensure((size_t)nq * capK * 4));
        PF_TRY(db->rr_done.ensure((size_t)nq * 4 < 1024 ? 1024 : (size_t)nq * 4));
        if (fresh) PF_CUDA(cudaMemsetAsync(db->rr_done.p, 0, db->rr_done.cap, db->ctx->stream));   // self-resetting afterwards
        a.xscores = db->rr_xscores.as<float>();
        a.done_cnt = db->rr_done.as<unsigned>();
    }
    ProfScope ps(db->ctx, K_RERANK, 39);
    rerank_kernel<<<dim3(nq, nsplit), RR_THREADS, smem, db->ctx->stream>>>(a);
    db->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

__global__ void pack_best_kernel(const float *score, const int *song, const float *time, int nq, float4 *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) out[i] = make_float4(score[i], __int_as_float(song[i]), time[i], 0.f);
}

int check_rerank_args(Db *db, int top_k, int fsm) {
    PF_CHECK(top_k > 0, PFANN_ERR_ARG, "rerank: top_k must be positive");
    PF_CHECK(fsm >= 1 && fsm <= 255, PFANN_ERR_ARG, "rerank: frame_shift_mul must be in 1..255");
    PF_CHECK(db->song_base + db->n_songs < (1LL << 28), PFANN_ERR_UNSUPPORTED, "rerank: more than 2^28 songs");
    return PFANN_OK;
}

// post-processing shared by the host-facing wrappers: the reference reads the answer back out of the
// zero-initialised song_scores table (database.py:176,190-191), so a best score <= 0 reports (0, 0).
void apply_zero_floor(int nq, float *score, const int *song, float *time) {
    for (int i = 0; i < nq; i++) {
        if (song[i] < 0 || !(score[i] > 0.f)) {
            score[i] = 0.f;
            time[i] = 0.f;
        }
    }
}

}  // namespace

extern "C" {

int pfann_db_rerank(pfann_db *h, const float *queries, const int64_t *query_index, int nq, const int64_t *labels,
                    int top_k, int fsm, float alpha, float *best_score, int32_t *best_song, float *best_time) {
    PF_CHECK(h && nq >= 0, PFANN_ERR_ARG, "pfann_db_rerank: bad argument");
    Db *db = reinterpret_cast<Db *>(h);
    PF_TRY(check_rerank_args(db, top_k, fsm));
    if (nq == 0) return PFANN_OK;
    PF_CHECK(queries && query_index && labels && best_score && best_song && best_time, PFANN_ERR_ARG,
             "pfann_db_rerank: NULL argument");
    PF_CUDA(cudaSetDevice(db->ctx->device));
    // query_index must be readable on the host to size the launch
    std::vector<int64_t> qi_host((size_t)nq * 2);
    PF_CUDA(cudaMemcpyAsync(qi_host.data(), query_index, sizeof(int64_t) * nq * 2, cudaMemcpyDefault, db->ctx->stream));
    PF_CUDA(cudaStreamSynchronize(db->ctx->stream));
    int64_t total = 0;
    int max_len = 0;
    for (int i = 0; i < nq; i++) {
        const int64_t s = qi_host[2 * i], l = qi_host[2 * i + 1];
        PF_CHECK(s >= 0 && l >= 0 && l < (1 << 20), PFANN_ERR_ARG, "pfann_db_rerank: bad query_index row %d", i);
        if (s + l > total) total = s + l;
        if (l > max_len) max_len = (int)l;
    }
    const void *qd, *qid, *ld;
    void *bs, *bg, *bt;
    PF_TRY(stage_input(db->ctx, 0, queries, (size_t)total * db->d * 4, &qd));
    PF_TRY(stage_input(db->ctx, 1, query_index, (size_t)nq * 16, &qid));
    PF_TRY(stage_input(db->ctx, 2, labels, (size_t)total * top_k * 8, &ld));
    PF_TRY(stage_output(db->ctx, 0, best_score, (size_t)nq * 4, &bs));
    PF_TRY(stage_output(db->ctx, 1, best_song, (size_t)nq * 4, &bg));
    PF_TRY(stage_output(db->ctx, 2, best_time, (size_t)nq * 4, &bt));
    PF_TRY(rerank_dev(db, (const float *)qd, (const int64_t *)qid, nq, max_len, (const int64_t *)ld, top_k, fsm, alpha,
                      (float *)bs, (int *)bg, (float *)bt, nullptr, nullptr, nullptr, nullptr));
    PF_TRY(finish_output(db->ctx, 0, best_score, (size_t)nq * 4));
    PF_TRY(finish_output(db->ctx, 1, best_song, (size_t)nq * 4));
    return finish_output(db->ctx, 2, best_time, (size_t)nq * 4);
}

/* all-device form for the sharded search: nothing is read back, the winners come out as one [nq][4] fp32 record
 * (score, song id bits, time in frames, 0) ready for a single all-gather */
int pfann_db_rerank_packed(pfann_db *h, const float *queries, const int64_t *query_index, int nq, int max_len,
                           const int64_t *labels, int top_k, int fsm, float alpha, float *packed) {
    PF_CHECK(h && nq >= 0 && max_len >= 0 && max_len < (1 << 20), PFANN_ERR_ARG, "pfann_db_rerank_packed: bad argument");
    Db *db = reinterpret_cast<Db *>(h);
    PF_TRY(check_rerank_args(db, top_k, fsm));
    if (nq == 0) return PFANN_OK;
    PF_CHECK(queries && query_index && labels && packed, PFANN_ERR_ARG, "pfann_db_rerank_packed: NULL argument");
    PF_CHECK(is_device_ptr(queries) && is_device_ptr(query_index) && is_device_ptr(labels) && is_device_ptr(packed),
             PFANN_ERR_ARG, "pfann_db_rerank_packed: device pointers only");
    PF_CUDA(cudaSetDevice(db->ctx->device));
    PF_TRY(db->rr_out.ensure((size_t)nq * 12));
    float *bs = db->rr_out.as<float>();
    int *bg = reinterpret_cast<int *>(bs + nq);
    float *bt = reinterpret_cast<float *>(bg + nq);
    PF_TRY(rerank_dev(db, queries, query_index, nq, max_len, labels, top_k, fsm, alpha, bs, bg, bt, nullptr, nullptr,
                      nullptr, nullptr));
    pack_best_kernel<<<cdiv(nq, 256), 256, 0, db->ctx->stream>>>(bs, bg, bt, nq, reinterpret_cast<float4 *>(packed));
    db->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

int pfann_db_seq_score(pfann_db *h, const int64_t *song_pos, int n_songs, const float *query, int query_len,
                       const int64_t *labels, int top_k, float *song_scores, int fsm, float alpha) {
    // error convention of the reference: the return value IS the song id, -1 = no candidate; failures
    // also return -1 with pfann_last_error() set (no exceptions, seqscore.cpp has no error codes either)
    if (!h || !query || !labels || !song_scores || query_len < 0) {
        set_error("pfann_db_seq_score: bad argument");
        return -1;
    }
    Db *db = reinterpret_cast<Db *>(h);
    if (check_rerank_args(db, top_k, fsm) < 0) return -1;
    if (song_pos && n_songs < db->song_base + db->n_songs) {
        set_error("pfann_db_seq_score: caller passes %d songs but this shard ends at song %lld", n_songs,
                  (long long)(db->song_base + db->n_songs));
        return -1;
    }
    if (query_len == 0) return -1;
    if (cudaSetDevice(db->ctx->device) != cudaSuccess) return -1;
    Ctx *ctx = db->ctx;
    const int64_t qi_host[2] = {0, query_len};
    int capK = 1;
    while (capK < query_len * top_k) capK <<= 1;
    const void *qd, *ld;
    if (stage_input(ctx, 0, query, (size_t)query_len * db->d * 4, &qd) < 0) return -1;
    if (stage_input(ctx, 2, labels, (size_t)query_len * top_k * 8, &ld) < 0) return -1;
    // outputs: [score, song, time] + sparse lists in one scratch buffer
    const size_t need = 64 + 16 + (size_t)capK * 12;
    if (db->rr_out.ensure(need) < 0) return -1;
    unsigned char *base = db->rr_out.as<unsigned char>();
    float *bs = reinterpret_cast<float *>(base);
    int *bg = reinterpret_cast<int *>(base + 16);
    float *bt = reinterpret_cast<float *>(base + 32);
    int64_t *qid = reinterpret_cast<int64_t *>(base + 64);
    int *sps = reinterpret_cast<int *>(base + 80);
    float *spc = reinterpret_cast<float *>(base + 80 + (size_t)capK * 4);
    float *spt = reinterpret_cast<float *>(base + 80 + (size_t)capK * 8);
    if (cudaMemcpyAsync(qid, qi_host, 16, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return -1;
    if (rerank_dev(db, (const float *)qd, qid, 1, query_len, (const int64_t *)ld, top_k, fsm, alpha, bs, bg, bt, sps,
                   spc, spt, nullptr) < 0)
        return -1;
    std::vector<unsigned char> hostbuf(need);
    if (cudaMemcpyAsync(hostbuf.data(), base, need, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        set_error("pfann_db_seq_score: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    const int *hs = reinterpret_cast<const int *>(hostbuf.data() + 80);
    const float *hc = reinterpret_cast<const float *>(hostbuf.data() + 80 + (size_t)capK * 4);
    const float *ht = reinterpret_cast<const float *>(hostbuf.data() + 80 + (size_t)capK * 8);
    for (int i = 0; i < capK; i++) {  // seqscore.cpp:126-132: only ever raises entries
        const int s = hs[i];
        if (s < 0) continue;
        if (hc[i] > song_scores[2 * (size_t)s]) {
            song_scores[2 * (size_t)s] = hc[i];
            song_scores[2 * (size_t)s + 1] = ht[i];
        }
    }
    return *reinterpret_cast<const int *>(hostbuf.data() + 16);
}

int pfann_db_query(pfann_db *h, const float *queries, const int64_t *query_index, int nq, int top_k, int fsm,
                   float alpha, float *best_score, int32_t *best_song, float *best_time, float *song_scores,
                   int64_t n_songs_total) {
    PF_CHECK(h && nq >= 0, PFANN_ERR_ARG, "pfann_db_query: bad argument");
    Db *db = reinterpret_cast<Db *>(h);
    PF_TRY(check_rerank_args(db, top_k, fsm));
    if (nq == 0) return PFANN_OK;
    PF_CHECK(queries && query_index && best_score && best_song && best_time, PFANN_ERR_ARG,
             "pfann_db_query: NULL argument");
    PF_CHECK(!is_device_ptr(query_index), PFANN_ERR_ARG, "pfann_db_query: query_index must be a host array");
    PF_CHECK(!is_device_ptr(best_score) && !is_device_ptr(best_song) && !is_device_ptr(best_time), PFANN_ERR_ARG,
             "pfann_db_query: best_* outputs must be host arrays");
    PF_CHECK(!song_scores || !is_device_ptr(song_scores), PFANN_ERR_ARG,
             "pfann_db_query: song_scores must be a host array");
    PF_CUDA(cudaSetDevice(db->ctx->device));
    Ctx *ctx = db->ctx;
    int64_t total = 0;
    int max_len = 0;
    for (int i = 0; i < nq; i++) {
        const int64_t s = query_index[2 * i], l = query_index[2 * i + 1];
        PF_CHECK(s >= 0 && l >= 0 && l < (1 << 20), PFANN_ERR_ARG, "pfann_db_query: bad query_index row %d", i);
        if (s + l > total) total = s + l;
        if (l > max_len) max_len = (int)l;
    }
    const void *qd, *qid;
    void *bs, *bg, *bt;
    PF_TRY(stage_input(ctx, 0, queries, (size_t)total * db->d * 4, &qd));
    PF_TRY(stage_input(ctx, 1, query_index, (size_t)nq * 16, &qid));
    PF_TRY(db->dist.ensure((size_t)total * top_k * 4));
    PF_TRY(db->labels.ensure((size_t)total * top_k * 8));
    // database.py:172: one search for all rows of all query files
    PF_TRY(db_search_dev(db, (const float *)qd, total, top_k, db->dist.as<float>(), db->labels.as<int64_t>()));
    PF_TRY(stage_output(ctx, 0, best_score, (size_t)nq * 4, &bs));
    PF_TRY(stage_output(ctx, 1, best_song, (size_t)nq * 4, &bg));
    PF_TRY(stage_output(ctx, 2, best_time, (size_t)nq * 4, &bt));
    int capK = 0;
    int *sps = nullptr;
    float *spc = nullptr, *spt = nullptr;
    if (song_scores) {
        int c = 1;
        while (c < max_len * top_k) c <<= 1;
        PF_TRY(db->rr_out.ensure((size_t)nq * c * 12));
        sps = db->rr_out.as<int>();
        spc = reinterpret_cast<float *>(sps + (size_t)nq * c);
        spt = spc + (size_t)nq * c;
    }
    PF_TRY(rerank_dev(db, (const float *)qd, (const int64_t *)qid, nq, max_len, db->labels.as<int64_t>(), top_k, fsm,
                      alpha, (float *)bs, (int *)bg, (float *)bt, sps, spc, spt, &capK));
    PF_TRY(finish_output(ctx, 0, best_score, (size_t)nq * 4));
    PF_TRY(finish_output(ctx, 1, best_song, (size_t)nq * 4));
    PF_TRY(finish_output(ctx, 2, best_time, (size_t)nq * 4));
    if (song_scores) {
        std::vector<unsigned char> hb((size_t)nq * capK * 12);
        PF_CUDA(cudaMemcpyAsync(hb.data(), sps, hb.size(), cudaMemcpyDeviceToHost, ctx->stream));
        PF_CUDA(cudaStreamSynchronize(ctx->stream));
        const int *hs = reinterpret_cast<const int *>(hb.data());
        const float *hc = reinterpret_cast<const float *>(hs + (size_t)nq * capK);
        const float *ht = hc + (size_t)nq * capK;
        memset(song_scores, 0, sizeof(float) * 2 * (size_t)nq * n_songs_total);  // database.py:176
        for (int q = 0; q < nq; q++) {
            float *ss = song_scores + 2 * (size_t)q * n_songs_total;
            for (int i = 0; i < capK; i++) {
                const int s = hs[(size_t)q * capK + i];
                if (s < 0 || s >= n_songs_total) continue;
                if (hc[(size_t)q * capK + i] > ss[2 * (size_t)s]) {
                    ss[2 * (size_t)s] = hc[(size_t)q * capK + i];
                    ss[2 * (size_t)s + 1] = ht[(size_t)q * capK + i];
                }
            }
        }
    }
    apply_zero_floor(nq, best_score, best_song, best_time);
    return PFANN_OK;
}

/* ---- reference-compatible symbols (cpp/seqscore.cpp:27-43) ---- */
long long version(void) { return 20220625002LL; }

int seq_score(void *index, const int64_t *song_pos, int n_songs, const float *query, int query_len,
              const int64_t *labels, int top_k, float *song_scores, int frame_shift_mul, float score_alpha) {
    return pfann_db_seq_score(reinterpret_cast<pfann_db *>(index), song_pos, n_songs, query, query_len, labels,
                              top_k, song_scores, frame_shift_mul, score_alpha);
}

}  // extern "C"
