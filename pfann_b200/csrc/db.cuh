// db.cuh -- stage 3 state shared by knn.cu (search), knn_tc.cu (tensor-core scan) and seqscore.cu (rerank).
#pragma once
#include <vector>

#include "common.cuh"

namespace pfann {

struct Db {
    Ctx *ctx;
    int64_t n;          // rows in this shard
    int d;
    int n_songs;        // songs in this shard
    int64_t id_base;    // global id of local row 0
    int64_t song_base;  // global id of local song 0
    float *emb32 = nullptr;          // [n][d] fp32, exact rescoring + rerank (database.py:155)
    __nv_bfloat16 *emb16 = nullptr;  // [n][d] bf16, tensor-core scan operand
    int64_t *song_pos = nullptr;     // device [n_songs+1]: GLOBAL start row of each local song
    std::vector<int64_t> song_pos_host;
    float max_norm = 0.f;            // max row L2 norm (error bound of the approximate scan)
    // tuning (tests shrink these to force the overflow / backstop paths)
    int cand_cap = 4096;             // candidate slots per query (power of two, <= 4096)
    int sample_rows = 16384;         // rows scanned in the threshold pre-pass (floor)
    float sample_scale = 1.f;        // multiplier of the k * n / 256 target (sharded search: thresholds are max-reduced)
    int use_tc = 1;                  // tensor-core bf16 scan when available, else fp32 CUDA-core scan
    // scratch
    DevBuf qbuf, qnorm, thr, cnt, cand, cand_v, sample, flags, dist, labels, rr_keys, rr_scores, rr_out, lab_stage, rr_xscores, rr_done;
    void *tc_state = nullptr;
    int *ovf_host = nullptr, *ovf_dev = nullptr;  // deferred overflow count of the sharded search (pinned, mapped)
};

// exact canonical inner product shared with the oracle (oracle/pfann_oracle.c dot_fma_seq):
// one fused multiply-add per element, k = 0..d-1 in order.
#ifdef __CUDACC__
__device__ __forceinline__ float dot_fma_seq(const float *__restrict__ a, const float *__restrict__ b, int d) {
    float acc = 0.f;
    for (int k = 0; k < d; k++) acc = __fmaf_rn(a[k], b[k], acc);
    return acc;
}
// Same arithmetic (one fmaf per element, k ascending -> bit-identical result), but the row is fetched one
// 128-byte line ahead of the dependent FMA chain so a thread keeps two lines in flight.  `row` is global and
// 16-byte aligned, `q` is in shared memory, d % 32 == 0.
__device__ __forceinline__ float dot_fma_seq_lines(const float *__restrict__ row, const float *__restrict__ q, int d) {
    const float4 *r4 = reinterpret_cast<const float4 *>(row);
    float4 cur[8], nxt[8];
#pragma unroll
    for (int i = 0; i < 8; i++) cur[i] = __ldg(r4 + i);
    float acc = 0.f;
    for (int l = 0; l < d; l += 32) {
        if (l + 32 < d) {
#pragma unroll
            for (int i = 0; i < 8; i++) nxt[i] = __ldg(r4 + (l >> 2) + 8 + i);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            acc = __fmaf_rn(cur[i].x, q[l + 4 * i], acc);
            acc = __fmaf_rn(cur[i].y, q[l + 4 * i + 1], acc);
            acc = __fmaf_rn(cur[i].z, q[l + 4 * i + 2], acc);
            acc = __fmaf_rn(cur[i].w, q[l + 4 * i + 3], acc);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) cur[i] = nxt[i];
    }
    return acc;
}
// order-preserving float -> uint32 (larger float -> larger key)
__device__ __forceinline__ uint32_t flipf(float f) {
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float unflipf(uint32_t u) {
    return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
// In-place bitonic sort of n (power of two) keys by one CTA; descending if DESC.
template <typename K, bool DESC>
__device__ void bitonic_sort(K *keys, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));  // index with bit `stride` cleared
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);  // ascending block?
                const K a = keys[lo], b = keys[hi];
                const bool swap = DESC ? (up ? a < b : a > b) : (up ? a > b : a < b);
                if (swap) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
        }
    }
    __syncthreads();
}
#endif

// knn.cu: device-pointer search (q, dist, labels all on the device, stream-ordered except for the
// overflow check which synchronises once per query group)
int db_search_dev(Db *db, const float *q, int64_t Q, int k, float *dist, int64_t *labels);
int db_search_thresholds_dev(Db *db, const float *q, int64_t Q, int k, float *thr, uint32_t *topk_out = nullptr);
int db_search_filtered_dev(Db *db, const float *q, int64_t Q, int k, float *thr, unsigned long long *keys, bool defer);
// knn_tc.cu
int knn_tc_prepare(Db *db);
int64_t knn_tc_sample_slots(Db *db, int64_t r0, int64_t r1, int ngroups);
void knn_tc_release(Db *db);
// approximate scan of rows [r0, r1) against Qtot queries in groups of `group` <= 256 (one database pass per group,
// all groups in ONE launch); mode 0: per-thread maxima of the scores to sample[q][slot] (ld = sample_ld);
// mode 1: push row ids with score >= thr[q] into cand/cnt
int knn_tc_scan(Db *db, const float *q, int Qtot, int group, int64_t r0, int64_t r1, int mode, float *sample,
                int64_t sample_ld, const float *thr, int *cnt, uint32_t *cand, uint32_t *cand_v, int cap);

}  // namespace pfann
