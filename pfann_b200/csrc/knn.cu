// knn.cu -- stage 3a: brute-force inner-product top-k over the HBM-resident fingerprint database.
//
// Replaces faiss IndexFlatIP.search as called from database.py:121,172.  Structure (per group of queries):
//   1. threshold pre-pass: scan a sample of rows, take the k-th best sample score per query.  That is a
//      LOWER bound of the global k-th best score, so (minus the scan's error bound) it is a safe filter;
//   2. one streaming scan of the whole shard (tcgen05 bf16 GEMM in knn_tc.cu, or the fp32 CUDA-core kernel
//      below): rows whose approximate score passes the filter are appended to a per-query candidate list;
//   3. select: every candidate is re-scored EXACTLY in fp32 with the canonical k-sequential fused
//      multiply-add (bit-identical to the oracle), the list is bitonic-sorted by (score desc, id asc) and the
//      first k are emitted.  Overflowing lists tighten the threshold from their own exact scores and rescan;
//      a brute-force exact kernel is the backstop for degenerate inputs (all-equal rows etc.).
// The result therefore does not depend on the precision of the scan: labels and distances are those of an
// exact fp32 search.
#include <float.h>
#include <math.h>
#include <string.h>

#include "db.cuh"
#include "pfann_b200.h"

using namespace pfann;

namespace {

constexpr int QG = 32;        // queries per CUDA-core scan pass
constexpr int SCAN_ROWS = 64; // rows per smem tile
constexpr int SCAN_THREADS = 256;

// ---- fp32 CUDA-core scan ---------------------------------------------------------------------------
// grid.x = row tiles (grid-stride), 256 threads; thread -> (row r = tid % 64, query octet qs = tid / 64)
__global__ void __launch_bounds__(SCAN_THREADS) knn_scan_fp32_kernel(
    const float *__restrict__ db, int64_t r0, int64_t r1, int d, const float *__restrict__ q, int Qg, int mode,
    float *sample, int64_t sample_ld, const float *__restrict__ thr, int *cnt, uint32_t *cand, uint32_t *cand_v,
    int cap) {
    extern __shared__ __align__(16) float sm[];
    const int ldr = d + 4;                 // padded row stride (16-byte aligned, conflict-free for float4)
    float *qs_ = sm;                       // [QG][d]
    float *rows = sm + QG * d;             // [SCAN_ROWS][ldr]
    const int tid = threadIdx.x;
    for (int i = tid; i < QG * d; i += SCAN_THREADS) qs_[i] = (i < Qg * d) ? q[i] : 0.f;
    const int r = tid & (SCAN_ROWS - 1), qo = (tid >> 6) * 8;
    const int64_t ntiles = (r1 - r0 + SCAN_ROWS - 1) / SCAN_ROWS;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = r0 + tile * SCAN_ROWS;
        __syncthreads();
        // coalesced tile load: SCAN_ROWS x d floats
        const int nvec = SCAN_ROWS * d / 4;
        for (int i = tid; i < nvec; i += SCAN_THREADS) {
            const int rr = (i * 4) / d, cc = (i * 4) - rr * d;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (base + rr < r1) v = __ldg(reinterpret_cast<const float4 *>(db + (base + rr) * d + cc));
            *reinterpret_cast<float4 *>(rows + rr * ldr + cc) = v;
        }
        __syncthreads();
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = 0.f;
        const float *xr = rows + r * ldr;
        for (int k = 0; k < d; k += 4) {
            const float4 x = *reinterpret_cast<const float4 *>(xr + k);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 w = *reinterpret_cast<const float4 *>(qs_ + (qo + j) * d + k);
                acc[j] = fmaf(x.x, w.x, acc[j]);
                acc[j] = fmaf(x.y, w.y, acc[j]);
                acc[j] = fmaf(x.z, w.z, acc[j]);
                acc[j] = fmaf(x.w, w.w, acc[j]);
            }
        }
        const int64_t row = base + r;
        if (row < r1) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int qi = qo + j;
                if (qi >= Qg) break;
                if (mode == 0) {
                    sample[qi * sample_ld + (row - r0)] = acc[j];
                } else if (acc[j] >= thr[qi]) {
                    const int pos = atomicAdd(cnt + qi, 1);
                    if (pos < cap) {
                        cand[(int64_t)qi * cap + pos] = (uint32_t)row;
                        cand_v[(int64_t)qi * cap + pos] = __float_as_uint(fmaxf(acc[j] - thr[qi], 0.f));
                    }
                }
            }
        }
    }
}

// ---- per-query threshold from the sample scores ---------------------------------------------------------
// Stage 1: grid (Qg, nchunks): every CTA sorts one chunk (<= 8192 scores) of a query's sample in shared memory
// and keeps its k best keys.  Stage 2: one CTA per query merges the nchunks*k survivors; the k-th best of the
// whole sample is then  thr[q] = kth - 2*eps*|q|*max_norm  (-inf when the sample has fewer than k rows).
__global__ void __launch_bounds__(256) knn_kth_chunk_kernel(const float *sample, int64_t sample_ld, int S, int chunk,
                                                            int cpad, int k, uint32_t *part) {
    extern __shared__ uint32_t keys[];  // [cpad] (full sort) or [256] (thread-max path)
    const int qi = blockIdx.x, c = blockIdx.y, nchunks = gridDim.y;
    const int lo = c * chunk;
    const int n = (S - lo) < chunk ? (S - lo) : chunk;
    uint32_t *out = part + ((int64_t)qi * nchunks + c) * k;
    if (k <= 128) {
        // Cheap and still safe: every thread keeps the maximum of its strided slice; the k-th largest of those 256
        // maxima is an actual sample score that is <= the chunk's true k-th largest (each maximum is a distinct
        // element), i.e. still a valid lower bound -- and nearly tight, since the top k rarely share a slice.
        uint32_t m = 0u;
        for (int i = threadIdx.x; i < n; i += 256) {
            const uint32_t key = flipf(sample[qi * sample_ld + lo + i]);
            m = key > m ? key : m;
        }
        keys[threadIdx.x] = m;
        bitonic_sort<uint32_t, true>(keys, 256);
        for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = keys[i];
        return;
    }
    for (int i = threadIdx.x; i < cpad; i += blockDim.x) keys[i] = i < n ? flipf(sample[qi * sample_ld + lo + i]) : 0u;
    bitonic_sort<uint32_t, true>(keys, cpad);
    for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = i < n ? keys[i] : 0u;
}

__global__ void __launch_bounds__(256) knn_kth_merge_kernel(const uint32_t *part, int nkeys, int P, int S,
                                                            const float *q, int d, int k, float eps_rel,
                                                            float max_norm, float *thr, float *qnorm, uint32_t *topk_out) {
    extern __shared__ uint32_t keys[];  // [P]
    __shared__ float red[8];
    const int qi = blockIdx.x;
    float ss = 0.f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) ss = fmaf(q[qi * d + i], q[qi * d + i], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = i < nkeys ? part[(int64_t)qi * nkeys + i] : 0u;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < 8; i++) tot += red[i];
    const float qn = sqrtf(tot);
    bitonic_sort<uint32_t, true>(keys, P);
    if (threadIdx.x == 0) {
        qnorm[qi] = qn;
        thr[qi] = (S >= k) ? unflipf(keys[k - 1]) - 2.f * eps_rel * qn * max_norm : -INFINITY;
    }
    // sharded search: the k best sampled scan scores themselves (sortable keys, 0 = none), so that the ranks can take
    // the k-th best of the UNION of their samples instead of the maximum of their k-th bests
    if (topk_out != nullptr)
        for (int i = threadIdx.x; i < k; i += blockDim.x) topk_out[(int64_t)qi * k + i] = (i < nkeys && i < S) ? keys[i] : 0u;
}

// thr[q] = k-th best of the gathered per-shard sample top-k lists [world][Q][k] minus the scan's error bound
__global__ void __launch_bounds__(256) thr_union_kernel(const uint32_t *gathered, int world, int64_t Q, int P, const float *q,
                                                        int d, int k, float eps_rel, float max_norm, float *thr) {
    extern __shared__ uint32_t keys[];  // [P]
    __shared__ float red[8];
    const int64_t qi = blockIdx.x;
    float ss = 0.f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) ss = fmaf(q[qi * d + i], q[qi * d + i], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int w = i / k, j = i - w * k;
        keys[i] = w < world ? gathered[((int64_t)w * Q + qi) * k + j] : 0u;
    }
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < 8; i++) tot += red[i];
    bitonic_sort<uint32_t, true>(keys, P);
    if (threadIdx.x == 0) thr[qi] = keys[k - 1] != 0u ? unflipf(keys[k - 1]) - 2.f * eps_rel * sqrtf(tot) * max_norm : -INFINITY;
}

// ---- select: rank by scan score, exact rescoring of the few that can matter, sort, emit ------------------------
// The scan leaves ~1-2 k survivors per query, each with v = (scan score - threshold).  |scan - exact| <= delta =
// eps_rel |q| max|x|, so with a_k the k-th best scan score every true top-k row has a scan score >= a_k - 2 delta:
// only that prefix of the ranking (a few dozen rows) is re-scored exactly in fp32 (gathering 512-byte rows of the
// fp32 database is what used to dominate this kernel), sorted by (score desc, id asc) and emitted.
// flags[0] += 1 for every query whose candidate list overflowed (its threshold is tightened in place).
__global__ void __launch_bounds__(256, 3) knn_select_kernel(const float *__restrict__ db, int d, int64_t id_base,
                                                         const float *__restrict__ q, const int *cnt,
                                                         const uint32_t *cand, const uint32_t *cand_v, int cap, int k,
                                                         float eps_rel, float max_norm, const float *qnorm, float *thr,
                                                         float *dist, int64_t *labels, int *flags) {
    extern __shared__ __align__(16) unsigned char smraw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smraw);  // [cap] survivors of the ranking
    float *qv = reinterpret_cast<float *>(keys + cap);                         // [d]
    __shared__ int hist[8][256];   // per-warp digit histograms of the radix select
    __shared__ uint32_t sel_prefix;
    __shared__ int sel_k, m_s;
    const int qi = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
    const int total = cnt[qi];
    const int n = total < cap ? total : cap;
    for (int i = tid; i < d; i += blockDim.x) qv[i] = q[(int64_t)qi * d + i];
    // this thread's share of the scan scores as order-preserving integer keys (cap <= 4096 -> at most 16)
    uint32_t kv[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int i = tid + 256 * j;
        kv[j] = i < n ? flipf(__uint_as_float(cand_v[(int64_t)qi * cap + i])) : 0u;
    }
    // (a lower bound of the) k-th largest key by a most-significant-digit radix select, 8 bits per round: no sort of the
    // ~2 k candidates
    uint32_t prefix = 0u, mask = 0u;
    if (tid == 0) {
        sel_k = k;
        m_s = 0;
    }
    const bool have_k = n >= k;
    // Two rounds (sign, exponent, 7 mantissa bits) are enough: the prefix with its low 16 bits cleared is a LOWER bound
    // of the k-th best key within 2^-7 of its value, far inside the 2 delta margin below -- a few more candidates
    // are re-scored exactly, the result is unchanged (half the rounds, and the rounds are fixed cost per query).
    for (int shift = 24; shift >= 16 && have_k; shift -= 8) {
        for (int i = tid; i < 8 * 256; i += 256) (&hist[0][0])[i] = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (tid + 256 * j < n && (kv[j] & mask) == prefix) atomicAdd(&hist[warp][(kv[j] >> shift) & 255u], 1);
        __syncthreads();
        if (warp == 0) {
            // lane l owns digits [8 l, 8 l + 8); walk from the largest digit down to the one holding the k-th key
            const int lane = tid;
            int c8[8], mine = 0;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                int c = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) c += hist[w][8 * lane + e];
                c8[e] = c;
                mine += c;
            }
            int suf = mine;  // inclusive suffix sum over lanes: keys in this lane's digits and all larger ones
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_down_sync(0xffffffffu, suf, o);
                if (lane + o < 32) suf += t;
            }
            const int above = suf - mine;
            const int kk = sel_k;
            __syncwarp();  // every lane has read sel_k before the owning lane replaces it
            if (above < kk && kk <= above + mine) {  // exactly one lane
                int acc = above;
                int digit = 0, newk = kk;
#pragma unroll
                for (int e = 7; e >= 0; e--) {
                    if (acc < kk && kk <= acc + c8[e]) {
                        digit = 8 * lane + e;
                        newk = kk - acc;
                    }
                    acc += c8[e];
                }
                sel_prefix = prefix | ((uint32_t)digit << shift);
                sel_k = newk;
            }
        }
        __syncthreads();
        prefix = sel_prefix;
        mask |= 255u << shift;
    }
    __syncthreads();
    const float delta2 = 2.f * eps_rel * qnorm[qi] * max_norm;
    const float vk = have_k ? unflipf(prefix) : -INFINITY;  // k-th best (scan score - threshold)
    if (total > cap) {
        // overflow: the stored subset still bounds the k-th best scan score from below
        if (tid == 0) {
            atomicAdd(flags, 1);
            if (have_k) thr[qi] = fmaxf(thr[qi], thr[qi] + vk - delta2);
        }
        return;
    }
    // every candidate that can still be a true top-k row: scan score >= k-th best - 2 delta
    const uint32_t kmin = have_k ? flipf(vk - delta2) : 0u;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int i = tid + 256 * j;
        if (i < n && kv[j] >= kmin) {
            const int pos = atomicAdd(&m_s, 1);
            keys[pos] = (unsigned long long)cand[(int64_t)qi * cap + i];
        }
    }
    __syncthreads();
    const int m = m_s;
    int P2 = 1;
    while (P2 < m) P2 <<= 1;
    for (int i = tid; i < P2; i += blockDim.x) {
        unsigned long long key = 0ull;
        if (i < m) {
            const uint32_t id = (uint32_t)keys[i];
            const float s = (d & 31) == 0 ? dot_fma_seq_lines(db + (int64_t)id * d, qv, d) : dot_fma_seq(db + (int64_t)id * d, qv, d);
            key = ((unsigned long long)flipf(s) << 32) | (unsigned long long)(0xFFFFFFFFu - id);
        }
        keys[i] = key;  // slot i is read and written by this thread only
    }
    bitonic_sort<unsigned long long, true>(keys, P2);
    for (int j = tid; j < k; j += blockDim.x) {
        if (j < m) {
            dist[(int64_t)qi * k + j] = unflipf((uint32_t)(keys[j] >> 32));
            labels[(int64_t)qi * k + j] = id_base + (int64_t)(0xFFFFFFFFu - (uint32_t)(keys[j] & 0xFFFFFFFFull));
        } else {
            dist[(int64_t)qi * k + j] = -FLT_MAX;
            labels[(int64_t)qi * k + j] = -1;
        }
    }
}

// ---- backstop: exact brute force, one CTA per query (degenerate inputs only) ------------------------------
__global__ void __launch_bounds__(256) knn_exact_kernel(const float *__restrict__ db, int64_t n, int d,
                                                        int64_t id_base, const float *__restrict__ q, int k,
                                                        float *dist, int64_t *labels) {
    extern __shared__ __align__(16) unsigned char smraw[];
    unsigned long long *best = reinterpret_cast<unsigned long long *>(smraw);  // [k] sorted descending
    unsigned long long *batch = best + k;                                      // [256]
    float *qv = reinterpret_cast<float *>(batch + 256);                        // [d]
    __shared__ int nbest;
    const int qi = blockIdx.x;
    for (int i = threadIdx.x; i < d; i += blockDim.x) qv[i] = q[(int64_t)qi * d + i];
    if (threadIdx.x == 0) nbest = 0;
    __syncthreads();
    for (int64_t r0 = 0; r0 < n; r0 += 256) {
        const int64_t r = r0 + threadIdx.x;
        unsigned long long key = 0ull;
        if (r < n) {
            const float s = dot_fma_seq(db + r * d, qv, d);
            key = ((unsigned long long)flipf(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)r);
        }
        batch[threadIdx.x] = key;
        __syncthreads();
        if (threadIdx.x == 0) {
            int nb = nbest;
            for (int i = 0; i < 256; i++) {
                const unsigned long long kk = batch[i];
                if (kk == 0ull) continue;
                if (nb == k && kk <= best[k - 1]) continue;
                int p = nb < k ? nb : k - 1;
                while (p > 0 && best[p - 1] < kk) {
                    best[p] = best[p - 1];
                    p--;
                }
                best[p] = kk;
                if (nb < k) nb++;
            }
            nbest = nb;
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        if (j < nbest) {
            dist[(int64_t)qi * k + j] = unflipf((uint32_t)(best[j] >> 32));
            labels[(int64_t)qi * k + j] = id_base + (int64_t)(0xFFFFFFFFu - (uint32_t)(best[j] & 0xFFFFFFFFull));
        } else {
            dist[(int64_t)qi * k + j] = -FLT_MAX;
            labels[(int64_t)qi * k + j] = -1;
        }
    }
}

// ---- helpers ---------------------------------------------------------------------------------------------
__global__ void row_norm_max_kernel(const float *db, int64_t n, int d, unsigned int *out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f;
    if (r < n)
        for (int k = 0; k < d; k++) s = fmaf(db[r * d + k], db[r * d + k], s);
    s = sqrtf(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fmaxf(s, __shfl_xor_sync(0xffffffffu, s, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(s));  // non-negative floats order like uints
}

__global__ void to_bf16_kernel(const float *in, __nv_bfloat16 *out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}

__global__ void fill_empty_kernel(float *dist, int64_t *labels, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        dist[i] = -FLT_MAX;
        labels[i] = -1;
    }
}

// merge G gathered top-k lists: in [G][Q][k] -> out [Q][k]; (score desc, id asc); -1 labels sort last
__global__ void topk_merge_kernel(const float *dist_g, const int64_t *labels_g, int G, int64_t Q, int k, int P,
                                  float *dist, int64_t *labels) {
    extern __shared__ __align__(16) unsigned char smraw[];
    // 96-bit ordering (score, -id) does not fit a 64-bit key with int64 ids: sort indices by comparing pairs
    float *s = reinterpret_cast<float *>(smraw);            // [P]
    int64_t *id = reinterpret_cast<int64_t *>(s + P + (P & 1));  // [P]
    const int64_t qi = blockIdx.x;
    const int n = G * k;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        if (i < n) {
            const int g = i / k, j = i - g * k;
            s[i] = dist_g[((int64_t)g * Q + qi) * k + j];
            id[i] = labels_g[((int64_t)g * Q + qi) * k + j];
        } else {
            s[i] = -FLT_MAX;
            id[i] = -1;
        }
    }
    // bitonic sort, descending by (valid, score, -id)
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const float sa = s[lo], sb = s[hi];
                const int64_t ia = id[lo], ib = id[hi];
                // a_before_b: a should precede b in descending order
                const bool a_valid = ia >= 0, b_valid = ib >= 0;
                bool a_before_b;
                if (a_valid != b_valid) a_before_b = a_valid;
                else if (sa != sb) a_before_b = sa > sb;
                else a_before_b = ia <= ib;
                const bool swap = up ? !a_before_b : (a_before_b && !(sa == sb && ia == ib));
                if (swap) {
                    s[lo] = sb; s[hi] = sa;
                    id[lo] = ib; id[hi] = ia;
                }
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        dist[qi * k + j] = s[j];
        labels[qi * k + j] = id[j];
    }
}

// one launch for up to SG database passes (tensor-core scan) or one launch per QG-query slice (fp32 CUDA-core scan)
int scan(Db *db, bool tc, const float *q, int Qs, int group, int64_t r0, int64_t r1, int mode, float *sample,
         int64_t sample_ld, const float *thr, int *cnt, uint32_t *cand, uint32_t *cand_v, int cap) {
    if (tc) return knn_tc_scan(db, q, Qs, group, r0, r1, mode, sample, sample_ld, thr, cnt, cand, cand_v, cap);
    const int64_t ntiles = (r1 - r0 + SCAN_ROWS - 1) / SCAN_ROWS;
    int64_t grid = (int64_t)db->ctx->sm_count * 4;
    if (grid > ntiles) grid = ntiles;
    const size_t smem = (size_t)(QG * db->d + SCAN_ROWS * (db->d + 4)) * 4;
    for (int q0 = 0; q0 < Qs; q0 += QG) {
        const int qn = (Qs - q0) < QG ? (Qs - q0) : QG;
        ProfScope ps(db->ctx, K_KNN_SCAN, mode == 0 ? 35 : 36);
        knn_scan_fp32_kernel<<<(unsigned)grid, SCAN_THREADS, smem, db->ctx->stream>>>(
            db->emb32, r0, r1, db->d, q + (int64_t)q0 * db->d, qn, mode, sample ? sample + q0 * sample_ld : nullptr,
            sample_ld, thr ? thr + q0 : nullptr, cnt ? cnt + q0 : nullptr, cand ? cand + (int64_t)q0 * cap : nullptr,
            cand_v ? cand_v + (int64_t)q0 * cap : nullptr, cap);
        db->ctx->launches++;
    }
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

// deferred overflow check: add the per-super-group overflow flags of one call to a host-visible counter (launches of
// one stream are serialised: a plain read-modify-write by one thread is enough)
__global__ void accumulate_flags_kernel(const int *flags, int n, volatile int *total) {
    int s = 0;
    for (int i = 0; i < n; i++) s += flags[i];
    if (s) *total = *total + s;
}

__global__ void fill_f32_kernel(float *dst, float v, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

// (dist, labels) -> one sortable 64-bit key per entry: flipped fp32 score in the high word, 0xFFFFFFFF - global id in
// the low word (descending key = score descending, id ascending); empty slots (-1) become 0.
__global__ void pack_keys_kernel(const float *dist, const int64_t *labels, int64_t n, unsigned long long *keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t id = labels[i];
    keys[i] = id < 0 ? 0ull
                     : (((unsigned long long)flipf(dist[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)id));
}

// merge G gathered key lists [G][Q][k] -> global top-k per query: dist/labels [Q][k] (-FLT_MAX / -1 padding)
__global__ void merge_keys_kernel(const unsigned long long *keys_g, int G, int64_t Q, int k, int P, float *dist,
                                  int64_t *labels) {
    extern __shared__ __align__(16) unsigned char smraw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smraw);
    const int64_t qi = blockIdx.x;
    const int n = G * k;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int g = i / k, j = i - g * k;
        keys[i] = i < n ? keys_g[((int64_t)g * Q + qi) * k + j] : 0ull;
    }
    bitonic_sort<unsigned long long, true>(keys, P);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const unsigned long long key = keys[j];
        if (key != 0ull) {
            if (dist) dist[qi * k + j] = unflipf((uint32_t)(key >> 32));
            labels[qi * k + j] = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
        } else {
            if (dist) dist[qi * k + j] = -FLT_MAX;
            labels[qi * k + j] = -1;
        }
    }
}

// per-shard winners [G][nq] x (score, song, time) -> global winner per query file: score desc, then lower song id
// (cpp/seqscore.cpp:121), then the reference's zero floor (database.py:176,190).  out: [nq] x (score, song bits, time, 0)
__global__ void combine_best_kernel(const float4 *packed_g, int G, int nq, float4 *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    float bs = 0.f, bt = 0.f;
    int bg = -1;
    for (int g = 0; g < G; g++) {
        const float4 v = packed_g[(size_t)g * nq + i];
        const int song = __float_as_int(v.y);
        if (song < 0) continue;
        if (bg < 0 || v.x > bs || (v.x == bs && song < bg)) {
            bs = v.x; bg = song; bt = v.z;
        }
    }
    if (!(bg >= 0 && bs > 0.f)) bs = 0.f, bt = 0.f;   // nothing above the zero-initialised table
    out[i] = make_float4(bs, __int_as_float(bg), bt, 0.f);
}

struct SearchPlan {
    bool tc;
    float eps_rel;
    int cap, group, sgroup, S, chunk;
    int64_t ngroups;
};

int plan_search(Db *db, int64_t Q, int k, SearchPlan *p) {
    PF_CHECK(k > 0 && k <= db->cand_cap && k <= 2048, PFANN_ERR_UNSUPPORTED, "database search: k=%d too large (max %d)", k,
             db->cand_cap < 2048 ? db->cand_cap : 2048);
    const int d = db->d;
    p->tc = db->use_tc && db->tc_state != nullptr;
    // |scan score - exact score| <= eps_rel * |q| * |x|: bf16 operand rounding 2^-8 (+ fp32 accumulation order)
    p->eps_rel = (p->tc ? 4.0e-3f : 0.f) + fmaxf(1.0e-5f, 1.2e-7f * (float)d);
    p->cap = db->cand_cap;
    p->group = p->tc ? (d <= 128 ? 256 : 128) : QG * 4;  // queries per database pass
    // sample size: aim at ~256 rows above the threshold (k * n / S ~ 256), within [sample_rows, 256 Ki].  Every
    // survivor costs ~100 instructions on the filter's hit path (measured: 360 k survivors per 256-query pass doubled
    // the scan time of a 1.25 M-row shard), a sampled row costs one more tile of the cheap pre-pass.
    // Measured on 10 M rows, 190 k queries: 262 k sampled rows -> pre-pass 26 ms + filtered scan 470 ms; 524 k -> 50 + 418;
    // 781 k -> 73 + 393 (the pre-pass costs twice the filtered scan per row: it keeps maxima instead of testing signs).
    int64_t cap_rows = 524288;
    if (const char *e = getenv("PFANN_B200_SAMPLE_ROWS_MAX")) cap_rows = atoll(e);   // experiment knob
    int64_t want = (int64_t)((double)k * db->n / 256);
    if (want > cap_rows) want = cap_rows;
    want = (int64_t)((double)want * db->sample_scale);   // sharded search: the threshold comes from all shards' samples
    if (want < db->sample_rows) want = db->sample_rows;
    p->chunk = 8192;  // one kth-select CTA sorts this many sample scores in shared memory
    int64_t max_chunks = 8192 / k;  // stage 2 sorts nchunks * k keys in shared memory
    if (want > max_chunks * p->chunk) want = max_chunks * p->chunk;
    p->S = (int)(db->n < want ? db->n : want);
    // A "super-group" = up to SG database passes (SG * group queries) that share ONE launch of every kernel (scans
    // included: the CTAs are split between the passes): on a shard of a million rows a pass is ~75 us and launch gaps,
    // prologues and tails would otherwise cost as much again.
    const int SG = 4;
    p->sgroup = SG * p->group;
    p->ngroups = (Q + p->sgroup - 1) / p->sgroup;
    PF_TRY(db->thr.ensure(sizeof(float) * (size_t)p->sgroup));
    PF_TRY(db->qnorm.ensure(sizeof(float) * (size_t)(Q > p->sgroup ? Q : p->sgroup)));
    PF_TRY(db->cnt.ensure(sizeof(int) * p->sgroup));
    PF_TRY(db->cand.ensure(sizeof(uint32_t) * (size_t)p->sgroup * p->cap));
    PF_TRY(db->cand_v.ensure(sizeof(uint32_t) * (size_t)p->sgroup * p->cap));
    PF_TRY(db->flags.ensure(sizeof(int) * (size_t)(p->ngroups + 4)));
    const size_t sel_smem = (size_t)p->cap * 8 + (size_t)d * 4;
    PF_CUDA(cudaFuncSetAttribute(knn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
    PF_CUDA(cudaFuncSetAttribute(knn_scan_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((QG * d + SCAN_ROWS * (d + 4)) * 4)));
    return PFANN_OK;
}

// 1. threshold pre-pass on the first S rows for Qs <= sgroup queries: thr[q] = a lower bound of the k-th best scan
// score of this shard (minus the scan's error bound), qnorm[q] = |q|
int prepass(Db *db, const SearchPlan &p, const float *qs, int Qs, int k, float *thr, float *qnorm, uint32_t *topk_out = nullptr) {
    cudaStream_t st = db->ctx->stream;
    const int d = db->d;
    const int ng = (Qs + p.group - 1) / p.group;
    // the tensor-core pre-pass leaves per-thread maxima (one per CTA and accumulator row) instead of every score
    const int S_keys = p.tc ? (int)knn_tc_sample_slots(db, 0, p.S, ng) : p.S;
    const int nchunks = (S_keys + p.chunk - 1) / p.chunk;
    int cpad = 1;
    while (cpad < (S_keys < p.chunk ? S_keys : p.chunk)) cpad <<= 1;
    int P2 = 1;
    while (P2 < nchunks * k) P2 <<= 1;
    PF_TRY(db->sample.ensure(sizeof(float) * (size_t)Qs * S_keys));
    PF_TRY(db->rr_keys.ensure(sizeof(uint32_t) * (size_t)Qs * nchunks * k));
    PF_TRY(scan(db, p.tc, qs, Qs, p.group, 0, p.S, 0, db->sample.as<float>(), S_keys, nullptr, nullptr, nullptr, nullptr, 0));
    ProfScope ps(db->ctx, K_KNN_SELECT, 37);
    knn_kth_chunk_kernel<<<dim3(Qs, nchunks), 256, (size_t)(cpad > 256 ? cpad : 256) * 4, st>>>(
        db->sample.as<float>(), S_keys, S_keys, p.chunk, cpad, k, db->rr_keys.as<uint32_t>());
    knn_kth_merge_kernel<<<Qs, 256, (size_t)P2 * 4, st>>>(db->rr_keys.as<uint32_t>(), nchunks * k, P2, S_keys, qs, d, k,
                                                         p.eps_rel, db->max_norm, thr, qnorm, topk_out);
    db->ctx->launches += 2;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

// 2. filtered scan of the whole shard, 3. ranking + exact rescoring; *flag += 1 per overflowed query
int filtered(Db *db, const SearchPlan &p, const float *qs, int Qs, int k, float *thr, const float *qnorm, float *dist,
             int64_t *labels, int *flag) {
    cudaStream_t st = db->ctx->stream;
    int *cnt = db->cnt.as<int>();
    uint32_t *cand = db->cand.as<uint32_t>(), *cand_v = db->cand_v.as<uint32_t>();
    PF_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * Qs, st));
    PF_TRY(scan(db, p.tc, qs, Qs, p.group, 0, db->n, 1, nullptr, 0, thr, cnt, cand, cand_v, p.cap));
    ProfScope ps(db->ctx, K_KNN_SELECT, 38);
    const size_t sel_smem = (size_t)p.cap * 8 + (size_t)db->d * 4;
    knn_select_kernel<<<Qs, 256, sel_smem, st>>>(db->emb32, db->d, db->id_base, qs, cnt, cand, cand_v, p.cap, k, p.eps_rel,
                                                 db->max_norm, qnorm, thr, dist, labels, flag);
    db->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

// Every super-group has run once without any host synchronisation; the (rare) ones with an overflowed candidate
// list are found with ONE read-back of the flags and redone with tightened thresholds.  `thr_all` (optional, [Q])
// holds externally supplied thresholds (sharded search); otherwise the shard's own pre-pass is repeated.
int redo_overflowed(Db *db, const SearchPlan &p, const float *q, int64_t Q, int k, float *thr_all, float *dist,
                    int64_t *labels) {
    cudaStream_t st = db->ctx->stream;
    const int d = db->d;
    int *flags = db->flags.as<int>();
    std::vector<int> hflags((size_t)p.ngroups);
    PF_CUDA(cudaMemcpyAsync(hflags.data(), flags, sizeof(int) * (size_t)p.ngroups, cudaMemcpyDeviceToHost, st));
    PF_CUDA(cudaStreamSynchronize(st));
    for (int64_t g = 0; g < p.ngroups; g++) {
        if (hflags[g] == 0) continue;
        const int64_t q0 = g * p.sgroup;
        const int Qs = (int)((Q - q0) < p.sgroup ? (Q - q0) : p.sgroup);
        const float *qs = q + q0 * d;
        float *thr = thr_all ? thr_all + q0 : db->thr.as<float>();
        float *qn = db->qnorm.as<float>() + (thr_all ? q0 : 0);
        if (!thr_all) PF_TRY(prepass(db, p, qs, Qs, k, thr, qn));
        bool done = false;
        for (int iter = 0; iter < 5 && !done; iter++) {  // without external thresholds iteration 0 repeats the overflow
            PF_CUDA(cudaMemsetAsync(flags + g, 0, sizeof(int), st));
            PF_TRY(filtered(db, p, qs, Qs, k, thr, qn, dist + q0 * k, labels + q0 * k, flags + g));
            int overflow = 0;
            PF_CUDA(cudaMemcpyAsync(&overflow, flags + g, sizeof(int), cudaMemcpyDeviceToHost, st));
            PF_CUDA(cudaStreamSynchronize(st));
            done = (overflow == 0);
        }
        if (!done) {
            // degenerate score distribution (e.g. massive exact ties): exact brute force for this super-group
            const size_t smem = (size_t)(k + 256) * 8 + (size_t)d * 4;
            PF_CUDA(cudaFuncSetAttribute(knn_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            knn_exact_kernel<<<Qs, 256, smem, st>>>(db->emb32, db->n, d, db->id_base, qs, k, dist + q0 * k,
                                                    labels + q0 * k);
            db->ctx->launches++;
            PF_CUDA(cudaGetLastError());
        }
    }
    return PFANN_OK;
}

}  // namespace

namespace pfann {

int db_search_dev(Db *db, const float *q, int64_t Q, int k, float *dist, int64_t *labels) {
    cudaStream_t st = db->ctx->stream;
    if (Q == 0) return PFANN_OK;
    SearchPlan p;
    PF_TRY(plan_search(db, Q, k, &p));   // every caller (pfann_db_search / _query / _rerank) ends up here: k is checked
    if (db->n == 0) {
        fill_empty_kernel<<<cdiv(Q * k, 256), 256, 0, st>>>(dist, labels, Q * k);
        db->ctx->launches++;
        PF_CUDA(cudaGetLastError());
        return PFANN_OK;
    }
    int *flags = db->flags.as<int>();
    PF_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)(p.ngroups + 4), st));
    for (int64_t g = 0; g < p.ngroups; g++) {
        const int64_t q0 = g * p.sgroup;
        const int Qs = (int)((Q - q0) < p.sgroup ? (Q - q0) : p.sgroup);
        PF_TRY(prepass(db, p, q + q0 * db->d, Qs, k, db->thr.as<float>(), db->qnorm.as<float>()));
        PF_TRY(filtered(db, p, q + q0 * db->d, Qs, k, db->thr.as<float>(), db->qnorm.as<float>(), dist + q0 * k,
                        labels + q0 * k, flags + g));
    }
    return redo_overflowed(db, p, q, Q, k, nullptr, dist, labels);
}

// Sharded search, phase 1 (device pointers): per-query thresholds of THIS shard for all Q queries.  A caller that
// searches several shards takes the element-wise maximum over the shards (each is a lower bound of the GLOBAL k-th
// best score, provided every shard uses the same error bound: pfann_db_set_max_norm) and hands it to phase 2.
int db_search_thresholds_dev(Db *db, const float *q, int64_t Q, int k, float *thr, uint32_t *topk_out) {
    if (Q == 0) return PFANN_OK;
    SearchPlan p;
    PF_TRY(plan_search(db, Q, k, &p));
    if (db->n == 0) {
        if (topk_out) PF_CUDA(cudaMemsetAsync(topk_out, 0, sizeof(uint32_t) * (size_t)Q * k, db->ctx->stream));
        fill_f32_kernel<<<cdiv(Q, 256), 256, 0, db->ctx->stream>>>(thr, -INFINITY, Q);   // an empty shard bounds nothing
        db->ctx->launches++;
        PF_CUDA(cudaGetLastError());
        return PFANN_OK;
    }
    for (int64_t g = 0; g < p.ngroups; g++) {
        const int64_t q0 = g * p.sgroup;
        const int Qs = (int)((Q - q0) < p.sgroup ? (Q - q0) : p.sgroup);
        PF_TRY(prepass(db, p, q + q0 * db->d, Qs, k, thr + q0, db->qnorm.as<float>() + q0, topk_out ? topk_out + q0 * k : nullptr));
    }
    return PFANN_OK;
}

// phase 2: filtered scans with the given thresholds (modified in place when a candidate list overflows) -> exact
// top-k of the rows that pass, as sortable keys.  Shards whose rows do not reach the global thresholds return fewer
// than k entries (0 keys) -- the merge only needs every member of the GLOBAL top-k, and those always pass.
int db_search_filtered_dev(Db *db, const float *q, int64_t Q, int k, float *thr, unsigned long long *keys, bool defer) {
    cudaStream_t st = db->ctx->stream;
    if (Q == 0) return PFANN_OK;
    SearchPlan p;
    PF_TRY(plan_search(db, Q, k, &p));
    PF_TRY(db->dist.ensure((size_t)Q * k * 4));
    PF_TRY(db->labels.ensure((size_t)Q * k * 8));
    float *dist = db->dist.as<float>();
    int64_t *labels = db->labels.as<int64_t>();
    if (db->n == 0) {
        PF_CUDA(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * (size_t)Q * k, st));
        return PFANN_OK;
    }
    int *flags = db->flags.as<int>();
    PF_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)(p.ngroups + 4), st));
    for (int64_t g = 0; g < p.ngroups; g++) {
        const int64_t q0 = g * p.sgroup;
        const int Qs = (int)((Q - q0) < p.sgroup ? (Q - q0) : p.sgroup);
        PF_TRY(filtered(db, p, q + q0 * db->d, Qs, k, thr + q0, db->qnorm.as<float>() + q0, dist + q0 * k, labels + q0 * k,
                        flags + g));
    }
    if (defer) {
        // no read-back here: the caller asks pfann_db_take_overflow() once after many calls and repeats them in the
        // checked mode if any candidate list overflowed (rare: thresholds are sized for ~256 survivors of 4096 slots)
        if (db->ovf_host == nullptr) {
            PF_CUDA(cudaHostAlloc(&db->ovf_host, sizeof(int), cudaHostAllocMapped));
            *db->ovf_host = 0;
            PF_CUDA(cudaHostGetDevicePointer(&db->ovf_dev, db->ovf_host, 0));
        }
        accumulate_flags_kernel<<<1, 1, 0, st>>>(flags, (int)p.ngroups, db->ovf_dev);
        db->ctx->launches++;
    } else {
        PF_TRY(redo_overflowed(db, p, q, Q, k, thr, dist, labels));
    }
    pack_keys_kernel<<<cdiv(Q * k, 256), 256, 0, st>>>(dist, labels, Q * k, keys);
    db->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

}  // namespace pfann

namespace {
int db_fill(Db *db, const float *emb, const int32_t *landmark_key);  // device allocations + copies of pfann_db_open
}

extern "C" {

int pfann_db_open(pfann_ctx *hctx, const float *emb, int64_t n, int d, const int32_t *landmark_key, int n_songs,
                  int64_t id_base, int64_t song_base, pfann_db **out) {
    PF_CHECK(hctx && out && n >= 0 && d > 0 && n_songs >= 0, PFANN_ERR_ARG, "pfann_db_open: bad argument");
    PF_CHECK(n == 0 || emb, PFANN_ERR_ARG, "pfann_db_open: emb is NULL");
    PF_CHECK(n_songs == 0 || landmark_key, PFANN_ERR_ARG, "pfann_db_open: landmark_key is NULL");
    PF_CHECK(d % 4 == 0 && d <= 1024, PFANN_ERR_UNSUPPORTED, "pfann_db_open: d=%d must be a multiple of 4, <= 1024", d);
    PF_CHECK(n < 0xFFFFFFF0LL, PFANN_ERR_UNSUPPORTED, "pfann_db_open: shard too large (%lld rows); shard it",
             (long long)n);
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    PF_CUDA(cudaSetDevice(ctx->device));
    Db *db = new Db();
    db->ctx = ctx;
    db->n = n; db->d = d; db->n_songs = n_songs; db->id_base = id_base; db->song_base = song_base;
    const int rc_fill = db_fill(db, emb, landmark_key);
    if (rc_fill != PFANN_OK) {  // nothing of a half-built shard survives an error
        pfann_db_close(reinterpret_cast<pfann_db *>(db));
        return rc_fill;
    }
    *out = reinterpret_cast<pfann_db *>(db);
    return PFANN_OK;
}

}  // extern "C"

namespace {
int db_fill(Db *db, const float *emb, const int32_t *landmark_key) {
    Ctx *ctx = db->ctx;
    const int64_t n = db->n, id_base = db->id_base;
    const int d = db->d, n_songs = db->n_songs;
    db->song_pos_host.resize((size_t)n_songs + 1);
    db->song_pos_host[0] = id_base;
    for (int i = 0; i < n_songs; i++) {  // database.py:83-86: cumulative sum with a leading zero
        PF_CHECK(landmark_key[i] >= 0, PFANN_ERR_ARG, "pfann_db_open: negative landmarkKey entry");
        db->song_pos_host[i + 1] = db->song_pos_host[i] + landmark_key[i];
    }
    PF_CHECK(db->song_pos_host[n_songs] - id_base == n || n_songs == 0, PFANN_ERR_ARG,
             "pfann_db_open: landmarkKey sums to %lld rows but %lld embeddings were given",
             (long long)(db->song_pos_host[n_songs] - id_base), (long long)n);
    PF_CUDA(cudaMalloc(&db->song_pos, sizeof(int64_t) * ((size_t)n_songs + 1)));
    PF_CUDA(cudaMemcpy(db->song_pos, db->song_pos_host.data(), sizeof(int64_t) * ((size_t)n_songs + 1),
                       cudaMemcpyHostToDevice));
    const size_t ne = (size_t)n * d;
    PF_CUDA(cudaMalloc(&db->emb32, sizeof(float) * (ne ? ne : 1)));
    PF_CUDA(cudaMalloc(&db->emb16, sizeof(__nv_bfloat16) * (ne ? ne : 1)));
    if (n > 0) {
        // on the context's stream: a device-resident `emb` produced by stream-ordered work is read in order
        PF_CUDA(cudaMemcpyAsync(db->emb32, emb, sizeof(float) * ne, cudaMemcpyDefault, ctx->stream));
        to_bf16_kernel<<<cdiv((long long)ne, 256), 256, 0, ctx->stream>>>(db->emb32, db->emb16, (int64_t)ne);
        unsigned int *dmax;
        PF_CUDA(cudaMalloc(&dmax, sizeof(unsigned int)));
        PF_CUDA(cudaMemsetAsync(dmax, 0, sizeof(unsigned int), ctx->stream));
        row_norm_max_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(db->emb32, n, d, dmax);
        ctx->launches += 2;
        unsigned int hm = 0;
        PF_CUDA(cudaMemcpyAsync(&hm, dmax, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        const cudaError_t e_sync = cudaStreamSynchronize(ctx->stream);
        cudaFree(dmax);
        PF_CUDA(e_sync);
        memcpy(&db->max_norm, &hm, sizeof(float));
    }
    return knn_tc_prepare(db);
}
}  // namespace

extern "C" {

void pfann_db_close(pfann_db *h) {
    Db *db = reinterpret_cast<Db *>(h);
    if (!db) return;
    cudaSetDevice(db->ctx->device);
    knn_tc_release(db);
    cudaFree(db->emb32);
    cudaFree(db->emb16);
    cudaFree(db->song_pos);
    DevBuf *bufs[] = {&db->qbuf, &db->qnorm, &db->thr, &db->cnt, &db->cand, &db->cand_v, &db->sample, &db->flags, &db->dist,
                      &db->labels, &db->rr_keys, &db->rr_scores, &db->rr_out, &db->lab_stage, &db->rr_xscores, &db->rr_done};
    for (DevBuf *b : bufs) b->release();
    if (db->ovf_host) cudaFreeHost(db->ovf_host);
    delete db;
}

int64_t pfann_db_ntotal(pfann_db *h) { return h ? reinterpret_cast<Db *>(h)->n : 0; }

int pfann_db_set_tuning(pfann_db *h, int cand_cap, int sample_rows, int use_tc) {
    PF_CHECK(h, PFANN_ERR_ARG, "pfann_db_set_tuning: NULL db");
    Db *db = reinterpret_cast<Db *>(h);
    if (cand_cap > 0) {
        PF_CHECK((cand_cap & (cand_cap - 1)) == 0 && cand_cap <= 4096 && cand_cap >= 2, PFANN_ERR_ARG,
                 "pfann_db_set_tuning: cand_cap must be a power of two in [2, 4096]");
        db->cand_cap = cand_cap;
    }
    if (sample_rows > 0) db->sample_rows = sample_rows;
    if (use_tc >= 0) db->use_tc = use_tc;
    return PFANN_OK;
}

int pfann_db_search(pfann_db *h, const float *q, int64_t Q, int k, float *dist, int64_t *labels) {
    PF_CHECK(h && Q >= 0 && k > 0 && (Q == 0 || (q && dist && labels)), PFANN_ERR_ARG, "pfann_db_search: bad argument");
    Db *db = reinterpret_cast<Db *>(h);
    PF_CHECK(k <= db->cand_cap && k <= 2048, PFANN_ERR_UNSUPPORTED, "pfann_db_search: k=%d too large (max %d)", k,
             db->cand_cap < 2048 ? db->cand_cap : 2048);
    if (Q == 0) return PFANN_OK;
    PF_CUDA(cudaSetDevice(db->ctx->device));
    const void *qd;
    void *dd, *ld;
    PF_TRY(stage_input(db->ctx, 0, q, (size_t)Q * db->d * 4, &qd));
    PF_TRY(stage_output(db->ctx, 0, dist, (size_t)Q * k * 4, &dd));
    PF_TRY(stage_output(db->ctx, 1, labels, (size_t)Q * k * 8, &ld));
    PF_TRY(db_search_dev(db, (const float *)qd, Q, k, (float *)dd, (int64_t *)ld));
    PF_TRY(finish_output(db->ctx, 0, dist, (size_t)Q * k * 4));
    return finish_output(db->ctx, 1, labels, (size_t)Q * k * 8);
}

int pfann_topk_merge(pfann_ctx *hctx, const float *dist_g, const int64_t *labels_g, int G, int64_t Q, int k,
                     float *dist, int64_t *labels) {
    PF_CHECK(hctx && G > 0 && Q >= 0 && k > 0, PFANN_ERR_ARG, "pfann_topk_merge: bad argument");
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    if (Q == 0) return PFANN_OK;
    PF_CUDA(cudaSetDevice(ctx->device));
    int P = 1;
    while (P < G * k) P <<= 1;
    PF_CHECK(P <= 8192, PFANN_ERR_UNSUPPORTED, "pfann_topk_merge: G*k too large");
    const size_t gb = (size_t)G * Q * k;
    const void *dg, *lg;
    void *dd, *ld;
    PF_TRY(stage_input(ctx, 0, dist_g, gb * 4, &dg));
    PF_TRY(stage_input(ctx, 1, labels_g, gb * 8, &lg));
    PF_TRY(stage_output(ctx, 0, dist, (size_t)Q * k * 4, &dd));
    PF_TRY(stage_output(ctx, 1, labels, (size_t)Q * k * 8, &ld));
    const size_t smem = (size_t)(P + (P & 1)) * 4 + (size_t)P * 8;
    PF_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_kernel<<<(unsigned)Q, 256, smem, ctx->stream>>>((const float *)dg, (const int64_t *)lg, G, Q, k, P,
                                                                (float *)dd, (int64_t *)ld);
    ctx->launches++;
    PF_CUDA(cudaGetLastError());
    PF_TRY(finish_output(ctx, 0, dist, (size_t)Q * k * 4));
    return finish_output(ctx, 1, labels, (size_t)Q * k * 8);
}

/* ---- sharded search in two phases + key merge (pfann_b200/dist.py); DEVICE pointers, stream-ordered ---- */
int pfann_db_search_thresholds(pfann_db *h, const float *q, int64_t Q, int k, float *thr) {
    PF_CHECK(h && Q >= 0 && k > 0 && (Q == 0 || (q && thr)), PFANN_ERR_ARG, "pfann_db_search_thresholds: bad argument");
    Db *db = reinterpret_cast<Db *>(h);
    PF_CHECK(Q == 0 || (is_device_ptr(q) && is_device_ptr(thr)), PFANN_ERR_ARG,
             "pfann_db_search_thresholds: q and thr must be device pointers");
    PF_CUDA(cudaSetDevice(db->ctx->device));
    return db_search_thresholds_dev(db, q, Q, k, thr);
}

int pfann_db_search_sample_topk(pfann_db *h, const float *q, int64_t Q, int k, float *thr, uint32_t *topk) {
    PF_CHECK(h && Q >= 0 && k > 0 && (Q == 0 || (q && thr && topk)), PFANN_ERR_ARG, "pfann_db_search_sample_topk: bad argument");
    Db *db = reinterpret_cast<Db *>(h);
    PF_CHECK(Q == 0 || (is_device_ptr(q) && is_device_ptr(thr) && is_device_ptr(topk)), PFANN_ERR_ARG,
             "pfann_db_search_sample_topk: q, thr and topk must be device pointers");
    PF_CUDA(cudaSetDevice(db->ctx->device));
    return db_search_thresholds_dev(db, q, Q, k, thr, topk);
}

int pfann_db_thresholds_from_topk(pfann_db *h, const uint32_t *gathered, int world, const float *q, int64_t Q, int k, float *thr) {
    PF_CHECK(h && world > 0 && Q >= 0 && k > 0 && (Q == 0 || (gathered && q && thr)), PFANN_ERR_ARG,
             "pfann_db_thresholds_from_topk: bad argument");
    if (Q == 0) return PFANN_OK;
    Db *db = reinterpret_cast<Db *>(h);
    PF_CHECK(is_device_ptr(gathered) && is_device_ptr(q) && is_device_ptr(thr), PFANN_ERR_ARG,
             "pfann_db_thresholds_from_topk: device pointers only");
    PF_CUDA(cudaSetDevice(db->ctx->device));
    SearchPlan p;
    PF_TRY(plan_search(db, Q, k, &p));
    int P2 = 1;
    while (P2 < world * k) P2 <<= 1;
    PF_CHECK((size_t)P2 * 4 <= 160 * 1024, PFANN_ERR_UNSUPPORTED, "pfann_db_thresholds_from_topk: world * k = %d too large", world * k);
    PF_CUDA(cudaFuncSetAttribute(thr_union_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2 * 4));
    ProfScope ps(db->ctx, K_KNN_SELECT, 37);
    // the gathered scores may come from another shard's bf16 scan whatever this shard runs: always the larger bound
    const float eps_union = 4.0e-3f + fmaxf(1.0e-5f, 1.2e-7f * (float)db->d);
    thr_union_kernel<<<(unsigned)Q, 256, (size_t)P2 * 4, db->ctx->stream>>>(gathered, world, Q, P2, q, db->d, k, eps_union,
                                                                           db->max_norm, thr);
    db->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

int pfann_db_search_filtered(pfann_db *h, const float *q, int64_t Q, int k, float *thr, uint64_t *keys, int defer_overflow_check) {
    PF_CHECK(h && Q >= 0 && k > 0 && (Q == 0 || (q && thr && keys)), PFANN_ERR_ARG, "pfann_db_search_filtered: bad argument");
    Db *db = reinterpret_cast<Db *>(h);
    PF_CHECK(Q == 0 || (is_device_ptr(q) && is_device_ptr(thr) && is_device_ptr(keys)), PFANN_ERR_ARG,
             "pfann_db_search_filtered: q, thr and keys must be device pointers");
    PF_CHECK(db->id_base + db->n <= 0xFFFFFFFFLL, PFANN_ERR_UNSUPPORTED,
             "pfann_db_search_filtered: global row ids beyond 2^32 do not fit the packed keys");
    PF_CUDA(cudaSetDevice(db->ctx->device));
    return db_search_filtered_dev(db, q, Q, k, thr, reinterpret_cast<unsigned long long *>(keys), defer_overflow_check != 0);
}

int pfann_db_take_overflow(pfann_db *h) {
    Db *db = reinterpret_cast<Db *>(h);
    if (!db || !db->ovf_host) return 0;
    const int v = *reinterpret_cast<volatile int *>(db->ovf_host);
    *db->ovf_host = 0;
    return v;
}

int pfann_topk_merge_keys(pfann_ctx *hctx, const uint64_t *keys_g, int G, int64_t Q, int k, float *dist, int64_t *labels) {
    PF_CHECK(hctx && G > 0 && Q >= 0 && k > 0 && (Q == 0 || (keys_g && labels)), PFANN_ERR_ARG,
             "pfann_topk_merge_keys: bad argument");
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    if (Q == 0) return PFANN_OK;
    PF_CHECK(is_device_ptr(keys_g) && is_device_ptr(labels) && (!dist || is_device_ptr(dist)), PFANN_ERR_ARG,
             "pfann_topk_merge_keys: device pointers only");
    PF_CUDA(cudaSetDevice(ctx->device));
    int P = 1;
    while (P < G * k) P <<= 1;
    PF_CHECK(P <= 8192, PFANN_ERR_UNSUPPORTED, "pfann_topk_merge_keys: G*k too large");
    PF_CUDA(cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P * 8));
    merge_keys_kernel<<<(unsigned)Q, P >= 512 ? 256 : 64, (size_t)P * 8, ctx->stream>>>(
        reinterpret_cast<const unsigned long long *>(keys_g), G, Q, k, P, dist, labels);
    ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

int pfann_best_combine(pfann_ctx *hctx, const float *packed_g, int G, int nq, float *packed_out) {
    PF_CHECK(hctx && G > 0 && nq >= 0 && (nq == 0 || (packed_g && packed_out)), PFANN_ERR_ARG, "pfann_best_combine: bad argument");
    Ctx *ctx = reinterpret_cast<Ctx *>(hctx);
    if (nq == 0) return PFANN_OK;
    PF_CHECK(is_device_ptr(packed_g) && is_device_ptr(packed_out), PFANN_ERR_ARG, "pfann_best_combine: device pointers only");
    PF_CUDA(cudaSetDevice(ctx->device));
    combine_best_kernel<<<cdiv(nq, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const float4 *>(packed_g), G, nq,
                                                                reinterpret_cast<float4 *>(packed_out));
    ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

int pfann_db_set_sample_scale(pfann_db *h, float scale) {
    PF_CHECK(h && scale > 0.f && scale <= 4.f, PFANN_ERR_ARG, "pfann_db_set_sample_scale: scale must be in (0, 4]");
    reinterpret_cast<Db *>(h)->sample_scale = scale;
    return PFANN_OK;
}

float pfann_db_max_norm(pfann_db *h) { return h ? reinterpret_cast<Db *>(h)->max_norm : 0.f; }

int pfann_db_set_max_norm(pfann_db *h, float max_norm) {
    PF_CHECK(h && max_norm >= 0.f, PFANN_ERR_ARG, "pfann_db_set_max_norm: bad argument");
    Db *db = reinterpret_cast<Db *>(h);
    PF_CHECK(max_norm >= db->max_norm, PFANN_ERR_ARG, "pfann_db_set_max_norm: %g is below this shard's own maximum %g",
             (double)max_norm, (double)db->max_norm);
    db->max_norm = max_norm;
    return PFANN_OK;
}

}  // extern "C"
