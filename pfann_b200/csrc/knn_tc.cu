// knn_tc.cu -- stage 3a scan on tcgen05 tensor cores: scores = DB_tile[128 x d] (bf16) . Q^T[d x N] (bf16).
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0   TMA producer: streams 128-row database tiles (SWIZZLE_128B, K-major) through a 4-stage
//            shared-memory ring -- the whole kernel is a single pass over the shard, i.e. HBM-bound when few
//            queries are searched per pass (the per-query-file regime of matcher.py:136) and tensor-bound when
//            hundreds are batched;
//   warp 1   MMA issuer: d/16 tcgen05.mma per tile into one of two TMEM accumulator buffers (128 lanes = the
//            tile's database rows, N columns = the queries, which stay resident in shared memory);
//   warps 2-5 epilogue: tcgen05.ld their lane quarter, compare against the per-query threshold and append the
//            few survivors to the candidate lists (mode 1), or dump the scores of the sample pre-pass (mode 0).
// The bf16 scores only SELECT candidates; knn.cu re-scores them exactly in fp32.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "db.cuh"
#include "pfann_b200.h"

using namespace pfann;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int KNN_EPI_WARPS = 16;  // four per TMEM lane quarter, 32 query columns each
constexpr int KNN_THREADS = 32 * (2 + KNN_EPI_WARPS);

struct KnnTcState {
    CUtensorMap mapA;  // [n][d] bf16, box (64, 128)
};

struct ScanArgs {
    const float *q;      // [Qtot][d] fp32; CTA row blockIdx.y serves queries [group * y, group * y + Qg)
    int Qtot, group;     // queries of the launch, queries per group (= per database pass)
    int Qg, d;           // queries of this CTA's group (set inside the kernel)
    long long r0, r1;    // row range
    int mode;
    float *sample;       // mode 0: [Qg][sample_ld]
    long long sample_ld;
    const float *thr;    // mode 1
    int *cnt;
    uint32_t *cand;      // [Qg][cap] row ids of the filter survivors ...
    uint32_t *cand_v;    // ... and their (scan score - threshold) as fp32 bits: the select kernel ranks by it
    int cap;
    unsigned long long *prof;  // optional [grid][8] cycle counters (tools/knn_probe.py)
    int debug;           // probe knob (PFANN_KNN_DEBUG): 1 skip filter, 2 skip TMEM loads too, 3 also skip the MMAs
};

// shared-memory budget: N = 256 queries per pass take 64 KB (queries) + 32 KB (threshold tile), leaving three
// 32 KB database stages and 2560 parked survivors; up to 128 queries: four stages, 2048 survivors
template <int N> struct ScanCfg {
    static constexpr int STAGES = N > 128 ? 3 : 4;
    static constexpr int QCAP = N > 128 ? 2560 : 2048;   // parked filter survivors per CTA
    static constexpr int TROWS = N > 128 ? N : 128;      // rows of the threshold tile
};

// gridDim.y query groups share one launch: group y is scanned by the gridDim.x CTAs of its row, so the launch
// gaps, kernel prologues (query staging, TMEM allocation) and tails of up to four database passes are paid once.
template <int N, bool SAMPLE>
__global__ void __launch_bounds__(KNN_THREADS, 1) knn_scan_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                     const ScanArgs a_in) {
    ScanArgs a = a_in;
    {
        const long long g0 = (long long)blockIdx.y * a.group;
        a.Qg = (int)((a.Qtot - g0) < a.group ? (a.Qtot - g0) : a.group);
        a.q += g0 * a.d;
        if (a.sample) a.sample += g0 * a.sample_ld;
        if (a.thr) a.thr += g0;
        if (a.cnt) a.cnt += g0;
        if (a.cand) a.cand += g0 * a.cap;
        if (a.cand_v) a.cand_v += g0 * a.cap;
        if (a.prof) a.prof += (size_t)blockIdx.y * gridDim.x * 8;
    }
    constexpr int STAGES = ScanCfg<N>::STAGES, QCAP = ScanCfg<N>::QCAP, TROWS = ScanCfg<N>::TROWS;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sbase = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int KBLK = a.d / BK;                         // K blocks per tile (d = 128 -> 2)
    const uint32_t A_KB_BYTES = BM * BK * 2;           // 16 KB per K block
    const uint32_t STAGE_BYTES = A_KB_BYTES * KBLK;
    const uint32_t B_KB_BYTES = N * BK * 2;
    unsigned char *sB = sbase;                         // [KBLK][N rows][128 B]
    unsigned char *sA = sbase + (size_t)B_KB_BYTES * KBLK;  // [STAGES][KBLK][128 rows][128 B]
    // survivors of the filter are parked here and pushed to the global candidate lists after the scan: a
    // global atomicAdd with return (~1 us) must not sit between the TMEM read and the buffer release
    uint2 *queue = reinterpret_cast<uint2 *>(sA + (size_t)STAGES * STAGE_BYTES + (size_t)TROWS * 128);  // [QCAP] (query, row)
    uint32_t *queue_v = reinterpret_cast<uint32_t *>(queue + QCAP);                         // [QCAP] score - threshold
    // threshold tile: row r holds [1 1 1 0..] in K columns 0-15 and [-thr_hi -thr_mid -thr_lo 0..] of query r in K
    // columns 16-31.  One extra MMA per tile (A = columns 0-15, B = columns 16-31 of the same rows) starts the
    // accumulator at -threshold, so "score reaches its threshold" is just a clear sign bit in TMEM.
    unsigned char *cthr = sA + (size_t)STAGES * STAGE_BYTES;
    __shared__ int qcount_s;
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ntiles = (a.r1 - a.r0 + BM - 1) / BM;

    if (tid == 0) {
        qcount_s = 0;
        ptx::prefetch_tmap(&mapA);
        for (int s = 0; s < STAGES; s++) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], N >= 128 ? KNN_EPI_WARPS : 4 * (N / 32));  // one arrive per participating warp
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&tmem_base_s, 2 * N < 32 ? 32 : 2 * N);
        ptx::tmem_relinquish();
    }
    // queries -> bf16, K-major, 128-byte swizzle (the layout TMA SWIZZLE_128B would have produced)
    for (int i = tid; i < N * (a.d / 8); i += KNN_THREADS) {
        const int n = i / (a.d / 8), ch = i - n * (a.d / 8);  // 16-byte chunk = 8 consecutive k
        const int kb = ch >> 3, c = ch & 7;
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (n < a.Qg) {
            const float4 *src = reinterpret_cast<const float4 *>(a.q + (long long)n * a.d + ch * 8);  // d % 64 == 0
            const float4 lo = __ldg(src), hi = __ldg(src + 1);
            __nv_bfloat162 p0 = __floats2bfloat162_rn(lo.x, lo.y), p1 = __floats2bfloat162_rn(lo.z, lo.w);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(hi.x, hi.y), p3 = __floats2bfloat162_rn(hi.z, hi.w);
            pk.x = *reinterpret_cast<uint32_t *>(&p0);
            pk.y = *reinterpret_cast<uint32_t *>(&p1);
            pk.z = *reinterpret_cast<uint32_t *>(&p2);
            pk.w = *reinterpret_cast<uint32_t *>(&p3);
        }
        *reinterpret_cast<uint4 *>(sB + (size_t)kb * B_KB_BYTES + (size_t)n * 128 + ((c ^ (n & 7)) << 4)) = pk;
    }
    if (tid < TROWS) {
        uint32_t w0 = 0u, w1 = 0u;  // (-thr) as three bf16 terms; queries beyond Qg: -inf (can never be reached)
        if (a.mode == 1) {
            if (tid < a.Qg) {
                // a hair below the threshold so that the 3-term bf16 representation can only admit MORE rows
                const float t = -(a.thr[tid] - 1e-6f * fabsf(a.thr[tid]) - 1e-30f);
                const __nv_bfloat16 h = __float2bfloat16_rn(t);
                const __nv_bfloat16 m = __float2bfloat16_rn(t - __bfloat162float(h));
                const __nv_bfloat16 l = __float2bfloat16_rn(t - __bfloat162float(h) - __bfloat162float(m));
                w0 = (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(m) << 16);
                w1 = (uint32_t)__bfloat16_as_ushort(l);
                if (!(fabsf(t) < 3.0e38f)) w0 = t > 0.f ? 0x7F80u : 0xFF80u, w1 = 0u;  // infinite threshold
            } else {
                w0 = 0xFF80u;
            }
        }
        const uint32_t one = 0x3F80u;
        uint4 *row = reinterpret_cast<uint4 *>(cthr + (size_t)tid * 128);
        const int x = tid & 7;  // 128-byte swizzle: logical 16-byte chunk c lives at chunk c ^ (row % 8)
        row[0 ^ x] = make_uint4(one | (one << 16), one, 0u, 0u);
        row[1 ^ x] = make_uint4(0u, 0u, 0u, 0u);
        row[2 ^ x] = make_uint4(w0, w1, 0u, 0u);
        row[3 ^ x] = make_uint4(0u, 0u, 0u, 0u);
    }
    ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    long long pc[5] = {0, 0, 0, 0, 0};  // cycle accounting: producer wait, mma wait tempty/full, epilogue wait/work

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            long long it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
                const int s = (int)(it % STAGES);
                const long long c0 = clock64();
                if (it >= STAGES) ptx::mbar_wait(&empty_bar[s], (uint32_t)((it / STAGES) - 1) & 1);
                pc[0] += clock64() - c0;
                ptx::mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                const int row0 = (int)(a.r0 + tile * BM);
                for (int kb = 0; kb < KBLK; kb++)
                    ptx::tma_load_2d(sA + (size_t)s * STAGE_BYTES + (size_t)kb * A_KB_BYTES, &mapA, &full_bar[s], kb * BK,
                                     row0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, N);
            long long it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
                const int s = (int)(it % STAGES), buf = (int)(it & 1);
                const long long c0 = clock64();
                if (it >= 2) ptx::mbar_wait(&tempty_bar[buf], (uint32_t)((it >> 1) - 1) & 1);
                const long long c1 = clock64();
                ptx::mbar_wait(&full_bar[s], (uint32_t)(it / STAGES) & 1);
                pc[1] += c1 - c0;
                pc[2] += clock64() - c1;
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N);
                {   // accumulator := -threshold (zero in the sample pre-pass)
                    const uint64_t dc = ptx::umma_desc_k_sw128(ptx::smem_u32(cthr));
                    ptx::umma_f16(d_tmem, dc, dc + 2, idesc, 0);
                }
                for (int kb = 0; kb < KBLK && a.debug < 3; kb++) {
                    const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(sA + (size_t)s * STAGE_BYTES + (size_t)kb * A_KB_BYTES));
                    const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(sB + (size_t)kb * B_KB_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)
                        ptx::umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1);
                }
                ptx::umma_commit(&empty_bar[s]);
                ptx::umma_commit(&tfull_bar[buf]);
            }
        }
    } else {
        // ===== epilogue: 16 warps, four per TMEM lane quarter (= warp % 4); warp `wc` of a quarter owns query columns
        // [32 wc, 32 wc + 32).  Warps whose columns do not exist (N = 32) only take part in the final flush. =====
        const int quarter = warp & 3, wc = (warp - 2) >> 2;
        const uint32_t t_quarter = tmem_base + ((uint32_t)(quarter * 32) << 16);
        long long it = 0;
        // sample pre-pass: every thread keeps, per query column, the maximum over the rows it has seen.  The k-th
        // largest of these maxima (each over a distinct subset of the sample) is a lower bound of the sample's k-th
        // largest score -- 148 x 128 values per query leave the kernel instead of one per sampled row.
        float mx[SAMPLE ? (N > 128 ? 64 : 32) : 1];
#pragma unroll
        for (int i = 0; i < (SAMPLE ? (N > 128 ? 64 : 32) : 1); i++) mx[i] = -INFINITY;
        if (wc * 32 < N) {
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
                const int buf = (int)(it & 1);
                const long long c0 = clock64();
                ptx::mbar_wait(&tfull_bar[buf], (uint32_t)(it >> 1) & 1);
                ptx::tc_fence_after();
                const long long c1 = clock64();
                pc[3] += c1 - c0;
                const long long row = a.r0 + tile * BM + quarter * 32 + lane;
                const bool rvalid = row < a.r1;
#pragma unroll
              for (int ci = 0; ci < (N > 128 ? 2 : 1); ci++) {   // N = 256: two 32-column chunks per warp
                const int c = wc * 32 + ci * 128;
                uint32_t v[32];
                if (a.debug < 2) {
                    ptx::tmem_ld_32x32b_x32(t_quarter + (uint32_t)(buf * N + c), v);
                    ptx::tmem_ld_wait();
                }
                if (ci == (N > 128 ? 1 : 0)) {
                    // the last values are in registers: hand the buffer back to the MMA issuer right away
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty_bar[buf]);
                }
                if (a.debug >= 2) {
                } else if (a.debug >= 1) {
                    uint32_t x = 0;
#pragma unroll
                    for (int i = 0; i < 32; i++) x ^= v[i];
                    if (x == 0x12345678u && a.cnt) a.cnt[0] = 1;  // keep the load alive
                } else if (SAMPLE) {
                    if (rvalid) {
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            mx[(SAMPLE ? ci * 32 : 0) + (SAMPLE ? i : 0)] =
                                fmaxf(mx[(SAMPLE ? ci * 32 : 0) + (SAMPLE ? i : 0)], __uint_as_float(v[i]));
                        }
                    }
                } else {
                    // accumulator = score - threshold: a hit is a clear sign bit.  AND-tree of the 32 words (3-input
                    // LOP3s), then the rare per-hit path.
                    uint32_t m8[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) m8[i] = v[4 * i] & v[4 * i + 1] & v[4 * i + 2] & v[4 * i + 3];
                    const uint32_t all = (m8[0] & m8[1] & m8[2] & m8[3]) & (m8[4] & m8[5] & m8[6] & m8[7]);
                    if (!(all & 0x80000000u) && rvalid) {
                        uint32_t mask = 0;
#pragma unroll
                        for (int i = 0; i < 32; i++) mask |= ((~v[i]) >> 31) << i;
                        while (mask) {
                            const int i = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const int qp = atomicAdd(&qcount_s, 1);
                            uint32_t vi = v[0];  // v[i] without dynamic register indexing
#pragma unroll
                            for (int j = 1; j < 32; j++) vi = (i == j) ? v[j] : vi;
                            if (qp < QCAP) {
                                queue[qp] = make_uint2((uint32_t)(c + i), (uint32_t)row);
                                queue_v[qp] = vi;
                            } else {  // queue full (degenerate thresholds): push directly
                                const int pos = atomicAdd(a.cnt + c + i, 1);
                                if (pos < a.cap) {
                                    a.cand[(long long)(c + i) * a.cap + pos] = (uint32_t)row;
                                    a.cand_v[(long long)(c + i) * a.cap + pos] = vi;
                                }
                            }
                        }
                    }
                }
              }
                pc[4] += clock64() - c1;
            }
            if (SAMPLE) {
#pragma unroll
                for (int ci = 0; ci < (N > 128 ? 2 : 1); ci++)
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        const int c = wc * 32 + ci * 128;
                        if (c + i < a.Qg)
                            a.sample[(long long)(c + i) * a.sample_ld + (long long)blockIdx.x * BM + quarter * 32 + lane] =
                                mx[(SAMPLE ? ci * 32 : 0) + (SAMPLE ? i : 0)];
                    }
            }
        }
        // flush the parked survivors: the 128 epilogue threads issue their global atomics side by side
        asm volatile("bar.sync 1, %0;" ::"n"(32 * KNN_EPI_WARPS) : "memory");
        if (a.mode == 1) {
            const int nq = qcount_s < QCAP ? qcount_s : QCAP;
            for (int i = tid - 64; i < nq; i += 32 * KNN_EPI_WARPS) {
                const uint2 e = queue[i];
                const int pos = atomicAdd(a.cnt + e.x, 1);
                if (pos < a.cap) {
                    a.cand[(long long)e.x * a.cap + pos] = e.y;
                    a.cand_v[(long long)e.x * a.cap + pos] = queue_v[i];
                }
            }
        }
    }
    if (a.prof && lane == 0) {
        unsigned long long *o = a.prof + (size_t)blockIdx.x * 8;
        if (warp == 0) o[0] = pc[0];
        if (warp == 1) { o[1] = pc[1]; o[2] = pc[2]; }
        if (warp >= 2) { atomicAdd(&o[3], (unsigned long long)pc[3]); atomicAdd(&o[4], (unsigned long long)pc[4]); }
        if (warp == 2) o[5] = (unsigned long long)ntiles;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 2 * N < 32 ? 32 : 2 * N);
    }
}

template <int N, bool SAMPLE>
int launch_scan(Db *db, KnnTcState *st, const ScanArgs &a) {
    const int KBLK = db->d / BK;
    const size_t smem = 1024 + (size_t)N * BK * 2 * KBLK + (size_t)ScanCfg<N>::STAGES * BM * BK * 2 * KBLK +
                        (size_t)ScanCfg<N>::QCAP * 12 + (size_t)ScanCfg<N>::TROWS * 128;
    PF_CUDA(cudaFuncSetAttribute(knn_scan_tc_kernel<N, SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = (a.r1 - a.r0 + BM - 1) / BM;
    const int ngroups = (a.Qtot + a.group - 1) / a.group;
    long long grid = db->ctx->sm_count / ngroups;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) return PFANN_OK;
    ProfScope ps(db->ctx, K_KNN_SCAN, a.mode == 0 ? 35 : 36);
    knn_scan_tc_kernel<N, SAMPLE><<<dim3((unsigned)grid, (unsigned)ngroups), KNN_THREADS, smem, db->ctx->stream>>>(st->mapA, a);
    db->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

}  // namespace

namespace pfann {

int knn_tc_prepare(Db *db) {
    db->tc_state = nullptr;
    if (db->n == 0 || db->d % BK != 0 || db->d > 256) return PFANN_OK;  // CUDA-core scan serves these
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    PF_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, PFANN_ERR_CUDA,
             "cuTensorMapEncodeTiled is not available from the driver");
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    KnnTcState *st = new KnnTcState();
    cuuint64_t dims[2] = {(cuuint64_t)db->d, (cuuint64_t)db->n};
    cuuint64_t str[1] = {(cuuint64_t)db->d * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&st->mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db->emb16, dims, str, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        delete st;
        set_error("knn_tc_prepare: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return PFANN_ERR_CUDA;
    }
    db->tc_state = st;
    return PFANN_OK;
}

void knn_tc_release(Db *db) {
    delete reinterpret_cast<KnnTcState *>(db->tc_state);
    db->tc_state = nullptr;
}

// sample pre-pass (mode 0): number of per-thread maxima written per query for a row range
int64_t knn_tc_sample_slots(Db *db, int64_t r0, int64_t r1, int ngroups) {
    const long long ntiles = (r1 - r0 + BM - 1) / BM;
    long long grid = db->ctx->sm_count / (ngroups > 0 ? ngroups : 1);
    if (grid > ntiles) grid = ntiles;
    return grid * BM;
}

int knn_tc_scan(Db *db, const float *q, int Qtot, int group, int64_t r0, int64_t r1, int mode, float *sample,
                int64_t sample_ld, const float *thr, int *cnt, uint32_t *cand, uint32_t *cand_v, int cap) {
    KnnTcState *st = reinterpret_cast<KnnTcState *>(db->tc_state);
    PF_CHECK(st != nullptr, PFANN_ERR_STATE, "knn_tc_scan: tensor-core state missing");
    PF_CHECK(group >= 1 && group <= 256 && Qtot >= 1 && (Qtot + group - 1) / group <= 8, PFANN_ERR_ARG,
             "knn_tc_scan: 1..256 queries per pass, at most 8 passes per launch");
    const int Qg = Qtot < group ? Qtot : group;   // widest group of the launch picks the instantiation
    ScanArgs a;
    a.q = q; a.Qtot = Qtot; a.group = group; a.Qg = Qg; a.d = db->d; a.r0 = r0; a.r1 = r1; a.mode = mode;
    a.sample = sample; a.sample_ld = sample_ld; a.thr = thr; a.cnt = cnt; a.cand = cand; a.cand_v = cand_v; a.cap = cap;
    const char *dbg = getenv("PFANN_KNN_DEBUG");
    a.debug = dbg ? atoi(dbg) : 0;
    const char *pp = getenv("PFANN_KNN_PROF_PTR");  // probe only: device address of a zeroed [grid][8] u64 buffer
    a.prof = pp ? reinterpret_cast<unsigned long long *>(strtoull(pp, nullptr, 0)) : nullptr;
    if (mode == 0) {
        if (Qg <= 32) return launch_scan<32, true>(db, st, a);
        if (Qg <= 128 || db->d > 128) return launch_scan<128, true>(db, st, a);
        return launch_scan<256, true>(db, st, a);
    }
    if (Qg <= 32) return launch_scan<32, false>(db, st, a);
    if (Qg <= 128 || db->d > 128) return launch_scan<128, false>(db, st, a);
    return launch_scan<256, false>(db, st, a);
}

}  // namespace pfann
