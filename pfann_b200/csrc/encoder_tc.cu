// encoder_tc.cu -- stage 2 convolutions as tcgen05 (5th-gen tensor core) implicit GEMMs.
//
// Every dense convolution of the encoder (model.py:20,26-27; all but layer-0 conv1 whose C_in is 1) is
//     Y[m][n] = sum_{tap j, channel c} X[b][fi(m,j)][ti(m,j)][c] * W[n][j*Ci + c] + bias[n]
// with bf16 operands and fp32 accumulation in TMEM.
//
//  * im2col is done by the TMA engine, not by threads: activations are channels-last bf16, so for one
//    kernel tap the 128 output positions of a tile are a strided 4-D box (c:64, to, f, b) of the input
//    tensor -- stride 2 along the convolved axis, base pointer shifted by the tap offset.  Positions that
//    fall into the TF-"same" zero padding (model.py:18-19,24-25) are out of bounds of the per-tap tensor
//    map and are zero-filled by the hardware.  Taps that only ever see padding are never issued.
//  * operands land in shared memory in the 128-byte swizzled K-major layout tcgen05.mma consumes; the kernel is
//    persistent and warp-specialised: a multi-stage mbarrier ring decouples the TMA producer (one thread)
//    from the MMA issuer (one thread) across tile boundaries;
//  * the accumulator tile (128 x BN fp32) lives in one of two TMEM buffers, so the four epilogue warps drain
//    tile i (tcgen05.ld, + bias, raw output staged through swizzled smem into coalesced 16-byte stores,
//    per-sample LayerNorm partial sums into fixed slots -- no atomics, bit-reproducible) while the tensor
//    core already accumulates tile i + 1.
// LayerNorm itself (global over (C,F,T) of a sample) is finished by ln_finalize_kernel and applied by
// ln_apply_kernel in encoder.cu.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "encoder.cuh"
#include "pfann_b200.h"

using namespace pfann;

namespace {

constexpr int BM = 128;          // rows per tile = TMEM lanes = UMMA M
constexpr int BK = 64;           // bf16 elements per K block = 128 bytes = one swizzle span
constexpr int TC_THREADS = 192;  // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue

struct TcConv {
    bool supported = false;
    CUtensorMap mapA[3];
    CUtensorMap mapB;
    int BN = 0, NT = 0;  // N tile, number of N tiles
    int box_to = 0, box_f = 0, box_b = 0, fdim = 0;
    int slots = 0;       // LayerNorm partial slots per sample
    const void *x_base = nullptr;  // activation buffer the A tensor maps point into
};

struct TcState {
    TcConv conv[16];
    int max_slots = 1;
    PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
};

struct TcArgs {
    void *Y;            // [M][Co], fp32 or bf16
    const float *bias;  // [Co]
    float2 *partials;   // [nb][slots]
    long long M;        // valid rows
    int m_tiles;        // ceil(M / 128)
    int Co, R;          // channels, rows per sample (Fo*To)
    int To, fdim;       // tile -> tensor coordinates
    int kb_per_tap, ntaps;
    int NT, slots;
};

template <int BN>
__host__ __device__ constexpr int tc_stages() {
    return BN >= 256 ? 4 : 6;
}

// Persistent: grid = #SMs, every CTA walks tiles t = blockIdx.x, +gridDim.x, ... of the (m_tile, n_tile) space.
//   warp 0      TMA producer  -- runs ahead of the MMAs by up to STAGES K-blocks, across tile boundaries
//   warp 1      MMA issuer    -- accumulates tile i into TMEM buffer i & 1 while the epilogue drains i - 1
//   warps 2..5  epilogue      -- TMEM -> registers -> (+bias, LayerNorm partial sums) -> swizzled smem ->
//                                coalesced 16-byte global stores of the raw output (bf16 or fp32)
template <int BN, typename YT>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0,
                                                                     const __grid_constant__ CUtensorMap mapA1,
                                                                     const __grid_constant__ CUtensorMap mapA2,
                                                                     const __grid_constant__ CUtensorMap mapB,
                                                                     const TcArgs a) {
    constexpr int STAGES = tc_stages<BN>();
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sbase =
        reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *stage_out = sbase + (size_t)STAGES * STAGE_BYTES;   // 4 warps x 4 KB epilogue staging
    float *bias_s = reinterpret_cast<float *>(stage_out + 4 * 4096);  // [Co]
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = a.kb_per_tap * a.ntaps;
    const long long ntiles = (long long)a.m_tiles * a.NT;

    if (tid == 0) {
        ptx::prefetch_tmap(&mapA0);
        ptx::prefetch_tmap(&mapB);
        for (int s = 0; s < STAGES; s++) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], 4);  // one arrive per epilogue warp
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&tmem_base_s, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    for (int i = tid; i < a.Co; i += TC_THREADS) bias_s[i] = a.bias[i];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            long long it = 0;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const long long m0 = (t / a.NT) * BM;
                const int n0 = (int)(t % a.NT) * BN;
                const int f0 = (int)((m0 / a.To) % a.fdim);
                const int b0 = (int)(m0 / ((long long)a.To * a.fdim));
                for (int kb = 0; kb < KB; kb++, it++) {
                    const int s = (int)(it % STAGES);
                    if (it >= STAGES) ptx::mbar_wait(&empty_bar[s], (uint32_t)((it / STAGES) - 1) & 1);
                    unsigned char *sa = sbase + (size_t)s * STAGE_BYTES;
                    ptx::mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    const int tap = kb / a.kb_per_tap, c0 = (kb - tap * a.kb_per_tap) * BK;
                    const CUtensorMap *mA = tap == 0 ? &mapA0 : (tap == 1 ? &mapA1 : &mapA2);
                    ptx::tma_load_4d(sa, mA, &full_bar[s], c0, 0, f0, b0);
                    ptx::tma_load_2d(sa + A_BYTES, &mapB, &full_bar[s], kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN);
            long long it = 0, ti = 0;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ti++) {
                const int buf = (int)(ti & 1);
                if (ti >= 2) ptx::mbar_wait(&tempty_bar[buf], (uint32_t)((ti >> 1) - 1) & 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < KB; kb++, it++) {
                    const int s = (int)(it % STAGES);
                    ptx::mbar_wait(&full_bar[s], (uint32_t)(it / STAGES) & 1);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(sbase + (size_t)s * STAGE_BYTES);
                    const uint64_t da = ptx::umma_desc_k_sw128(sa), db = ptx::umma_desc_k_sw128(sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)
                        ptx::umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    ptx::umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
                }
                ptx::umma_commit(&tfull_bar[buf]);    // accumulator of this tile complete
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        uint4 *stg = reinterpret_cast<uint4 *>(stage_out + (size_t)(warp - 2) * 4096);
        constexpr int CPR = 32 * (int)sizeof(YT) / 16;  // 16-byte chunks per staged row of 32 columns: 8 fp32 / 4 bf16
        constexpr int RPI = 32 / CPR;                   // rows per store instruction: 4 (fp32) or 8 (bf16)
        long long ti = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ti++) {
            const int buf = (int)(ti & 1);
            const long long m0 = (t / a.NT) * BM;
            const int n_tile = (int)(t % a.NT), n0 = n_tile * BN;
            ptx::mbar_wait(&tfull_bar[buf], (uint32_t)(ti >> 1) & 1);
            ptx::tc_fence_after();
            const long long mrow0 = m0 + quarter * 32;
            const long long m = mrow0 + lane;
            const bool valid = m < a.M;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + c), v);
                ptx::tmem_ld_wait();
                float o[32];
#pragma unroll
                for (int j = 0; j < 8; j++) {  // bias by vector loads into distinct registers (no LDS serialisation)
                    const float4 b4 = *reinterpret_cast<const float4 *>(&bias_s[n0 + c + 4 * j]);
                    o[4 * j] = __uint_as_float(v[4 * j]) + b4.x;
                    o[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
                    o[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
                    o[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
                }
                float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};  // 4 independent chains
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    p1[i & 3] += o[i];
                    p2[i & 3] = fmaf(o[i], o[i], p2[i & 3]);
                }
                s1 += (p1[0] + p1[1]) + (p1[2] + p1[3]);
                s2 += (p2[0] + p2[1]) + (p2[2] + p2[3]);
                // registers (row = lane) -> XOR-swizzled smem (conflict-free both ways) -> row-contiguous stores
                const int sw = sizeof(YT) == 4 ? (lane & 7) : ((lane >> 1) & 3);
                if (sizeof(YT) == 4) {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        stg[lane * CPR + (j ^ sw)] =
                            make_uint4(__float_as_uint(o[4 * j]), __float_as_uint(o[4 * j + 1]),
                                       __float_as_uint(o[4 * j + 2]), __float_as_uint(o[4 * j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        __nv_bfloat162 q0 = __floats2bfloat162_rn(o[8 * j], o[8 * j + 1]);
                        __nv_bfloat162 q1 = __floats2bfloat162_rn(o[8 * j + 2], o[8 * j + 3]);
                        __nv_bfloat162 q2 = __floats2bfloat162_rn(o[8 * j + 4], o[8 * j + 5]);
                        __nv_bfloat162 q3 = __floats2bfloat162_rn(o[8 * j + 6], o[8 * j + 7]);
                        stg[lane * CPR + (j ^ sw)] =
                            make_uint4(*reinterpret_cast<uint32_t *>(&q0), *reinterpret_cast<uint32_t *>(&q1),
                                       *reinterpret_cast<uint32_t *>(&q2), *reinterpret_cast<uint32_t *>(&q3));
                    }
                }
                __syncwarp();
#pragma unroll
                for (int r0 = 0; r0 < 32; r0 += RPI) {
                    const int r = r0 + lane / CPR, ch = lane % CPR;
                    const int rsw = sizeof(YT) == 4 ? (r & 7) : ((r >> 1) & 3);
                    const uint4 val = stg[r * CPR + (ch ^ rsw)];
                    if (mrow0 + r < a.M) {
                        unsigned char *dst = reinterpret_cast<unsigned char *>(a.Y) +
                                             ((mrow0 + r) * a.Co + n0 + c) * (long long)sizeof(YT) + ch * 16;
                        *reinterpret_cast<uint4 *>(dst) = val;
                    }
                }
                __syncwarp();
            }
            // accumulator fully read: hand the TMEM buffer back before the (cheap) statistics write
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[buf]);
            // per-sample LayerNorm partials in fixed slots (deterministic: no atomics)
            if (a.R >= 32) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if (lane == 0 && valid) {
                    const long long sample = m / a.R;
                    const int slot = (int)((m % a.R) >> 5) * a.NT + n_tile;
                    a.partials[sample * a.slots + slot] = make_float2(s1, s2);
                }
            } else {
                for (int o = a.R >> 1; o > 0; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if ((lane & (a.R - 1)) == 0 && valid) a.partials[(m / a.R) * a.slots + n_tile] = make_float2(s1, s2);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// stats[b] = (mean, rstd) from the partial slots; one warp per sample, fixed summation order
__global__ void ln_finalize_kernel(const float2 *partials, int slots, long long E, float2 *stats, int nb) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= nb) return;
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < slots; i += 32) {
        const float2 p = partials[(long long)b * slots + i];
        s1 += (double)p.x;
        s2 += (double)p.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
        const double mean = s1 / (double)E;
        double var = s2 / (double)E - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[b] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
    }
}

bool is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }

int encode_map(TcState *st, CUtensorMap *map, const void *base, int rank, const cuuint64_t *dims,
               const cuuint64_t *strides_bytes /* rank-1 */, const cuuint32_t *box) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = st->encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), dims,
                            strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]", (int)r,
                  rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                  box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return PFANN_ERR_CUDA;
    }
    return PFANN_OK;
}

template <int BN, typename YT>
int launch_tc(Model *m, const TcConv &tc, const TcArgs &args) {
    const size_t smem =
        (size_t)tc_stages<BN>() * (BM * BK * 2 + BN * BK * 2) + 4 * 4096 + (size_t)args.Co * 4 + 1024;
    PF_CHECK(smem + 2048 <= 227 * 1024, PFANN_ERR_UNSUPPORTED, "conv GEMM needs %zu B of shared memory", smem);
    static size_t attr_smem = 0;  // per instantiation: raise the opt-in limit only when a launch needs more
    if (smem > attr_smem) {
        PF_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, YT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        attr_smem = smem;
    }
    const long long ntiles = (long long)args.m_tiles * args.NT;
    long long grid = m->ctx->sm_count;
    if (grid > ntiles) grid = ntiles;
    ProfScope ps(m->ctx, K_CONV_TC);
    conv_gemm_tc_kernel<BN, YT><<<(unsigned)grid, TC_THREADS, smem, m->ctx->stream>>>(tc.mapA[0], tc.mapA[1],
                                                                                       tc.mapA[2], tc.mapB, args);
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

}  // namespace

namespace pfann {

bool tc_supported(const ConvGeom &g) {
    if (g.depthwise) return false;
    if (g.Ci % BK != 0 || g.Co % 64 != 0 || g.Co > 8192) return false;
    const long long R = g.rows_per_sample();
    if (!is_pow2(R) || !is_pow2(g.To) || g.To > BM) return false;
    if (!is_pow2(g.axis == 0 ? g.Fi : g.Fo)) return false;
    for (int j = 0; j < g.ntaps; j++)
        if (g.tap_off[j] < 0) return false;  // left padding would need negative box origins
    return true;
}

int tc_prepare(Model *m) {
    TcState *st = new TcState();
    m->tc_state = st;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    PF_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, PFANN_ERR_CUDA,
             "cuTensorMapEncodeTiled is not available from the driver");
    st->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    // workspaces must exist before tensor maps can point into them
    PF_TRY(plan_workspace(m));
    for (int i = 1; i < 16; i++) {
        const bool front = i < 2 * m->front_layers;
        const cuuint64_t nb = (cuuint64_t)(front ? m->front_sub : m->chunk);
        const ConvGeom &g = m->conv[i].g;
        TcConv &tc = st->conv[i];
        tc.supported = tc_supported(g);
        if (!tc.supported) continue;
        const __nv_bfloat16 *X = reinterpret_cast<const __nv_bfloat16 *>(
            front ? ((i & 1) == 0 ? m->fxb.p : m->fxa.p) : ((i & 1) == 0 ? m->xb.p : m->xa.p));
        tc.x_base = X;
        tc.BN = g.Co >= 256 ? 256 : (g.Co >= 128 ? 128 : 64);
        tc.NT = g.Co / tc.BN;
        tc.fdim = g.axis == 0 ? g.Fi : g.Fo;
        tc.box_to = g.To;
        tc.box_f = tc.fdim < BM / g.To ? tc.fdim : BM / g.To;
        tc.box_b = BM / (g.To * tc.box_f);
        const long long R = g.rows_per_sample();
        tc.slots = (int)((R >= 32 ? R / 32 : 1) * tc.NT);
        if (tc.slots > st->max_slots) st->max_slots = tc.slots;
        for (int j = 0; j < g.ntaps; j++) {
            const int off = g.tap_off[j];
            cuuint64_t dims[4], str[3];
            cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)tc.box_to, (cuuint32_t)tc.box_f, (cuuint32_t)tc.box_b};
            const __nv_bfloat16 *base;
            if (g.axis == 0) {
                // (c, to, f, b): X[b][f][2 to + off][c]
                const int nv = (g.Ti - off + 1) / 2;  // output positions whose tap is inside the input
                dims[0] = g.Ci; dims[1] = nv > 0 ? nv : 1; dims[2] = g.Fi; dims[3] = nb;
                str[0] = (cuuint64_t)2 * g.Ci * 2; str[1] = (cuuint64_t)g.Ti * g.Ci * 2;
                str[2] = (cuuint64_t)g.Fi * g.Ti * g.Ci * 2;
                base = X + (long long)off * g.Ci;
            } else {
                // (c, to, fo, b): X[b][2 fo + off][to][c]
                const int nv = (g.Fi - off + 1) / 2;
                dims[0] = g.Ci; dims[1] = g.Ti; dims[2] = nv > 0 ? nv : 1; dims[3] = nb;
                str[0] = (cuuint64_t)g.Ci * 2; str[1] = (cuuint64_t)2 * g.Ti * g.Ci * 2;
                str[2] = (cuuint64_t)g.Fi * g.Ti * g.Ci * 2;
                base = X + (long long)off * g.Ti * g.Ci;
            }
            PF_TRY(encode_map(st, &tc.mapA[j], base, 4, dims, str, box));
        }
        for (int j = g.ntaps; j < 3; j++) tc.mapA[j] = tc.mapA[0];
        {
            cuuint64_t dims[2] = {(cuuint64_t)g.K(), (cuuint64_t)g.Co};
            cuuint64_t str[1] = {(cuuint64_t)g.K() * 2};
            cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)tc.BN};
            PF_TRY(encode_map(st, &tc.mapB, m->conv[i].w_nk, 2, dims, str, box));
        }
    }
    PF_TRY(m->partials.ensure((size_t)m->chunk * st->max_slots * sizeof(float2)));
    if (m->front_layers > 0) PF_TRY(m->fpartials.ensure((size_t)m->front_sub * st->max_slots * sizeof(float2)));
    return PFANN_OK;
}

void tc_release(Model *m) {
    delete reinterpret_cast<TcState *>(m->tc_state);
    m->tc_state = nullptr;
}

int tc_conv(Model *m, int idx, const __nv_bfloat16 *X, void *Y, bool y_bf16, int nb) {
    TcState *st = reinterpret_cast<TcState *>(m->tc_state);
    PF_CHECK(st != nullptr, PFANN_ERR_STATE, "tc_conv: tensor-core state missing");
    const TcConv &tc = st->conv[idx];
    const ConvGeom &g = m->conv[idx].g;
    PF_CHECK(tc.supported, PFANN_ERR_UNSUPPORTED, "tc_conv: conv %d has no tensor-core geometry", idx);
    PF_CHECK(X == tc.x_base, PFANN_ERR_STATE, "tc_conv: input buffer moved since the tensor maps were built");
    TcArgs a;
    a.Y = Y; a.bias = m->conv[idx].bias; a.partials = m->cur_partials;
    a.M = (long long)nb * g.rows_per_sample();
    a.m_tiles = (int)((a.M + BM - 1) / BM);
    a.Co = g.Co; a.R = (int)g.rows_per_sample(); a.To = g.To; a.fdim = tc.fdim;
    a.kb_per_tap = g.Ci / BK; a.ntaps = g.ntaps; a.NT = tc.NT; a.slots = tc.slots;
    if (y_bf16) {
        if (tc.BN == 256) PF_TRY((launch_tc<256, __nv_bfloat16>(m, tc, a)));
        else if (tc.BN == 128) PF_TRY((launch_tc<128, __nv_bfloat16>(m, tc, a)));
        else PF_TRY((launch_tc<64, __nv_bfloat16>(m, tc, a)));
    } else {
        if (tc.BN == 256) PF_TRY((launch_tc<256, float>(m, tc, a)));
        else if (tc.BN == 128) PF_TRY((launch_tc<128, float>(m, tc, a)));
        else PF_TRY((launch_tc<64, float>(m, tc, a)));
    }
    ProfScope ps(m->ctx, K_LN);
    ln_finalize_kernel<<<cdiv(nb, 8), 256, 0, m->ctx->stream>>>(m->cur_partials, tc.slots, g.out_per_sample(),
                                                              m->cur_stats, nb);
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

}  // namespace pfann
