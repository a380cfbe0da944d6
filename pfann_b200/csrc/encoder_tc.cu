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
// Three kernels live here:
//   conv_gemm_tc_kernel   the plain implicit GEMM described above (raw output + LayerNorm partial sums; LayerNorm is
//                         finished by ln_finalize_kernel and applied by ln_apply_kernel in encoder.cu) -- used for the
//                         eight small tail convolutions;
//   conv_ln_tc_kernel     convolutions 1-7 with LayerNorm + ReLU inside the epilogue (statistics exchanged between
//                         CTAs through an L2 table), no raw output at all;
//   l0_tc_kernel          layer-0 conv1 (C_in = 1) + ln1 + ReLU as one K = 16 hi/lo-split MMA per 128 positions.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

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
    CUtensorMap mapB128;  // same weights, 128-row boxes (fused conv+LayerNorm kernel)
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
    int n_stages, b_res;  // smem ring depth; 1 = the whole weight matrix stays resident in shared memory
};


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
    constexpr int MAX_STAGES = 12;
    const int STAGES = a.n_stages;
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    // weights either stream through the ring next to A, or (when K x BN fits) are loaded once and stay resident:
    // that halves the L2 -> SM traffic of the small-K layers, which is what paces them (measured ~40 B/cycle/SM)
    const uint32_t STAGE_BYTES = A_BYTES + (a.b_res ? 0u : B_BYTES);
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sbase =
        reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int KBall = a.kb_per_tap * a.ntaps;
    unsigned char *sBres = sbase + (size_t)STAGES * STAGE_BYTES;       // [KB][BN x 128 B] when b_res
    unsigned char *stage_out = sBres + (a.b_res ? (size_t)KBall * B_BYTES : 0);  // 4 warps x 4 KB epilogue staging
    float *bias_s = reinterpret_cast<float *>(stage_out + 4 * 4096);  // [Co]
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], bres_bar, tfull_bar[2], tempty_bar[2];
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
        ptx::mbar_init(&bres_bar, 1);
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
            // ring position kept incrementally: a 64-bit divide by the run-time ring depth per K block would cost
            // more issue cycles than the four MMAs it feeds
            int s = 0;
            uint32_t round = 0;
            if (a.b_res) {  // the whole [Co x K] weight matrix, once
                ptx::mbar_expect_tx(&bres_bar, (uint32_t)KB * B_BYTES);
                for (int kb = 0; kb < KB; kb++) ptx::tma_load_2d(sBres + (size_t)kb * B_BYTES, &mapB, &bres_bar, kb * BK, 0);
            }
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const long long m0 = (t / a.NT) * BM;
                const int n0 = (int)(t % a.NT) * BN;
                const int f0 = (int)((m0 / a.To) % a.fdim);
                const int b0 = (int)(m0 / ((long long)a.To * a.fdim));
                int tap = 0, kbt = 0;
                for (int kb = 0; kb < KB; kb++) {
                    if (round > 0) ptx::mbar_wait(&empty_bar[s], (round - 1) & 1);
                    unsigned char *sa = sbase + (size_t)s * STAGE_BYTES;
                    ptx::mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    const CUtensorMap *mA = tap == 0 ? &mapA0 : (tap == 1 ? &mapA1 : &mapA2);
                    ptx::tma_load_4d(sa, mA, &full_bar[s], kbt * BK, 0, f0, b0);
                    if (!a.b_res) ptx::tma_load_2d(sa + A_BYTES, &mapB, &full_bar[s], kb * BK, n0);
                    if (++kbt == a.kb_per_tap) kbt = 0, tap++;
                    if (++s == STAGES) s = 0, round++;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN);
            long long ti = 0;
            int s = 0;
            uint32_t round = 0;
            if (a.b_res) ptx::mbar_wait(&bres_bar, 0);
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ti++) {
                const int buf = (int)(ti & 1);
                if (ti >= 2) ptx::mbar_wait(&tempty_bar[buf], (uint32_t)((ti >> 1) - 1) & 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < KB; kb++) {
                    ptx::mbar_wait(&full_bar[s], round & 1);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(sbase + (size_t)s * STAGE_BYTES);
                    const uint64_t da = ptx::umma_desc_k_sw128(sa);
                    const uint64_t db = ptx::umma_desc_k_sw128(a.b_res ? ptx::smem_u32(sBres + (size_t)kb * B_BYTES) : sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)
                        ptx::umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    ptx::umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
                    if (++s == STAGES) s = 0, round++;
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

// ------------------------------------------------------------------------------------------------------------
// Fused convolution + LayerNorm + ReLU: the raw convolution output never leaves the SM.
//
// Every CTA owns one fixed POSITION p = (row block rb of 128 output rows within a sample, 128-channel slice nh) and
// walks the samples (or, for samples shorter than 128 rows, groups of 128 / R samples) q = lane, lane + L, ... of
// its lane; the P = RB * NT CTAs of a lane together produce whole samples.  Because the position is fixed, the
// LayerNorm affine of the CTA (gamma, beta: 128 x 128 bf16 each, model.py:21,30) is loaded ONCE -- 32 registers per
// epilogue thread -- and for K <= 384 the weights are loaded once into shared memory: per tile only the im2col'd
// activations stream in (TMA) and the bf16 result streams out.
//
// Four 128 x 128 fp32 accumulators rotate through TMEM.  Per tile i the sixteen epilogue warps run
//   pass 1 (tile i)     tcgen05.ld, + bias, per-warp (sum, sum of squares); each warp publishes its pair as ONE
//                       64-bit word into a global exchange table (all-ones = not written yet);
//   pass 2 (tile i - 3) tcgen05.ld again, normalise with the sample statistics, affine, ReLU, bf16, one 32-byte
//                       sector store per thread and 16 columns (STG.256);
// one more "statistics" warp polls the table until the P * 16 words of a sample group are there, adds them in a fixed
// order (double) and hands (mean, rstd) to pass 2 through shared memory.  The exchange latency (a few microseconds
// through L2) is hidden behind three tiles of work, the tensor core runs one tile ahead of the epilogue, and no
// cluster launch is needed: the only requirement is that the <= 148 CTAs are co-resident (cooperative launch).
// A poll that does not complete (never observed; would mean a peer CTA is not running) times out, raises a flag
// and lets the kernel finish instead of hanging the GPU.
// ------------------------------------------------------------------------------------------------------------
struct TcLnArgs {
    __nv_bfloat16 *X;                    // [M][Co] normalised output
    const float *bias;                   // [Co]
    const __nv_bfloat16 *gb;             // [P][64 KB] gamma/beta per CTA position in the epilogue's shared-memory layout
    unsigned long long *part;            // [n_groups][P][16] packed (sum, sumsq) per epilogue warp; ~0 = not written
    int *err;                            // set to 1 when a poll timed out
    long long M;                         // valid rows
    int n_groups;                        // tiles per position
    int Co, R, To, fdim, kb_per_tap, ntaps;
    int NT, P, L;                        // channel slices, positions per sample group, lanes
    int sg_shift;                        // log2(samples per tile): 0, 1 or 2
    int n_stages, b_res;                 // smem ring depth; 1 = this CTA's weight slice stays resident in shared memory
    unsigned long long *prof;            // optional cycle counters [grid][8] (tools/ln_probe.py)
};

constexpr int LN_EPI_WARPS = 16;  // four per TMEM lane quarter, 32 columns each
constexpr int LN_THREADS = 32 * (3 + LN_EPI_WARPS);  // warp 0 TMA producer, 1 MMA issuer, 2-17 epilogue, 18 statistics
constexpr int LN_BN = 128;        // accumulator width: 4 x 128 columns = all of TMEM
constexpr int LN_DEFER = 3;       // pass 2 runs this many tiles behind pass 1 (4 TMEM slots: 3 waiting + 1 accumulating)
constexpr uint32_t LN_GB_BYTES = 128 * 128 * 2 * 2;
constexpr unsigned long long LN_UNSET = ~0ull;

__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <bool PROF>
__global__ void __launch_bounds__(LN_THREADS, 1) conv_ln_tc_kernel(const __grid_constant__ CUtensorMap mapA0,
                                                                   const __grid_constant__ CUtensorMap mapA1,
                                                                   const __grid_constant__ CUtensorMap mapA2,
                                                                   const __grid_constant__ CUtensorMap mapB,
                                                                   const TcLnArgs a) {
    constexpr int MAX_STAGES = 12;
    constexpr int BN = LN_BN;
    const int STAGES = a.n_stages;
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    const uint32_t STAGE_BYTES = A_BYTES + (a.b_res ? 0u : B_BYTES);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sbase =
        reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int KB = a.kb_per_tap * a.ntaps;
    unsigned char *sBres = sbase + (size_t)STAGES * STAGE_BYTES;                 // [KB][128 x 128 B] when b_res
    // bias tile: row r holds [1 1 1 0..] in K columns 0-15 and [b_hi b_mid b_lo 0..] of channel n0 + r in K columns
    // 16-31 (same swizzled K-major layout as the operands).  One extra MMA per tile, A = columns 0-15 of every row,
    // B = columns 16-31, initialises the accumulator with the bias -- the epilogue never touches it (bias reads
    // from shared memory used to cost more L1 data-pipe cycles than the tensor core's own operand reads).
    unsigned char *cb_s = sBres + (a.b_res ? (size_t)KB * B_BYTES : 0);
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], bres_bar;
    __shared__ __align__(8) uint64_t tfull_bar[4], tempty_bar[4], sready_bar[4];
    __shared__ float2 stat_s[4][4];                                               // [slot][sample of the tile] (mean, rstd)
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = (int)(blockIdx.x % (unsigned)a.P), lane_id = (int)(blockIdx.x / (unsigned)a.P);
    const int rb = p / a.NT, nh = p - rb * a.NT;
    const int n0 = nh * BN;
    const int GR = a.R > BM ? a.R : BM;                       // rows between consecutive tiles of one position
    const int nt_cta = lane_id < a.n_groups ? (a.n_groups - lane_id + a.L - 1) / a.L : 0;

    if (tid == 0) {
        ptx::prefetch_tmap(&mapA0);
        ptx::prefetch_tmap(&mapB);
        for (int s = 0; s < STAGES; s++) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 4; s++) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], LN_EPI_WARPS);  // one arrive per epilogue warp
            ptx::mbar_init(&sready_bar[s], 1);
        }
        ptx::mbar_init(&bres_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&tmem_base_s, 512);
        ptx::tmem_relinquish();
    }
    if (tid < BN) {
        const float b = a.bias[n0 + tid];
        const __nv_bfloat16 bh = __float2bfloat16_rn(b);
        const __nv_bfloat16 bm = __float2bfloat16_rn(b - __bfloat162float(bh));
        const __nv_bfloat16 bl = __float2bfloat16_rn(b - __bfloat162float(bh) - __bfloat162float(bm));
        const uint32_t one = 0x3F80u;  // bf16 1.0
        uint4 *row = reinterpret_cast<uint4 *>(cb_s + (size_t)tid * 128);
        const int x = tid & 7;         // 128-byte swizzle: logical 16-byte chunk c lives at chunk c ^ (row % 8)
        row[0 ^ x] = make_uint4(one | (one << 16), one, 0u, 0u);
        row[1 ^ x] = make_uint4(0u, 0u, 0u, 0u);
        row[2 ^ x] = make_uint4((uint32_t)__bfloat16_as_ushort(bh) | ((uint32_t)__bfloat16_as_ushort(bm) << 16),
                                (uint32_t)__bfloat16_as_ushort(bl), 0u, 0u);
        row[3 ^ x] = make_uint4(0u, 0u, 0u, 0u);
    }
    ptx::fence_proxy_async();  // the tensor core reads the bias tile through the async proxy
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            if (a.b_res) {  // this CTA's [128 x K] weight slice, once
                ptx::mbar_expect_tx(&bres_bar, (uint32_t)KB * B_BYTES);
                for (int kb = 0; kb < KB; kb++) ptx::tma_load_2d(sBres + (size_t)kb * B_BYTES, &mapB, &bres_bar, kb * BK, n0);
            }
            int s = 0;
            uint32_t round = 0;
            long long prod_wait = 0;
            for (int i = 0; i < nt_cta; i++) {
                const long long m0 = (long long)(lane_id + i * a.L) * GR + (long long)rb * BM;
                const int f0 = (int)((m0 / a.To) % a.fdim);
                const int b0 = (int)(m0 / ((long long)a.To * a.fdim));
                int tap = 0, kbt = 0;
                for (int kb = 0; kb < KB; kb++) {
                    const long long c0 = PROF ? clock64() : 0;
                    if (round > 0) ptx::mbar_wait(&empty_bar[s], (round - 1) & 1);
                    if (PROF) prod_wait += clock64() - c0;
                    unsigned char *sa = sbase + (size_t)s * STAGE_BYTES;
                    ptx::mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    const CUtensorMap *mA = tap == 0 ? &mapA0 : (tap == 1 ? &mapA1 : &mapA2);
                    ptx::tma_load_4d(sa, mA, &full_bar[s], kbt * BK, 0, f0, b0);
                    if (!a.b_res) ptx::tma_load_2d(sa + A_BYTES, &mapB, &full_bar[s], kb * BK, n0);
                    if (++kbt == a.kb_per_tap) kbt = 0, tap++;
                    if (++s == STAGES) s = 0, round++;
                }
            }
            if (PROF && a.prof) a.prof[(size_t)blockIdx.x * 8 + 7] = (unsigned long long)prod_wait;
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN);
            int s = 0;
            uint32_t round = 0;
            if (a.b_res) ptx::mbar_wait(&bres_bar, 0);
            long long mw_empty = 0, mw_full = 0;
            for (int i = 0; i < nt_cta; i++) {
                const int slot = i & 3;
                const long long c0 = PROF ? clock64() : 0;
                if (i >= 4) ptx::mbar_wait(&tempty_bar[slot], (uint32_t)((i >> 2) - 1) & 1);
                if (PROF) mw_empty += clock64() - c0;
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(slot * BN);
                {   // accumulator := bias (ones x bias rows)
                    const uint64_t dc = ptx::umma_desc_k_sw128(ptx::smem_u32(cb_s));
                    ptx::umma_f16(d_tmem, dc, dc + 2, idesc, 0);
                }
                for (int kb = 0; kb < KB; kb++) {
                    const long long c1 = PROF ? clock64() : 0;
                    ptx::mbar_wait(&full_bar[s], round & 1);
                    if (PROF) mw_full += clock64() - c1;
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(sbase + (size_t)s * STAGE_BYTES);
                    const uint64_t da = ptx::umma_desc_k_sw128(sa);
                    const uint64_t db = ptx::umma_desc_k_sw128(a.b_res ? ptx::smem_u32(sBres + (size_t)kb * B_BYTES) : sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)
                        ptx::umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1);
                    ptx::umma_commit(&empty_bar[s]);
                    if (++s == STAGES) s = 0, round++;
                }
                ptx::umma_commit(&tfull_bar[slot]);
            }
            if (PROF && a.prof) {
                a.prof[(size_t)blockIdx.x * 8 + 5] = (unsigned long long)mw_empty;
                a.prof[(size_t)blockIdx.x * 8 + 6] = (unsigned long long)mw_full;
            }
        }
    } else if (warp < 2 + LN_EPI_WARPS) {
        // ===== epilogue: 16 warps, four per TMEM lane quarter; warp `wc` of a quarter owns columns [32 wc, 32 wc + 32) =====
        const int quarter = warp & 3, ew = warp - 2, wc = ew >> 2;
        const int cc = wc * 32;                         // first column of this warp inside the CTA's 128-channel slice
        const int pub = wc * 4 + quarter;               // index of this warp's word in the exchange table
        const int sgq = quarter >> (2 - a.sg_shift);    // which sample of the tile this warp's 32 rows belong to
        // gamma/beta of (row = this lane, this warp's 32 columns) never change for a CTA position: 32 registers,
        // loaded once ([position][chunk][quarter][8 x 16 B][lane]: one 512-byte request per warp load)
        uint4 gq[4], bq[4];
        {
            const uint4 *gl = reinterpret_cast<const uint4 *>(a.gb) + (size_t)p * (LN_GB_BYTES / 16) +
                              ((wc * 4 + quarter) * 8) * 32 + lane;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                gq[q] = __ldg(gl + q * 32);
                bq[q] = __ldg(gl + (4 + q) * 32);
            }
        }
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cc;
        long long pc[4] = {0, 0, 0, 0};
        // Inside an iteration pass 2 (tile i - 3) runs BEFORE pass 1 (tile i): its TMEM slot is the one the tensor
        // core needs for tile i + 1, so it is released first, and the statistics exchange of a tile gets three tile
        // times to complete (the CTAs of a lane drift against each other; with two it cost ~20 % of the epilogue).
        const int p2_first = 1;
        for (int hs = 0; hs < 2 * (nt_cta + LN_DEFER); hs++) {
            const int i = hs >> 1;
            const bool do_p1 = ((hs & 1) ^ p2_first) == 0;
            if (do_p1 && i < nt_cta) {
                // ---------------- pass 1 (tile i): per-warp sums ----------------
                const int slot = i & 3;
                const long long tw = PROF ? clock64() : 0;
                ptx::mbar_wait(&tfull_bar[slot], (uint32_t)(i >> 2) & 1);
                ptx::tc_fence_after();
                const long long t0 = PROF ? clock64() : 0;
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(t_lane + (uint32_t)(slot * BN), v);
                ptx::tmem_ld_wait();
                float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const float o0 = __uint_as_float(v[4 * q]), o1 = __uint_as_float(v[4 * q + 1]);
                    const float o2 = __uint_as_float(v[4 * q + 2]), o3 = __uint_as_float(v[4 * q + 3]);
                    p1[0] += o0; p1[1] += o1; p1[2] += o2; p1[3] += o3;
                    p2[0] = fmaf(o0, o0, p2[0]); p2[1] = fmaf(o1, o1, p2[1]);
                    p2[2] = fmaf(o2, o2, p2[2]); p2[3] = fmaf(o3, o3, p2[3]);
                }
                float s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]);
                float s2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if (lane == 0) {
                    unsigned long long w = ((unsigned long long)__float_as_uint(s2) << 32) | (unsigned long long)__float_as_uint(s1);
                    if (w == LN_UNSET) w = 0x7FC000007FC00000ull;  // a (NaN, NaN) that is not the "unset" pattern
                    const long long q = (long long)lane_id + (long long)i * a.L;
                    st_relaxed_gpu_u64(a.part + (q * a.P + p) * LN_EPI_WARPS + pub, w);
                }
                if (PROF) pc[0] += t0 - tw, pc[1] += clock64() - t0;
            }
            if (!do_p1 && i >= LN_DEFER) {
                // ---------------- pass 2 (tile j): normalise + affine + ReLU + bf16 store ----------------
                const int j = i - LN_DEFER, slot = j & 3;
                const long long tw = PROF ? clock64() : 0;
                ptx::mbar_wait(&sready_bar[slot], (uint32_t)(j >> 2) & 1);
                const long long t0 = PROF ? clock64() : 0;
                const float2 st = stat_s[slot][sgq];
                const float rstd = st.y, nmr = -st.x * st.y;
                const long long mrow0 = (long long)(lane_id + j * a.L) * GR + (long long)rb * BM + quarter * 32;
                const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
                uint32_t va[16], vb[16];
                ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)(slot * BN), va);
                ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)(slot * BN + 16), vb);
                ptx::tmem_ld_wait();
                // the TMEM slot is free as soon as its values are in registers
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tempty_bar[slot]);
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    // 16 columns: normalise, affine (packed bf16 FMA), ReLU
                    uint32_t pk[8];
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const uint4 g4 = gq[2 * hh + q], b4 = bq[2 * hh + q];
                        const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
                        const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int c = 8 * q + 2 * e;
                            const float a0 = __uint_as_float(hh ? vb[c] : va[c]), a1 = __uint_as_float(hh ? vb[c + 1] : va[c + 1]);
                            const float x0 = fmaf(a0, rstd, nmr);
                            const float x1 = fmaf(a1, rstd, nmr);
                            __nv_bfloat162 y2 = __hfma2(__floats2bfloat162_rn(x0, x1),
                                                        *reinterpret_cast<const __nv_bfloat162 *>(&gw[e]),
                                                        *reinterpret_cast<const __nv_bfloat162 *>(&bw[e]));
                            y2 = __hmax2(y2, zero2);
                            pk[4 * q + e] = *reinterpret_cast<const uint32_t *>(&y2);
                        }
                    }
                    // one full 32-byte sector per thread (STG.256), straight from registers: staging the rows through
                    // shared memory for wider row pieces cost more L1-data-pipe wavefronts and two warp barriers
                    if (mrow0 + lane < a.M)
                        ptx::st_global_v8(reinterpret_cast<unsigned char *>(a.X) +
                                              ((mrow0 + lane) * a.Co + n0 + cc + hh * 16) * 2, pk);
                }
                if (PROF) pc[2] += t0 - tw, pc[3] += clock64() - t0;
            }
        }
        if (PROF && a.prof && lane == 0) {
            unsigned long long *o = a.prof + (size_t)blockIdx.x * 8;
            atomicAdd(&o[0], (unsigned long long)pc[0]);
            atomicAdd(&o[1], (unsigned long long)pc[1]);
            atomicAdd(&o[2], (unsigned long long)pc[2]);
            atomicAdd(&o[3], (unsigned long long)pc[3]);
            if (warp == 2) atomicAdd(&o[4], (unsigned long long)nt_cta);
        }
    } else {
        // ===== statistics warp: gather the P * 16 partial words of each sample group, fixed-order sum, (mean, rstd) =====
        const int n = a.P * LN_EPI_WARPS;                    // words per sample group (32 ... 256)
        const int cls = (lane & 3) >> (2 - a.sg_shift);      // sample (within the tile) of the words this lane reads
        const int n_sg = 1 << a.sg_shift;
        const double invE = 1.0 / ((double)a.R * (double)a.Co);
        bool dead = false;
        for (int i = 0; i < nt_cta; i++) {
            const int slot = i & 3;
            const long long q = (long long)lane_id + (long long)i * a.L;
            const unsigned long long *src = a.part + q * a.P * LN_EPI_WARPS;
            // word e of the group belongs to lane e % 32 (e % 4 == lane % 4: all words of a lane are of one sample);
            // the up to 8 loads of a lane are issued together and re-polled until every word has been written
            unsigned long long w[8];
#pragma unroll
            for (int k = 0; k < 8; k++) w[k] = LN_UNSET;
            for (int tries = 0; tries < (dead ? 1 : (1 << 21)); tries++) {
                bool all = true;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int e = lane + 32 * k;
                    if (e < n && w[k] == LN_UNSET) w[k] = ld_relaxed_gpu_u64(src + e);
                }
#pragma unroll
                for (int k = 0; k < 8; k++) all = all && (lane + 32 * k >= n || w[k] != LN_UNSET);
                if (__all_sync(0xffffffffu, all)) break;
                __nanosleep(32);
            }
            double d1 = 0.0, d2 = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (lane + 32 * k < n) {
                    if (w[k] == LN_UNSET) {
                        dead = true;
                        *a.err = 1;
                        w[k] = 0;
                    }
                    d1 += (double)__uint_as_float((uint32_t)w[k]);
                    d2 += (double)__uint_as_float((uint32_t)(w[k] >> 32));
                }
            }
            dead = __any_sync(0xffffffffu, dead);
            for (int sg = 0; sg < n_sg; sg++) {
                double t1 = cls == sg ? d1 : 0.0, t2 = cls == sg ? d2 : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
                    t2 += __shfl_xor_sync(0xffffffffu, t2, o);
                }
                if (lane == sg) {
                    const double mean = t1 * invE;
                    double var = t2 * invE - mean * mean;
                    if (var < 0.0) var = 0.0;
                    stat_s[slot][sg] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&sready_bar[slot]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Layer 0: conv1 (C_in = 1, three taps along time, model.py:20) + ln1 + ReLU on the tensor core.
//
// K = 3 is far too small for a GEMM -- but one 128 x 128 x 16 tcgen05.mma per 128 output positions costs 64 tensor
// cycles and replaces 3 FMAs + bias per output on the CUDA cores.  The 16 K-columns carry a bf16 hi/lo split of both
// operands so that the product is exact to ~2^-17 (fp32 accumulation in TMEM):
//     A row (position p): [mh0 mh1 mh2 | ml0 ml1 ml2 | mh0 mh1 mh2 | 1  1  1  0 ...]      m_j = mel under tap j
//     B row (channel  n): [wh0 wh1 wh2 | wh0 wh1 wh2 | wl0 wl1 wl2 | bh bm bl 0 ...]      bias as three bf16 terms
// Builder warps assemble the A tiles (two 16-byte stores per row, swizzled K-major layout) from the fp32 log-mel,
// one thread issues the MMAs into a ring of four TMEM accumulators, and sixteen epilogue warps read them back
// (tcgen05.ld), normalise with the statistics the moments kernel derived analytically, apply the per-element affine
// (register-resident per position block: tiles are walked position-block-major) and ReLU, and store bf16 rows.
// ------------------------------------------------------------------------------------------------------------
struct L0TcArgs {
    const float *mel;          // [nb][F][T]
    const float2 *stats;       // [nb] (mean, rstd)
    const __nv_bfloat16 *gb;   // [PB][64 KB] gamma/beta per block of 128 positions, epilogue register layout
    const uint4 *btile;        // 16 KB weight tile in the tensor core's shared-memory layout
    __nv_bfloat16 *X;          // [nb][P][128]
    int nb, F, T, To, ntaps, off[3];
    int PB;                    // position blocks per sample (P / 128)
    int direct_store;          // 1 (default): 32-byte stores straight from registers; 0: staged 64-byte row pieces
};
constexpr int L0_BUILD_WARPS = 2;
constexpr int L0_THREADS = 32 * (L0_BUILD_WARPS + 1 + 16);
constexpr int L0_STAGES = 8;

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&v);
}

__global__ void __launch_bounds__(L0_THREADS, 1) l0_tc_kernel(const L0TcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sbase =
        reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = sbase;                         // L0_STAGES x (128 rows x 128 B); 32 B per row are used
    unsigned char *sB = sA + L0_STAGES * 16384;        // 16 KB weight tile
    unsigned char *stage_out = sB + 16384;             // 16 warps x 2 KB store staging
    __shared__ __align__(8) uint64_t a_full[L0_STAGES], a_empty[L0_STAGES], tfull_bar[4], tempty_bar[4];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // tiles in position-block-major order t = pb * nb + s; every CTA takes one contiguous range
    const long long T_total = (long long)a.PB * a.nb;
    const long long t_begin = T_total * blockIdx.x / gridDim.x, t_end = T_total * (blockIdx.x + 1) / gridDim.x;
    const int nt = (int)(t_end - t_begin);
    const int pb_begin = (int)(t_begin / a.nb), s_begin = (int)(t_begin - (long long)pb_begin * a.nb);

    if (tid == 0) {
        for (int s = 0; s < L0_STAGES; s++) {
            ptx::mbar_init(&a_full[s], 32 * L0_BUILD_WARPS);
            ptx::mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < 4; s++) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], 16);
        }
        ptx::fence_mbar_init();
    }
    if (warp == L0_BUILD_WARPS) {
        ptx::tmem_alloc(&tmem_base_s, 512);
        ptx::tmem_relinquish();
    }
    for (int i = tid; i < 1024; i += L0_THREADS) reinterpret_cast<uint4 *>(sB)[i] = a.btile[i];
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < L0_BUILD_WARPS) {
        // ===== A-tile builders: thread bt assembles rows bt and bt + 64 of every tile =====
        int pb = pb_begin, s = s_begin, stage = 0;
        uint32_t round = 0;
        float mv[2][3];
        auto load_rows = [&](int pbi, int si, float(&out)[2][3]) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int p = pbi * 128 + tid + 64 * h;
                const int f = p / a.To, to = p - f * a.To;
                const float *row = a.mel + ((long long)si * a.F + f) * a.T;
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const int t = 2 * to + a.off[j];
                    out[h][j] = (j < a.ntaps && t >= 0 && t < a.T) ? __ldg(row + t) : 0.f;
                }
            }
        };
        if (nt > 0) load_rows(pb, s, mv);
        for (int i = 0; i < nt; i++) {
            float cur[2][3];
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int j = 0; j < 3; j++) cur[h][j] = mv[h][j];
            int pbn = pb, sn = s + 1;
            if (sn == a.nb) sn = 0, pbn++;
            if (i + 1 < nt) load_rows(pbn, sn, mv);  // next tile's inputs are in flight while this one is assembled
            if (round > 0) ptx::mbar_wait(&a_empty[stage], (round - 1) & 1);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = tid + 64 * h;
                float mh[3], ml[3];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    mh[j] = __bfloat162float(__float2bfloat16_rn(cur[h][j]));
                    ml[j] = cur[h][j] - mh[j];
                }
                uint4 *row = reinterpret_cast<uint4 *>(sA + (size_t)stage * 16384 + (size_t)r * 128);
                const int x = r & 7;  // 128-byte swizzle: logical 16-byte chunk c lives at chunk c ^ (row % 8)
                row[0 ^ x] = make_uint4(pack_bf16(mh[0], mh[1]), pack_bf16(mh[2], ml[0]), pack_bf16(ml[1], ml[2]),
                                        pack_bf16(mh[0], mh[1]));
                row[1 ^ x] = make_uint4(pack_bf16(mh[2], 1.f), pack_bf16(1.f, 1.f), 0u, 0u);
            }
            ptx::fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
            ptx::mbar_arrive(&a_full[stage]);
            pb = pbn; s = sn;
            if (++stage == L0_STAGES) stage = 0, round++;
        }
    } else if (warp == L0_BUILD_WARPS) {
        if (lane == 0) {
            // ===== MMA issuer: one 128 x 128 x 16 MMA per tile =====
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, 128);
            const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(sB));
            int stage = 0;
            uint32_t round = 0;
            for (int i = 0; i < nt; i++) {
                const int slot = i & 3;
                if (i >= 4) ptx::mbar_wait(&tempty_bar[slot], (uint32_t)((i >> 2) - 1) & 1);
                ptx::mbar_wait(&a_full[stage], round & 1);
                ptx::tc_fence_after();
                const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(sA + (size_t)stage * 16384));
                ptx::umma_f16(tmem_base + (uint32_t)(slot * 128), da, db, idesc, 0);
                ptx::umma_commit(&a_empty[stage]);
                ptx::umma_commit(&tfull_bar[slot]);
                if (++stage == L0_STAGES) stage = 0, round++;
            }
        }
    } else {
        // ===== epilogue: 16 warps, four per TMEM lane quarter; warp `wc` of a quarter owns columns [32 wc, 32 wc + 32) =====
        const int quarter = warp & 3, ew = warp - (L0_BUILD_WARPS + 1), wc = ew >> 2;
        const int cc = wc * 32;
        uint4 *stg = reinterpret_cast<uint4 *>(stage_out + (size_t)ew * 2048);
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cc;
        const long long P = (long long)a.PB * 128;
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
        const int sw = (lane >> 1) & 3;
        uint4 gq[4], bq[4];
        int pb = pb_begin, s = s_begin, cur_pb = -1;
        float2 st_next = nt > 0 ? __ldg(a.stats + s) : make_float2(0.f, 1.f);
        for (int i = 0; i < nt; i++) {
            const int slot = i & 3;
            if (pb != cur_pb) {  // at most once inside a CTA's range: gamma/beta of this position block
                const uint4 *gl = reinterpret_cast<const uint4 *>(a.gb) + (size_t)pb * (LN_GB_BYTES / 16) +
                                  ((wc * 4 + quarter) * 8) * 32 + lane;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    gq[q] = __ldg(gl + q * 32);
                    bq[q] = __ldg(gl + (4 + q) * 32);
                }
                cur_pb = pb;
            }
            const float2 st = st_next;
            int pbn = pb, sn = s + 1;
            if (sn == a.nb) sn = 0, pbn++;
            if (i + 1 < nt) st_next = __ldg(a.stats + sn);
            const float rstd = st.y, nmr = -st.x * st.y;
            const long long mrow0 = (long long)s * P + (long long)pb * 128 + quarter * 32;
            ptx::mbar_wait(&tfull_bar[slot], (uint32_t)(i >> 2) & 1);
            ptx::tc_fence_after();
            uint32_t va[16], vb[16];
            ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)(slot * 128), va);
            ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)(slot * 128 + 16), vb);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[slot]);
            uint32_t pd[8];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint4 g4 = gq[q], b4 = bq[q];
                const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
                const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int c = 8 * (q & 1) + 2 * e;
                    const float a0 = __uint_as_float(q >= 2 ? vb[c] : va[c]), a1 = __uint_as_float(q >= 2 ? vb[c + 1] : va[c + 1]);
                    __nv_bfloat162 y2 = __hfma2(__floats2bfloat162_rn(fmaf(a0, rstd, nmr), fmaf(a1, rstd, nmr)),
                                                *reinterpret_cast<const __nv_bfloat162 *>(&gw[e]),
                                                *reinterpret_cast<const __nv_bfloat162 *>(&bw[e]));
                    y2 = __hmax2(y2, zero2);
                    pk[e] = *reinterpret_cast<const uint32_t *>(&y2);
                }
                if (a.direct_store) {
#pragma unroll
                    for (int e = 0; e < 4; e++) pd[4 * (q & 1) + e] = pk[e];
                    if (q & 1)
                        ptx::st_global_v8(reinterpret_cast<unsigned char *>(a.X) + ((mrow0 + lane) * 128 + cc + (q >> 1) * 16) * 2, pd);
                } else {
                    stg[lane * 4 + (q ^ sw)] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
            if (a.direct_store) {
                pb = pbn; s = sn;
                continue;
            }
            __syncwarp();
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 8) {
                const int r = r0 + (lane >> 2), ch = lane & 3;
                const uint4 val = stg[r * 4 + (ch ^ ((r >> 1) & 3))];
                *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(a.X) + ((mrow0 + r) * 128 + cc) * 2 + ch * 16) = val;
            }
            __syncwarp();
            pb = pbn; s = sn;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == L0_BUILD_WARPS) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
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

// Shared-memory plan: resident weights when K x BN x 2 B leaves room for >= 4 A stages, else A+B stream together.
static void plan_ring(int BN, int KB, int NT, size_t fixed_bytes, int *n_stages, int *b_res, size_t *smem) {
    const size_t budget = 225 * 1024 - 1024;  // dynamic smem available next to the static barriers
    const size_t A = (size_t)BM * BK * 2, B = (size_t)BN * BK * 2;
    const size_t bres = (size_t)KB * B;
    static const bool off = getenv("PFANN_B200_NO_BRES") != nullptr;
    if (!off && NT == 1 && fixed_bytes + bres + 4 * A <= budget) {
        size_t st = (budget - fixed_bytes - bres) / A;
        if (st > 12) st = 12;
        *n_stages = (int)st; *b_res = 1;
        *smem = st * A + bres + fixed_bytes + 1024;
        return;
    }
    size_t st = (budget - fixed_bytes) / (A + B);
    if (st > 12) st = 12;
    *n_stages = (int)st; *b_res = 0;
    *smem = st * (A + B) + fixed_bytes + 1024;
}

template <int BN, typename YT>
int launch_tc(Model *m, const TcConv &tc, const TcArgs &args_in) {
    TcArgs args = args_in;
    size_t smem = 0;
    plan_ring(BN, args.kb_per_tap * args.ntaps, args.NT, 4 * 4096 + (size_t)args.Co * 4, &args.n_stages, &args.b_res, &smem);
    PF_CHECK(args.n_stages >= 2, PFANN_ERR_UNSUPPORTED, "conv GEMM: no room for a shared-memory ring");
    // the opt-in is per device/context: set it on every launch (a process-wide cache skipped it on a second GPU)
    PF_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, YT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = (long long)args.m_tiles * args.NT;
    long long grid = m->ctx->sm_count;
    if (grid > ntiles) grid = ntiles;
    ProfScope ps(m->ctx, K_CONV_TC, m->prof_idx);
    conv_gemm_tc_kernel<BN, YT><<<(unsigned)grid, TC_THREADS, smem, m->ctx->stream>>>(tc.mapA[0], tc.mapA[1],
                                                                                       tc.mapA[2], tc.mapB, args);
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}


int launch_tc_ln(Model *m, const TcConv &tc, const LnGeom &lg, const TcLnArgs &args_in) {
    TcLnArgs args = args_in;
    // shared-memory plan: the 16 KB bias tile is fixed; the weight slice stays resident
    // when that leaves room for >= 3 activation stages, otherwise weights stream through the ring next to them
    const size_t budget = 225 * 1024, fixed = 16384;
    const size_t A = (size_t)BM * BK * 2, B = (size_t)LN_BN * BK * 2;
    const size_t bres = (size_t)args.kb_per_tap * args.ntaps * B;
    static const char *env_bres = getenv("PFANN_B200_LN_BRES");
    const bool want_bres = env_bres ? atoi(env_bres) != 0 : true;
    size_t st;
    if (want_bres && fixed + bres + 3 * A <= budget) {
        st = (budget - fixed - bres) / A;
        args.b_res = 1;
    } else {
        st = (budget - fixed) / (A + B);
        args.b_res = 0;
    }
    if (st > 12) st = 12;
    if (const char *e = getenv("PFANN_B200_LN_STAGES")) {  // experiment knob: shallower ring
        const size_t cap = (size_t)atoi(e);
        if (cap >= 2 && cap < st) st = cap;
    }
    PF_CHECK(st >= 2, PFANN_ERR_UNSUPPORTED, "fused conv+LN: no room for a shared-memory ring");
    args.n_stages = (int)st;
    const size_t smem = st * (A + (args.b_res ? 0 : B)) + (args.b_res ? bres : 0) + fixed + 1024;
    if (args.prof)
        PF_CUDA(cudaFuncSetAttribute(conv_ln_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
        PF_CUDA(cudaFuncSetAttribute(conv_ln_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // exchange table: all-ones = "not written"
    const size_t part_bytes = (size_t)args.n_groups * lg.P * LN_EPI_WARPS * sizeof(unsigned long long);
    PF_TRY(m->ln_part.ensure(part_bytes));
    PF_TRY(tc_ln_err_ptr(m, &args.err));
    PF_CUDA(cudaMemsetAsync(m->ln_part.p, 0xFF, part_bytes, m->ctx->stream));
    args.part = m->ln_part.as<unsigned long long>();
    int L = m->ctx->sm_count / lg.P;
    if (L > args.n_groups) L = args.n_groups;
    PF_CHECK(L >= 1, PFANN_ERR_UNSUPPORTED, "fused conv+LN: %d positions do not fit %d SMs", lg.P, m->ctx->sm_count);
    args.L = L;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(L * lg.P));
    cfg.blockDim = dim3(LN_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = m->ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;  // the CTAs of a lane wait for each other: all must be resident
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ProfScope ps(m->ctx, K_CONV_TC, m->prof_idx);
    if (args.prof)
        PF_CUDA(cudaLaunchKernelEx(&cfg, conv_ln_tc_kernel<true>, tc.mapA[0], tc.mapA[1], tc.mapA[2], tc.mapB128, args));
    else
        PF_CUDA(cudaLaunchKernelEx(&cfg, conv_ln_tc_kernel<false>, tc.mapA[0], tc.mapA[1], tc.mapA[2], tc.mapB128, args));
    m->ctx->launches++;
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
        const cuuint64_t nb = (cuuint64_t)m->chunk;
        const ConvGeom &g = m->conv[i].g;
        TcConv &tc = st->conv[i];
        tc.supported = tc_supported(g);
        if (!tc.supported) continue;
        const __nv_bfloat16 *X = reinterpret_cast<const __nv_bfloat16 *>(
            (i & 1) == 0 ? m->xb.p : m->xa.p);
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
            box[1] = (cuuint32_t)(g.Co < LN_BN ? g.Co : LN_BN);
            PF_TRY(encode_map(st, &tc.mapB128, m->conv[i].w_nk, 2, dims, str, box));
        }
    }
    PF_TRY(m->partials.ensure((size_t)m->chunk * st->max_slots * sizeof(float2)));
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
    ProfScope ps(m->ctx, K_LN, 16 + m->prof_idx);
    ln_finalize_kernel<<<cdiv(nb, 8), 256, 0, m->ctx->stream>>>(m->cur_partials, tc.slots, g.out_per_sample(),
                                                              m->cur_stats, nb);
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

LnGeom ln_geom(const ConvGeom &g) {
    LnGeom r;
    if (!tc_supported(g) || g.Co % LN_BN) return r;
    const long long R = g.rows_per_sample();
    if (R < 32) return r;                      // a warp's 32 accumulator rows must belong to one sample
    r.NT = g.Co / LN_BN;
    r.RB = R >= BM ? (int)(R / BM) : 1;
    r.sg_shift = R >= BM ? 0 : (R == 64 ? 1 : 2);
    r.P = r.RB * r.NT;
    if (r.P > 16) return r;                    // the statistics warp gathers P * 16 <= 256 words per sample group
    if (r.sg_shift > 0 && r.P * LN_EPI_WARPS > 32) return r;  // several samples per tile: one word per lane
    r.ok = true;
    return r;
}

bool tc_ln_supported(Model *m, int idx) {
    if (idx < 1 || idx > 14 || m->tc_state == nullptr) return false;
    if (getenv("PFANN_B200_NO_FUSED_LN") != nullptr) return false;  // read per call: tests toggle it
    return m->conv[idx].gb16 != nullptr && ln_geom(m->conv[idx].g).ok;
}

int tc_conv_ln(Model *m, int idx, const __nv_bfloat16 *X, __nv_bfloat16 *Xout, int nb) {
    TcState *st = reinterpret_cast<TcState *>(m->tc_state);
    PF_CHECK(st != nullptr, PFANN_ERR_STATE, "tc_conv_ln: tensor-core state missing");
    const TcConv &tc = st->conv[idx];
    const ConvGeom &g = m->conv[idx].g;
    const LnGeom lg = ln_geom(g);
    PF_CHECK(tc.supported && lg.ok, PFANN_ERR_UNSUPPORTED, "tc_conv_ln: conv %d has no fused geometry", idx);
    PF_CHECK(X == tc.x_base, PFANN_ERR_STATE, "tc_conv_ln: input buffer moved since the tensor maps were built");
    TcLnArgs a = {};
    a.X = Xout; a.bias = m->conv[idx].bias; a.gb = m->conv[idx].gb16;
    const long long R = g.rows_per_sample();
    a.M = (long long)nb * R;
    a.n_groups = R >= BM ? nb : (int)((a.M + BM - 1) / BM);
    a.Co = g.Co; a.R = (int)R; a.To = g.To; a.fdim = tc.fdim;
    a.kb_per_tap = g.Ci / BK; a.ntaps = g.ntaps; a.NT = lg.NT; a.P = lg.P; a.sg_shift = lg.sg_shift;
    const char *pp = getenv("PFANN_LN_PROF_PTR");  // probe only: zeroed device buffer [16][grid<=148][8] u64
    a.prof = pp ? reinterpret_cast<unsigned long long *>(strtoull(pp, nullptr, 0)) + (size_t)idx * 148 * 8 : nullptr;
    return launch_tc_ln(m, tc, lg, a);
}

bool tc_l0_supported(Model *m) {
    const bool off = getenv("PFANN_B200_NO_L0_TC") != nullptr;  // read per call: tests toggle it
    const ConvGeom &g = m->conv[0].g;
    return !off && m->l0_gb16 != nullptr && m->l0_btile != nullptr && g.Co == 128 && (g.Fo * g.To) % 128 == 0;
}

// layer-0 conv1 + ln1 + ReLU on the tensor core; `stats` = per-sample (mean, rstd) from the moments kernel
int tc_l0(Model *m, const float *mel, const float2 *stats, __nv_bfloat16 *X, int nb) {
    const ConvGeom &g = m->conv[0].g;
    L0TcArgs a = {};
    a.mel = mel; a.stats = stats; a.gb = m->l0_gb16; a.btile = reinterpret_cast<const uint4 *>(m->l0_btile); a.X = X;
    a.nb = nb; a.F = g.Fi; a.T = g.Ti; a.To = g.To; a.ntaps = g.ntaps;
    for (int j = 0; j < 3; j++) a.off[j] = j < g.ntaps ? g.tap_off[j] : 0;
    a.PB = g.Fo * g.To / 128;
    static const char *env_ds = getenv("PFANN_B200_L0_DIRECT_STORE");
    a.direct_store = env_ds ? atoi(env_ds) : 1;  // measured 2 % faster than the staged 64-byte row pieces
    const size_t smem = (size_t)L0_STAGES * 16384 + 16384 + 16 * 2048 + 1024;
    PF_CUDA(cudaFuncSetAttribute(l0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = m->ctx->sm_count;
    const long long tiles = (long long)a.PB * nb;
    if (grid > tiles) grid = tiles;
    ProfScope ps(m->ctx, K_CONV_TC, 0);
    l0_tc_kernel<<<(unsigned)grid, L0_THREADS, smem, m->ctx->stream>>>(a);
    m->ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFANN_OK;
}

int tc_ln_err_ptr(Model *m, int **dev_flag) {
    if (m->ln_err_host == nullptr) {
        PF_CUDA(cudaHostAlloc(&m->ln_err_host, sizeof(int), cudaHostAllocMapped));
        *m->ln_err_host = 0;
        PF_CUDA(cudaHostGetDevicePointer(&m->ln_err_dev, m->ln_err_host, 0));
    }
    *dev_flag = m->ln_err_dev;
    return PFANN_OK;
}

const CUtensorMap *tc_weight_map128(Model *m, int idx) {
    TcState *st = reinterpret_cast<TcState *>(m->tc_state);
    return (st != nullptr && st->conv[idx].supported) ? &st->conv[idx].mapB128 : nullptr;
}

int tc_ln_check(Model *m, bool sync) {
    if (m->ln_err_host == nullptr) return PFANN_OK;
    if (sync) PF_CUDA(cudaStreamSynchronize(m->ctx->stream));
    const int flag = *reinterpret_cast<volatile int *>(m->ln_err_host);
    if (flag != 0) {
        *m->ln_err_host = 0;  // reported once; later calls start clean
        set_error("fused conv+LayerNorm: statistics exchange timed out (CTAs not co-resident?); the embeddings of "
                  "the affected call are invalid");
        return PFANN_ERR_STATE;
    }
    return PFANN_OK;
}

}  // namespace pfann
