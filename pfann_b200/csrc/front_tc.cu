// front_tc.cu -- the first SeparableConv2d of the encoder (model.py:54-66, layer 0) as ONE kernel:
//
//     log-mel [B][F][T] fp32 --conv1 (C_in = 1, 1x3 along time, stride 2) + ln1 + ReLU--> X0 [B][F][T/2][C]
//                            --conv2 (3x1 along frequency, stride 2, dense)  + ln2 + ReLU--> X1 [B][F/2][T/2][C] bf16
//
// X0 (1 MB of bf16 per segment, the largest activation of the network) never exists in HBM: it is produced on the
// CUDA cores straight into the shared-memory operand tiles of conv2's tcgen05 GEMM.  Round 1 wrote it with one kernel
// (l0_tc_kernel) and read it back with the next (conv_ln_tc_kernel on conv 1): 48 % of the extraction step.
//
// Mapping.  One CTA per output frequency row fo of conv2 (grid = F/2 = 128 CTAs, all co-resident: cooperative
// launch); a tile is that row for a group of 8 segments:  M = 16 time steps x 8 segments = 128 accumulator rows
// (row m = to * 8 + s), N = 128 output channels, K = 3 taps x 128 channels.  Tap j of the tile is exactly X0 row
// f = 2 fo + j for the 8 segments, i.e. one 128 x 128 bf16 operand tile (two 16 KB K-blocks in the 128-byte-swizzled
// K-major layout); the 96 KB weight matrix stays resident in shared memory.  Because fo is fixed per CTA,
//   * the ln1 affine of the three X0 rows a CTA ever produces lives in registers (12 per thread),
//   * the ln2 affine of its output row lives in 8.5 KB of shared memory,
//   * only 3 KB of log-mel per tile come in and 32 KB of bf16 go out: HBM traffic 0.53 MB per segment instead of 2.6.
// The price: X0 rows with even f are produced twice (by fo = f/2 as tap 0 and by fo = f/2 - 1 as tap 2): 1.5 x the
// layer-0 arithmetic, which is 3 FMAs per element.
//
// Warp roles (19 warps):
//   warp 0        stager: log-mel values of the next tile (global -> registers -> shared memory, pre-scaled with the
//                 ln1 statistics the moments kernel derived analytically, encoder.cu l0_moments_kernel)
//   warp 1        one thread issues the tcgen05.mma (bias through an extra K = 16 MMA like conv_ln_tc_kernel)
//   warps 2-17    workers.  Per tile: produce the three X0 rows (thread = (time step, 4 channels), loop over segment
//                 pairs with packed fp32 FMAs), then the two LayerNorm passes over TMEM of older tiles:
//                 pass 1 (tile i-1) per-segment sum / sum of squares, pass 2 (tile i-3) normalise + affine + ReLU +
//                 bf16 + 32-byte stores
//   warp 18       statistics: adds the 16 worker-warp partials of a tile in a fixed order, publishes them with 64-bit
//                 integer atomics (fixed point: integer addition is associative, so the result does not depend on
//                 the order in which the 128 CTAs arrive -- bit-reproducible), waits for the arrival counter of the
//                 segment group and hands (mean, rstd) of the 8 segments to pass 2 through shared memory
#include <cuda.h>
#include <stdlib.h>

#include "encoder.cuh"
#include "pfann_b200.h"

using namespace pfann;

namespace pfann {
const CUtensorMap *tc_weight_map128(Model *m, int idx);  // encoder_tc.cu: [Co][K] bf16 weights, boxes of 64 x 128
}

namespace {

constexpr int FR_TO = 16;                  // time steps after conv1 (T = 32, stride 2)
constexpr int FR_S = 8;                    // segments per tile
constexpr int FR_C = 128;                  // channels of layer 0 (both convolutions)
constexpr int FR_PROD = 16;                // producer warps (one per time step)
constexpr int FR_EPI = 8;                  // epilogue warps (TMEM lane quarter x 64-column half)
constexpr int FR_W_PROD = 4, FR_W_EPI = FR_W_PROD + FR_PROD;   // first warp of each role (epilogue: multiple of 4)
constexpr int FR_THREADS = 32 * (FR_W_EPI + FR_EPI);
constexpr int FR_DEFER = 2;                // pass 2 runs this many tiles behind pass 1
constexpr uint32_t FR_UNIT = 128 * 128;    // one operand tile K-block: 128 rows x 128 bytes
constexpr int FR_GSTRIDE = 136;            // ln2 affine row stride in bf16 (272 B: the 4 rows a warp reads hit distinct banks)
constexpr double FR_FIX = 1048576.0;       // fixed-point scale of the exchanged sums (2^20)

struct FrontArgs {
    const float *mel;            // [nb][F][T]
    const float2 *stats0;        // [nb] (mean, rstd) of ln1 (from the mel moments)
    const float *w0;             // [C][3] conv1 weights (live taps, tap-major per channel)
    const float *b0;             // [C]
    const float *g0, *be0;       // ln1 affine, channels-last [F][To][C] fp32
    const float *b1;             // [C] conv2 bias
    const float *g1, *be1;       // ln2 affine, channels-last [Fo][To][C] fp32
    __nv_bfloat16 *X;            // [nb][Fo][To][C]
    unsigned long long *gsum;    // [n_groups][8][2] (fixed-point sum << 8 | arrivals), zeroed before the launch
    int *err;
    int nb, n_groups, F, T, Fo;
    unsigned long long *prof;    // optional cycle counters [grid][16] (tools/front_probe.py)
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&v);
}
// max(x * g + b, 0) on a bf16 pair: one instruction (HFMA2.BF16.RELU)
__device__ __forceinline__ uint32_t affine_relu(uint32_t x, uint32_t g, uint32_t b) {
    uint32_t y;
    asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(y) : "r"(x), "r"(g), "r"(b));
    return y;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
// Shared-memory matrix descriptor for a K-major bf16 tile of exactly one K = 16 step: rows of 32 bytes with the
// 32-byte swizzle (16-byte chunk c of row r lives at chunk c ^ ((r >> 2) & 1)), 8-row groups 256 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_k_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused: the tile is one swizzle span wide)
    d |= (uint64_t)(256 >> 4) << 32;              // stride byte offset
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)6 << 61;                       // layout: SWIZZLE_32B
    return d;
}
// explicit shared-space accesses (32-bit addresses): the carve-up of the dynamic buffer goes through generic pointers
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void red_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <bool PROF>
__global__ void __launch_bounds__(FR_THREADS, 1) front_tc_kernel(const __grid_constant__ CUtensorMap mapB, const FrontArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sbase =
        reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sW = sbase;                          // 6 K-blocks of conv2 weights: [tap][kb][128 n x 128 B]
    unsigned char *sA = sW + 6 * FR_UNIT;               // X0 operand tiles: [row slot r = tap][kb][128 m x 128 B]
    // conv2 bias through the tensor core: one extra K = 16 MMA per tile, A = rows of [1 1 1 0..], B = rows of
    // [b_hi b_mid b_lo 0..] (three bf16 terms of the fp32 bias), two compact 4 KB tiles in the 32-byte swizzle
    unsigned char *cb_s = sA + 6 * FR_UNIT;
    __nv_bfloat16 *g1_s = reinterpret_cast<__nv_bfloat16 *>(cb_s + 8192);        // [16][FR_GSTRIDE] ln2 gamma
    __nv_bfloat16 *be1_s = g1_s + FR_TO * FR_GSTRIDE;                           // [16][FR_GSTRIDE] ln2 beta
    float *P_s = reinterpret_cast<float *>(be1_s + FR_TO * FR_GSTRIDE);        // [2][3][16][4 segment pairs][8] pre-scaled mel
    float4 *AS_s = reinterpret_cast<float4 *>(P_s + 2 * 3 * FR_TO * 4 * 8);     // [2][4] (a_s, a_s+1, d_s, d_s+1)
    __shared__ __align__(8) uint64_t wres_bar, rfull[3], rempty[3], pfull[2], pempty[2];
    __shared__ __align__(8) uint64_t tfull[4], tempty[4], sready[4], p1done[2];
    __shared__ float2 part_s[2][FR_EPI][FR_S];          // [tile parity][epilogue warp][segment] (sum, sum of squares)
    __shared__ float2 stat_s[4][FR_S];                  // [slot][segment] (mean, rstd) of ln2
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fo = blockIdx.x;
    const int nrows = (2 * fo + 2 < a.F) ? 3 : 2;       // X0 rows 2 fo + j inside the input (the rest is zero padding)
    const int nt = a.n_groups;

    if (tid == 0) {
        ptx::prefetch_tmap(&mapB);
        for (int r = 0; r < 3; r++) {
            ptx::mbar_init(&rfull[r], FR_PROD);
            ptx::mbar_init(&rempty[r], 1);
        }
        for (int s = 0; s < 4; s++) {
            ptx::mbar_init(&tfull[s], 1);
            ptx::mbar_init(&tempty[s], FR_EPI);
            ptx::mbar_init(&sready[s], 1);
        }
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(&pfull[s], 1);
            ptx::mbar_init(&pempty[s], FR_PROD);
            ptx::mbar_init(&p1done[s], FR_EPI);
        }
        ptx::mbar_init(&wres_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&tmem_base_s, 512);
        ptx::tmem_relinquish();
    }
    if (tid < FR_C) {
        const float b = a.b1[tid];
        const __nv_bfloat16 bh = __float2bfloat16_rn(b);
        const __nv_bfloat16 bm = __float2bfloat16_rn(b - __bfloat162float(bh));
        const __nv_bfloat16 bl = __float2bfloat16_rn(b - __bfloat162float(bh) - __bfloat162float(bm));
        const uint32_t one = 0x3F80u;   // bf16 1.0
        const int x = (tid >> 2) & 1;
        uint4 *ra = reinterpret_cast<uint4 *>(cb_s + (size_t)tid * 32), *rb = reinterpret_cast<uint4 *>(cb_s + 4096 + (size_t)tid * 32);
        ra[0 ^ x] = make_uint4(one | (one << 16), one, 0u, 0u);
        ra[1 ^ x] = make_uint4(0u, 0u, 0u, 0u);
        rb[0 ^ x] = make_uint4((uint32_t)__bfloat16_as_ushort(bh) | ((uint32_t)__bfloat16_as_ushort(bm) << 16),
                               (uint32_t)__bfloat16_as_ushort(bl), 0u, 0u);
        rb[1 ^ x] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < FR_TO * FR_C; i += FR_THREADS) {   // ln2 affine of this CTA's output row
        const int to = i / FR_C, c = i - to * FR_C;
        const size_t src = ((size_t)fo * FR_TO + to) * FR_C + c;
        g1_s[to * FR_GSTRIDE + c] = __float2bfloat16_rn(a.g1[src]);
        be1_s[to * FR_GSTRIDE + c] = __float2bfloat16_rn(a.be1[src]);
    }
    ptx::fence_proxy_async();   // the tensor core reads the bias tiles through the async proxy
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== stager: entry e = (row r, time step to, segment s) -> the three mel values under the taps of conv1,
        // scaled by rstd_s; 12 entries per lane; two buffers, so that it runs a tile ahead of the producers =====
        if (lane == 0) {   // conv2 weights, once: K-block kb_all = tap * 2 + kb
            ptx::mbar_expect_tx(&wres_bar, 6 * FR_UNIT);
            for (int kb = 0; kb < 6; kb++) ptx::tma_load_2d(sW + (size_t)kb * FR_UNIT, &mapB, &wres_bar, kb * 64, 0);
        }
        float mv[12][3];
        float2 st = make_float2(0.f, 0.f);   // (a_s, d_s) of segment s = lane (lanes 0-7)
        auto fetch = [&](int g) {
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const int e = lane + 32 * k;
                const int s = e & 7, to = (e >> 3) & 15, r = e >> 7;
                const long long smp = (long long)g * FR_S + s;
                const bool ok = smp < a.nb && r < nrows;
                const float *row = a.mel + (smp * a.F + (2 * fo + r)) * a.T + 2 * to;
                mv[k][0] = ok ? __ldg(row) : 0.f;
                mv[k][1] = ok ? __ldg(row + 1) : 0.f;
                mv[k][2] = (ok && 2 * to + 2 < a.T) ? __ldg(row + 2) : 0.f;
            }
            if (lane < FR_S) {
                const long long smp = (long long)g * FR_S + lane;
                if (smp < a.nb) {
                    const float2 ms = __ldg(a.stats0 + smp);
                    st = make_float2(ms.y, -ms.x * ms.y);
                } else {
                    st = make_float2(0.f, 0.f);
                }
            }
        };
        if (nt > 0) fetch(0);
        for (int i = 0; i < nt; i++) {
            const int b = i & 1;
            if (i >= 2) ptx::mbar_wait(&pempty[b], (uint32_t)((i >> 1) - 1) & 1);
            float *Pb = P_s + b * (3 * FR_TO * 4 * 8);
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const int e = lane + 32 * k;
                const int s = e & 7, to = (e >> 3) & 15, r = e >> 7;
                const float as = __shfl_sync(0xffffffffu, st.x, s);
                float *dst = Pb + (((r * FR_TO + to) * 4 + (s >> 1)) * 8) + (s & 1);
                dst[0] = as * mv[k][0];
                dst[2] = as * mv[k][1];
                dst[4] = as * mv[k][2];
            }
            {
                const float a_hi = __shfl_down_sync(0xffffffffu, st.x, 1), d_hi = __shfl_down_sync(0xffffffffu, st.y, 1);
                if (lane < FR_S && (lane & 1) == 0) AS_s[b * 4 + (lane >> 1)] = make_float4(st.x, a_hi, st.y, d_hi);
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&pfull[b]);
            if (i + 1 < nt) fetch(i + 1);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 128);
            ptx::mbar_wait(&wres_bar, 0);
            const long long t_begin = PROF ? clock64() : 0;
            long long w_slot = 0, w_rows = 0;
            for (int i = 0; i < nt; i++) {
                const int slot = i & 3;
                const long long m0 = PROF ? clock64() : 0;
                if (i >= 4) ptx::mbar_wait(&tempty[slot], (uint32_t)((i >> 2) - 1) & 1);
                if (PROF) w_slot += clock64() - m0;
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(slot * 128);
                ptx::umma_f16(d_tmem, umma_desc_k_sw32(ptx::smem_u32(cb_s)), umma_desc_k_sw32(ptx::smem_u32(cb_s + 4096)), idesc, 0);
                for (int r = 0; r < nrows; r++) {
                    const long long m1 = PROF ? clock64() : 0;
                    ptx::mbar_wait(&rfull[r], (uint32_t)i & 1);
                    if (PROF) w_rows += clock64() - m1;
                    ptx::tc_fence_after();
#pragma unroll
                    for (int kb = 0; kb < 2; kb++) {
                        const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(sA + (size_t)(r * 2 + kb) * FR_UNIT));
                        const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(sW + (size_t)(r * 2 + kb) * FR_UNIT));
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            ptx::umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1);
                    }
                    ptx::umma_commit(&rempty[r]);   // the row slot may be overwritten once these MMAs retire
                }
                ptx::umma_commit(&tfull[slot]);
            }
            if (PROF && a.prof) {
                unsigned long long *o = a.prof + (size_t)blockIdx.x * 16;
                unsigned int smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                atomicAdd(&o[8], (unsigned long long)(clock64() - t_begin));
                o[9] = smid;
                atomicAdd(&o[10], (unsigned long long)w_slot);
                atomicAdd(&o[11], (unsigned long long)w_rows);
            }
        }
    } else if (warp == 2) {
        // ===== statistics warp: lane = 2 s + k (k = 0 sum, 1 sum of squares) of the 8 segments of a tile.  Every word of
        // the exchange table carries its own arrival count in the low byte (each CTA adds (fixed-point value << 8) + 1),
        // so no ordering between different addresses is needed: relaxed atomics, no fences =====
        const double invE = 1.0 / ((double)FR_TO * (double)a.Fo * (double)FR_C);
        const int s = lane >> 1, k = lane & 1;
        bool dead = false;
        for (int j = 0; j < nt; j++) {
            ptx::mbar_wait(&p1done[j & 1], (uint32_t)(j >> 1) & 1);
            unsigned long long *gs = a.gsum + (size_t)j * (FR_S * 2) + lane;
            long long tot = 0;
            if (lane < 2 * FR_S) {
                double d = 0.0;
#pragma unroll
                for (int ww = 0; ww < FR_EPI; ww++) {
                    const float2 pp = part_s[j & 1][ww][s];
                    d += (double)(k ? pp.y : pp.x);
                }
                red_add_u64(gs, ((unsigned long long)__double2ll_rn(d * FR_FIX) << 8) + 1ull);
                unsigned long long wv = 0;
                if (!dead) {
                    int tries = 0;
                    while (((wv = ld_relaxed_u64(gs)) & 0xFFull) != (unsigned long long)gridDim.x) {
                        __nanosleep(40);
                        if (++tries > (1 << 20)) {
                            dead = true;
                            *a.err = 1;
                            break;
                        }
                    }
                }
                tot = (long long)wv >> 8;
            }
            dead = __any_sync(0xffffffffu, dead);
            const double t = (double)tot / FR_FIX;
            const double t2 = __shfl_down_sync(0xffffffffu, t, 1);
            if (lane < 2 * FR_S && k == 0) {
                const double mean = t * invE;
                double var = t2 * invE - mean * mean;
                if (var < 0.0) var = 0.0;
                stat_s[j & 3][s] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&sready[j & 3]);
        }
    } else if (warp >= FR_W_PROD && warp < FR_W_EPI) {
        // ===== producers: time step w, channels 4 lane .. 4 lane + 3, all 8 segments of the tile =====
        const int w = warp - FR_W_PROD;
        const int c0 = 4 * lane, kbp = lane >> 4;
        const uint32_t chunk = (uint32_t)(lane & 15) >> 1, half8 = (uint32_t)(lane & 1) * 8u;
        float wd[4][3], bd[4];                                   // conv1 weights / bias
        uint32_t gg[3][2], gb[3][2];                             // ln1 affine of rows 2 fo + r: bf16 pairs (c0,c0+1), (c0+2,c0+3)
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int j = 0; j < 3; j++) wd[c][j] = __ldg(a.w0 + (c0 + c) * 3 + j);
            bd[c] = __ldg(a.b0 + c0 + c);
        }
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int f = 2 * fo + r < a.F ? 2 * fo + r : a.F - 1;
            const float4 g4 = __ldg(reinterpret_cast<const float4 *>(a.g0 + ((size_t)f * FR_TO + w) * FR_C + c0));
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.be0 + ((size_t)f * FR_TO + w) * FR_C + c0));
            gg[r][0] = pack2(g4.x, g4.y); gg[r][1] = pack2(g4.z, g4.w);
            gb[r][0] = pack2(b4.x, b4.y); gb[r][1] = pack2(b4.z, b4.w);
        }
        const uint32_t sA_u = ptx::smem_u32(sA), P_u = ptx::smem_u32(P_s);
        const uint32_t AS_u = ptx::smem_u32(AS_s);
        long long pc[3] = {0, 0, 0};
        for (int i = 0; i < nt; i++) {
            const int b = i & 1;
            const long long c0t = PROF ? clock64() : 0;
            ptx::mbar_wait(&pfull[b], (uint32_t)(i >> 1) & 1);
            if (PROF) pc[0] += clock64() - c0t;
#pragma unroll
            for (int r = 0; r < 3; r++) {   // unrolled: the per-row register arrays must be indexed statically
                if (r >= nrows) break;
                const long long c1 = PROF ? clock64() : 0;
                if (i > 0) ptx::mbar_wait(&rempty[r], (uint32_t)(i - 1) & 1);
                const long long c2 = PROF ? clock64() : 0;
                if (PROF) pc[1] += c2 - c1;
                // row m = w * 8 + s, 16-byte chunk c of row m lives at chunk c ^ (m % 8) = c ^ s: this thread's 8 bytes
                // are at T ^ (s * 0x90) (bits 4-9 of the unit / row-group base are zero)
                const uint32_t T = sA_u + (uint32_t)(r * 2 + kbp) * FR_UNIT + (uint32_t)w * 1024u + (chunk << 4) + half8;
                const uint32_t Pr = P_u + (uint32_t)((b * 3 + r) * FR_TO + w) * 128u;
#pragma unroll
                for (int sp = 0; sp < 4; sp++) {
                    const float4 p0 = lds128f(Pr + sp * 32);       // m0(s) m0(s+1) m1(s) m1(s+1)   (pre-scaled by rstd_s)
                    const float2 p1 = lds64f(Pr + sp * 32 + 16);   // m2(s) m2(s+1)
                    const float4 as = lds128f(AS_u + (b * 4 + sp) * 16);   // rstd_s, rstd_s+1, -mean rstd (s, s+1)
                    const float2 av = make_float2(as.x, as.y), dv = make_float2(as.z, as.w);
                    const float2 m0 = make_float2(p0.x, p0.y), m1 = make_float2(p0.z, p0.w);
                    float2 acc[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        // x^ = rstd * (w . m + b - mean) for segments (s, s + 1) of channel c0 + c
                        float2 t = ffma2(make_float2(bd[c], bd[c]), av, dv);
                        t = ffma2(make_float2(wd[c][0], wd[c][0]), m0, t);
                        t = ffma2(make_float2(wd[c][1], wd[c][1]), m1, t);
                        acc[c] = ffma2(make_float2(wd[c][2], wd[c][2]), p1, t);
                    }
                    const uint32_t lo0 = affine_relu(pack2(acc[0].x, acc[1].x), gg[r][0], gb[r][0]);
                    const uint32_t lo1 = affine_relu(pack2(acc[2].x, acc[3].x), gg[r][1], gb[r][1]);
                    const uint32_t hi0 = affine_relu(pack2(acc[0].y, acc[1].y), gg[r][0], gb[r][0]);
                    const uint32_t hi1 = affine_relu(pack2(acc[2].y, acc[3].y), gg[r][1], gb[r][1]);
                    sts64(T ^ (uint32_t)(2 * sp * 0x90), lo0, lo1);
                    sts64(T ^ (uint32_t)((2 * sp + 1) * 0x90), hi0, hi1);
                }
                ptx::fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&rfull[r]);
                if (PROF) pc[2] += clock64() - c2;
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&pempty[b]);   // the staged mel of this tile has been consumed
        }
        if (PROF && a.prof && lane == 0) {
            unsigned long long *o = a.prof + (size_t)blockIdx.x * 16;
#pragma unroll
            for (int q = 0; q < 3; q++) atomicAdd(&o[q], (unsigned long long)pc[q]);
            if (w == 0) atomicAdd(&o[7], (unsigned long long)nt);
        }
    } else if (warp >= FR_W_EPI) {
        // ===== epilogue: TMEM lane quarter (rows = 4 time steps x 8 segments), 64 accumulator columns =====
        const int ew = warp - FR_W_EPI;
        const int quarter = warp & 3, cc = (ew >> 2) * 64;
        const int to_e = quarter * 4 + (lane >> 3), s_e = lane & 7;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cc;
        const uint32_t g1_u = ptx::smem_u32(g1_s) + (uint32_t)(to_e * FR_GSTRIDE + cc) * 2u;
        const uint32_t be1_u = ptx::smem_u32(be1_s) + (uint32_t)(to_e * FR_GSTRIDE + cc) * 2u;
        long long pc[4] = {0, 0, 0, 0};
        for (int i = 0; i < nt + FR_DEFER; i++) {
            if (i >= FR_DEFER) {
                // ---------------- pass 2 (tile j): normalise, affine, ReLU, bf16 store ----------------
                const int j = i - FR_DEFER, slot = j & 3;
                const long long c3 = PROF ? clock64() : 0;
                ptx::mbar_wait(&sready[slot], (uint32_t)(j >> 2) & 1);
                const long long c4 = PROF ? clock64() : 0;
                if (PROF) pc[0] += c4 - c3;
                const float2 st = stat_s[slot][s_e];
                const float rstd = st.y, nmr = -st.x * st.y;
                const long long smp = (long long)j * FR_S + s_e;
                unsigned char *dst = reinterpret_cast<unsigned char *>(a.X) +
                                     ((((size_t)smp * a.Fo + fo) * FR_TO + to_e) * FR_C + cc) * 2;
                ptx::tc_fence_after();
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    uint32_t v[16];
                    ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)(slot * 128 + h * 16), v);
                    ptx::tmem_ld_wait();
                    if (h == 3) {   // the TMEM slot is free once its last values are in registers
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&tempty[slot]);
                    }
                    uint32_t pk[8];
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const uint4 g4 = lds128u(g1_u + h * 32 + q * 16), b4 = lds128u(be1_u + h * 32 + q * 16);
                        const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
                        const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int c = 8 * q + 2 * e;
                            const float2 x = ffma2(make_float2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])),
                                                   make_float2(rstd, rstd), make_float2(nmr, nmr));
                            pk[4 * q + e] = affine_relu(pack2(x.x, x.y), gw[e], bw[e]);
                        }
                    }
                    if (smp < a.nb) ptx::st_global_v8(dst + h * 32, pk);
                }
                if (PROF) pc[1] += clock64() - c4;
            }
            if (i < nt) {
                // ---------------- pass 1 (tile i): per-segment sums of this warp's 32 rows x 64 columns ----------------
                const int slot = i & 3;
                const long long c5 = PROF ? clock64() : 0;
                ptx::mbar_wait(&tfull[slot], (uint32_t)(i >> 2) & 1);
                const long long c6 = PROF ? clock64() : 0;
                if (PROF) pc[2] += c6 - c5;
                ptx::tc_fence_after();
                float2 p1[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, p2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(t_lane + (uint32_t)(slot * 128 + h * 32), v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; q++) {   // packed fp32 pairs: half the instructions of the scalar form
                        const float2 o0 = make_float2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]));
                        const float2 o1 = make_float2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                        p1[0] = fadd2(p1[0], o0); p1[1] = fadd2(p1[1], o1);
                        p2[0] = ffma2(o0, o0, p2[0]); p2[1] = ffma2(o1, o1, p2[1]);
                    }
                }
                float s1 = (p1[0].x + p1[0].y) + (p1[1].x + p1[1].y);
                float s2 = (p2[0].x + p2[0].y) + (p2[1].x + p2[1].y);
                // lanes with equal lane % 8 hold the same segment (4 time steps per warp)
                s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
                s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
                s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
                if (lane < FR_S) part_s[i & 1][ew][lane] = make_float2(s1, s2);   // at most two tiles are in flight here
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&p1done[i & 1]);
                if (PROF) pc[3] += clock64() - c6;
            }
        }
        if (PROF && a.prof && lane == 0) {
            unsigned long long *o = a.prof + (size_t)blockIdx.x * 16;
#pragma unroll
            for (int q = 0; q < 4; q++) atomicAdd(&o[3 + q], (unsigned long long)pc[q]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

constexpr size_t FR_SMEM = 1024 + 12 * (size_t)FR_UNIT + 8192 + 2 * FR_TO * FR_GSTRIDE * 2 + 2 * (3 * FR_TO * 4 * 8 * 4 + 4 * 16);

}  // namespace

namespace pfann {

// The fused kernel covers conv 0 + conv 1 of configs like default.json: one input channel, 128 channels out of both
// convolutions, 16 time steps after conv1, taps {0, 1, 2} on both axes, at most one CTA per SM.
bool tc_front_supported(Model *m) {
    // read per call: tests toggle these to compare against the round-1 kernels (l0_tc_kernel + conv_ln_tc_kernel),
    // the unfused LayerNorm chain and the CUDA-core layer 0
    if (getenv("PFANN_B200_NO_FRONT") != nullptr || getenv("PFANN_B200_NO_FUSED_LN") != nullptr ||
        getenv("PFANN_B200_NO_L0_TC") != nullptr)
        return false;
    if (m->tc_state == nullptr || !m->l0_fused) return false;
    const ConvGeom &g0 = m->conv[0].g, &g1 = m->conv[1].g;
    if (g0.Ci != 1 || g0.Co != FR_C || g1.Ci != FR_C || g1.Co != FR_C || g1.depthwise) return false;
    if (g0.To != FR_TO || g0.Ti != 2 * FR_TO || g1.To != FR_TO || g0.axis != 0 || g1.axis != 1) return false;
    if (g0.ntaps != 3 || g1.ntaps != 3) return false;
    for (int j = 0; j < 3; j++)
        if (g0.tap_off[j] != j || g1.tap_off[j] != j) return false;
    if (g1.Fo > m->ctx->sm_count || g1.Fo > 255 || g1.Fi != g0.Fo || g1.Fo * 2 != g1.Fi) return false;  // arrivals: 1 byte
    return tc_weight_map128(m, 1) != nullptr;
}

int tc_front(Model *m, const float *mel, const float2 *stats0, __nv_bfloat16 *Xout, int nb) {
    const ConvGeom &g0 = m->conv[0].g, &g1 = m->conv[1].g;
    FrontArgs a = {};
    a.mel = mel; a.stats0 = stats0; a.w0 = m->l0_w; a.b0 = m->conv[0].bias;
    a.g0 = m->conv[0].gamma; a.be0 = m->conv[0].beta;
    a.b1 = m->conv[1].bias; a.g1 = m->conv[1].gamma; a.be1 = m->conv[1].beta;
    a.X = Xout; a.nb = nb; a.n_groups = (nb + FR_S - 1) / FR_S; a.F = g0.Fi; a.T = g0.Ti; a.Fo = g1.Fo;
    // exchange table: fixed-point sums + arrival counters, zeroed per launch
    const size_t tab_bytes = (size_t)a.n_groups * FR_S * 2 * sizeof(unsigned long long);
    PF_TRY(m->ln_part.ensure(tab_bytes));
    PF_TRY(tc_ln_err_ptr(m, &a.err));
    PF_CUDA(cudaMemsetAsync(m->ln_part.p, 0, tab_bytes, m->ctx->stream));
    a.gsum = m->ln_part.as<unsigned long long>();
    const char *pp = getenv("PFANN_FRONT_PROF_PTR");   // probe only: zeroed device buffer [grid][8] u64
    a.prof = pp ? reinterpret_cast<unsigned long long *>(strtoull(pp, nullptr, 0)) : nullptr;
    if (a.prof)
        PF_CUDA(cudaFuncSetAttribute(front_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FR_SMEM));
    else
        PF_CUDA(cudaFuncSetAttribute(front_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FR_SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.Fo);
    cfg.blockDim = dim3(FR_THREADS);
    cfg.dynamicSmemBytes = FR_SMEM;
    cfg.stream = m->ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;   // the CTAs wait for each other's statistics: all must be resident
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ProfScope ps(m->ctx, K_CONV_TC, 1);
    if (a.prof)
        PF_CUDA(cudaLaunchKernelEx(&cfg, front_tc_kernel<true>, *tc_weight_map128(m, 1), a));
    else
        PF_CUDA(cudaLaunchKernelEx(&cfg, front_tc_kernel<false>, *tc_weight_map128(m, 1), a));
    m->ctx->launches++;
    return PFANN_OK;
}

}  // namespace pfann
