// common.cuh -- shared host/device helpers for libpfann_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------
// error handling: nothing throws across the C-ABI; every entry point returns <0 on error and
// stores a message retrievable with pfann_last_error().
// ------------------------------------------------------------------------------------------
namespace pfann {

void set_error(const char *fmt, ...);
const char *get_error();

enum {
    PFANN_OK = 0,
    PFANN_ERR_CUDA = -1,
    PFANN_ERR_ARG = -2,
    PFANN_ERR_UNSUPPORTED = -3,
    PFANN_ERR_STATE = -4,
    PFANN_ERR_NOMEM = -5,
};

#define PF_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::pfann::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                               cudaGetErrorString(_e));                                       \
            cudaGetLastError(); /* do not leave a stale error for the next check */           \
            return ::pfann::PFANN_ERR_CUDA;                                                   \
        }                                                                                     \
    } while (0)

#define PF_CHECK(cond, code, ...)                                                             \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ::pfann::set_error(__VA_ARGS__);                                                  \
            return (code);                                                                    \
        }                                                                                     \
    } while (0)

#define PF_TRY(expr)                                                                          \
    do {                                                                                      \
        int _rc = (expr);                                                                     \
        if (_rc < 0) return _rc;                                                              \
    } while (0)

// Grow-only device buffer owned by a context.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// kernel classes for the optional per-class CUDA-event timing (bench.py roofline)
enum KernelClass { K_MEL = 0, K_CONV_TC, K_CONV_CC, K_LN, K_HEAD, K_KNN_SCAN, K_KNN_SELECT, K_RERANK, K_MISC, K_NCLASS };

// Per-device context: device ordinal, the stream all work is enqueued on, scratch, launch counter.
struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // H2D / D2H pipelines of pfann_extract_pcm16 (host buffers)
    // extraction overlap: the encoder of chunk k on a high-priority stream, the mel kernel of chunk k + 1 on a
    // low-priority one (its CTAs fill the SMs the cooperative encoder kernels leave idle)
    cudaStream_t enc_stream = nullptr, mel_stream = nullptr;
    int sm_count = 148;
    long long launches = 0;  // kernels of OURS launched through this context (bench: gpu_launches)
    DevBuf stage_in[4], stage_out[4];
    bool profile = false;    // record a CUDA-event pair around every kernel launch, per class
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[K_NCLASS];
    // finer breakdown of the same event pairs: (detail slot, index into prof_events[class]); slots are
    // 0-15 conv idx, 16-31 LayerNorm kernels after conv idx, 32 mel, 33 head, 34 layer-0 moments
    std::vector<std::pair<int, std::pair<int, int>>> prof_detail;
    std::vector<cudaEvent_t> prof_pool;
    double detail_ms[48] = {};       // filled by pfann_ctx_profile_read, returned by pfann_ctx_profile_detail
    long long detail_count[48] = {};
};

// RAII: events on the launching stream around one kernel launch (no-op unless ctx->profile)
struct ProfScope {
    Ctx *c;
    int k, sub;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ProfScope(Ctx *ctx, int klass, int detail = -1);
    ~ProfScope();
};

// Where does a caller pointer live?  Host buffers are staged through ctx scratch.
bool is_device_ptr(const void *p);

// RAII-free staging helpers (return device pointer usable on ctx->stream).
int stage_input(Ctx *ctx, int slot, const void *src, size_t bytes, const void **dev);
// For outputs: returns a device pointer to write to; call finish_output afterwards.
int stage_output(Ctx *ctx, int slot, void *dst, size_t bytes, void **dev);
int finish_output(Ctx *ctx, int slot, void *dst, size_t bytes);  // D2H + sync when dst is host

static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

}  // namespace pfann

// ------------------------------------------------------------------------------------------
// device-side PTX helpers (mbarrier, bulk/TMA copies, tcgen05)
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace pfann {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- thread-block-cluster helpers (distributed shared memory) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of the same shared-memory variable in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f64x2(uint32_t addr, double a, double b) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
// arrive (release at cluster scope) on an mbarrier that may live in another CTA of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); 16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Tiled TMA loads (cp.async.bulk.tensor, SASS UTMALDG) with an mbarrier completion.
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const void *tmap, uint64_t *bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const void *tmap, uint64_t *bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6, %7}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16/fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit all prior tcgen05.mma of this thread; arrives (count 1) on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32-bit, 32 consecutive columns: thread i of the warp gets TMEM lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 32-byte global store (sm_100: STG.256): one full sector per thread
__device__ __forceinline__ void st_global_v8(void *p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 bytes with the
// 128-byte swizzle (what TMA SWIZZLE_128B writes): SBO = 1024 B (8 rows), LBO unused, version 1.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                       // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // layout: SWIZZLE_128B
    return d;
}
// Instruction descriptor: kind::f16, A/B = bf16 K-major, D = fp32, shape M x N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4)                      // c_format = F32
           | (1u << 7)                    // a_format = BF16
           | (1u << 10)                   // b_format = BF16
           | ((uint32_t)(N >> 3) << 17)   // n_dim
           | ((uint32_t)(M >> 4) << 24);  // m_dim
}

}  // namespace ptx
}  // namespace pfann
#endif  // __CUDACC__
