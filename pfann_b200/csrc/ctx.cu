// ctx.cu -- context, error string, host<->device staging for the C-ABI (include/pfann_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "pfann_b200.h"

namespace pfann {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return PFANN_OK;
    if (p) {
        PF_CUDA(cudaFree(p));  // implicit device sync: nothing in flight still uses it
        p = nullptr;
        cap = 0;
    }
    size_t want = bytes + (bytes >> 3) + 256;  // small slack so slowly growing batches do not thrash
    PF_CUDA(cudaMalloc(&p, want));
    cap = want;
    return PFANN_OK;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int stage_input(Ctx *ctx, int slot, const void *src, size_t bytes, const void **dev) {
    if (bytes == 0 || is_device_ptr(src)) {
        *dev = src;
        return PFANN_OK;
    }
    PF_TRY(ctx->stage_in[slot].ensure(bytes));
    PF_CUDA(cudaMemcpyAsync(ctx->stage_in[slot].p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *dev = ctx->stage_in[slot].p;
    return PFANN_OK;
}

int stage_output(Ctx *ctx, int slot, void *dst, size_t bytes, void **dev) {
    if (bytes == 0 || is_device_ptr(dst)) {
        *dev = dst;
        return PFANN_OK;
    }
    PF_TRY(ctx->stage_out[slot].ensure(bytes));
    *dev = ctx->stage_out[slot].p;
    return PFANN_OK;
}

int finish_output(Ctx *ctx, int slot, void *dst, size_t bytes) {
    if (bytes == 0 || dst == nullptr) return PFANN_OK;
    if (ctx->stage_out[slot].p && dst != ctx->stage_out[slot].p && !is_device_ptr(dst)) {
        PF_CUDA(cudaMemcpyAsync(dst, ctx->stage_out[slot].p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        PF_CUDA(cudaStreamSynchronize(ctx->stream));  // host results must be complete on return
    }
    return PFANN_OK;
}

static cudaEvent_t prof_get_event(Ctx *c) {
    if (!c->prof_pool.empty()) {
        cudaEvent_t e = c->prof_pool.back();
        c->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(Ctx *ctx, int klass, int detail) : c(ctx), k(klass), sub(detail) {
    if (!c->profile) return;
    e0 = prof_get_event(c);
    e1 = prof_get_event(c);
    cudaEventRecord(e0, c->stream);
}
ProfScope::~ProfScope() {
    if (!e0) return;
    cudaEventRecord(e1, c->stream);
    if (sub >= 0) c->prof_detail.push_back({sub, {k, (int)c->prof_events[k].size()}});
    c->prof_events[k].push_back({e0, e1});
}

}  // namespace pfann

using namespace pfann;

extern "C" {

long long pfann_version(void) { return PFANN_B200_VERSION; }

const char *pfann_last_error(void) { return get_error(); }

int pfann_ctx_create(int device, pfann_ctx **out) {
    PF_CHECK(out != nullptr, PFANN_ERR_ARG, "pfann_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("pfann_ctx_create: no CUDA device available (%s); libpfann_b200 has no CPU fallback",
                  cudaGetErrorString(e));
        cudaGetLastError();
        return PFANN_ERR_CUDA;
    }
    PF_CHECK(device >= 0 && device < n, PFANN_ERR_ARG, "pfann_ctx_create: device %d out of range (%d devices)",
             device, n);
    PF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PF_CUDA(cudaGetDeviceProperties(&prop, device));
    PF_CHECK(prop.major == 10, PFANN_ERR_UNSUPPORTED,
             "pfann_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
             prop.major, prop.minor);
    Ctx *c = new Ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->stream = nullptr;  // legacy default stream until pfann_ctx_set_stream
    *out = reinterpret_cast<pfann_ctx *>(c);
    return PFANN_OK;
}

void pfann_ctx_destroy(pfann_ctx *h) {
    Ctx *c = reinterpret_cast<Ctx *>(h);
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    for (int i = 0; i < 4; i++) {
        c->stage_in[i].release();
        c->stage_out[i].release();
    }
    delete c;
}

int pfann_ctx_set_stream(pfann_ctx *h, void *stream) {
    PF_CHECK(h != nullptr, PFANN_ERR_ARG, "pfann_ctx_set_stream: ctx is NULL");
    reinterpret_cast<Ctx *>(h)->stream = reinterpret_cast<cudaStream_t>(stream);
    return PFANN_OK;
}

int pfann_ctx_sync(pfann_ctx *h) {
    PF_CHECK(h != nullptr, PFANN_ERR_ARG, "pfann_ctx_sync: ctx is NULL");
    Ctx *c = reinterpret_cast<Ctx *>(h);
    PF_CUDA(cudaSetDevice(c->device));
    PF_CUDA(cudaStreamSynchronize(c->stream));
    return PFANN_OK;
}

int pfann_ctx_profile(pfann_ctx *h, int enable) {
    PF_CHECK(h != nullptr, PFANN_ERR_ARG, "pfann_ctx_profile: ctx is NULL");
    reinterpret_cast<Ctx *>(h)->profile = enable != 0;
    return PFANN_OK;
}

int pfann_ctx_profile_read(pfann_ctx *h, double *ms, long long *count, int n_classes) {
    PF_CHECK(h && ms && count, PFANN_ERR_ARG, "pfann_ctx_profile_read: NULL argument");
    Ctx *c = reinterpret_cast<Ctx *>(h);
    PF_CUDA(cudaSetDevice(c->device));
    PF_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < n_classes; k++) {
        ms[k] = 0.0;
        count[k] = 0;
    }
    for (int i = 0; i < PFANN_N_PROFILE_DETAIL; i++) c->detail_ms[i] = 0.0, c->detail_count[i] = 0;
    for (auto &d : c->prof_detail) {
        auto &pr = c->prof_events[d.second.first][d.second.second];
        float t = 0.f;
        cudaEventElapsedTime(&t, pr.first, pr.second);
        if (d.first < PFANN_N_PROFILE_DETAIL) c->detail_ms[d.first] += t, c->detail_count[d.first]++;
    }
    c->prof_detail.clear();
    for (int k = 0; k < K_NCLASS; k++) {
        for (auto &pr : c->prof_events[k]) {
            float t = 0.f;
            cudaEventElapsedTime(&t, pr.first, pr.second);
            if (k < n_classes) {
                ms[k] += t;
                count[k]++;
            }
            c->prof_pool.push_back(pr.first);
            c->prof_pool.push_back(pr.second);
        }
        c->prof_events[k].clear();
    }
    return PFANN_OK;
}

int pfann_ctx_profile_detail(pfann_ctx *h, double *ms, long long *count, int n) {
    PF_CHECK(h && ms && count, PFANN_ERR_ARG, "pfann_ctx_profile_detail: NULL argument");
    Ctx *c = reinterpret_cast<Ctx *>(h);
    for (int i = 0; i < n; i++) {
        ms[i] = i < PFANN_N_PROFILE_DETAIL ? c->detail_ms[i] : 0.0;
        count[i] = i < PFANN_N_PROFILE_DETAIL ? c->detail_count[i] : 0;
    }
    return PFANN_OK;
}

long long pfann_ctx_launches(pfann_ctx *h) { return h ? reinterpret_cast<Ctx *>(h)->launches : 0; }

int pfann_ctx_sm_count(pfann_ctx *h) { return h ? reinterpret_cast<Ctx *>(h)->sm_count : 0; }

}  // extern "C"
