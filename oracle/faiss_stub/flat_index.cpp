/* Flat fp32 index behind the faiss::Index stub, so that the reference's seq_score
 * (compiled unmodified from /root/reference/cpp/seqscore.cpp) has something to
 * reconstruct() from.  TEST INFRASTRUCTURE ONLY. */
#include <cstring>
#include <faiss/Index.h>

namespace {
struct FlatIndex : faiss::Index {
    const float *data;
    void reconstruct(faiss::idx_t key, float *recons) const override {
        std::memcpy(recons, data + (size_t)key * d, sizeof(float) * d);
    }
};
}  // namespace

/* `data` is borrowed: the caller keeps it alive for the lifetime of the handle. The returned
 * pointer is the faiss::Index* that seq_score() casts its first argument to. */
extern "C" void *ref_flat_index_new(const float *data, int64_t n, int d) {
    FlatIndex *ix = new FlatIndex();
    ix->d = d;
    ix->ntotal = n;
    ix->data = data;
    return static_cast<faiss::Index *>(ix);
}

extern "C" void ref_flat_index_free(void *p) { delete static_cast<faiss::Index *>(p); }
