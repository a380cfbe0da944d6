/* Minimal stand-in for <faiss/Index.h>, written for this repo (NOT a copy of faiss).
 *
 * The reference's cpp/seqscore.cpp includes <faiss/Index.h> but touches exactly two members
 * of faiss::Index: the dimension `d` (seqscore.cpp:46) and the virtual
 * `reconstruct(key, float*)` (seqscore.cpp:96).  faiss is not installed in this image and
 * cannot be fetched, so oracle/Makefile compiles the UNMODIFIED reference source against
 * this stub plus flat_index.cpp (which provides the only concrete Index we need: a flat
 * fp32 matrix).  The ABI is self-consistent because both sides are compiled together.
 * TEST INFRASTRUCTURE ONLY (see oracle/pfann_oracle.c header).
 */
#pragma once
#include <cstdint>

namespace faiss {

typedef int64_t idx_t;

struct Index {
    int d;
    idx_t ntotal;
    virtual ~Index() {}
    virtual void reconstruct(idx_t key, float *recons) const = 0;
};

}  // namespace faiss
