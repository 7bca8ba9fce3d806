// Per-thread sorted candidate list used by every selection kernel.
#pragma once

#include <math_constants.h>
#include <stdint.h>

// Keeps the k best (largest raw score) candidates seen so far, sorted best-first.  Candidates must
// be offered in ascending id order: insertion is strict (`>` the current k-th score) and goes after
// equal scores, which yields the canonical (score desc, id asc) order Faiss's handlers produce
// (distances.cpp: Top1BlockResultHandler / HeapBlockResultHandler, strict comparisons).
template <int KMAX>
struct TopKList {
    float v[KMAX];
    int id[KMAX];
    int k;
    float thr;  // score of the current k-th entry (-inf while fewer than k real entries)

    __device__ __forceinline__ void init(int k_) {
        k = k_;
#pragma unroll 1
        for (int i = 0; i < k_; ++i) {
            v[i] = -CUDART_INF_F;
            id[i] = -1;
        }
        thr = -CUDART_INF_F;
    }
    // precondition: val > thr
    __device__ __noinline__ void insert(float val, int idx) {
        int p = k - 1;
        while (p > 0 && v[p - 1] < val) {
            v[p] = v[p - 1];
            id[p] = id[p - 1];
            --p;
        }
        v[p] = val;
        id[p] = idx;
        thr = v[k - 1];
    }
};
