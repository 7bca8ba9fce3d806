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

    // empty slots carry `seed` (id -1): only candidates beating the seed are ever inserted
    __device__ __forceinline__ void init(int k_) { init(k_, -CUDART_INF_F); }
    __device__ __forceinline__ void init(int k_, float seed) {
        k = k_;
#pragma unroll 1
        for (int i = 0; i < k_; ++i) {
            v[i] = seed;
            id[i] = -1;
        }
        thr = seed;
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


// Register-resident candidate set for k <= 32: UNSORTED values in 32 registers, ids in thread-local
// memory (only ever stored while scanning).  An insert overwrites the current minimum and recomputes
// the minimum with a max/min tree -- no dependent load/store chain, unlike sorted insertion, which is
// what makes the selection epilogue keep up with a single-product (coarse) main loop.  finalize()
// sorts once per work item into the canonical (score desc, id asc) order.
struct RegList32 {
    float v[32];
    int id[32];
    int k;
    float thr;

    __device__ __forceinline__ void init(int k_) { init(k_, -CUDART_INF_F); }
    __device__ __forceinline__ void init(int k_, float seed) {
        k = k_;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            v[i] = (i < k_) ? seed : CUDART_INF_F;   // slots >= k never become the minimum
            id[i] = -1;
        }
        thr = seed;
    }
    // precondition: val > thr.  Evicts the current minimum; among several slots holding the same
    // minimum the one with the LARGEST id goes, so exact ties keep their lowest ids (Faiss order).
    __device__ __forceinline__ void insert(float val, int idx) {
        int pos = 0, cnt = 0;
#pragma unroll
        for (int i = 31; i >= 0; --i) {
            const bool eq = (v[i] == thr);
            pos = eq ? i : pos;
            cnt += eq ? 1 : 0;
        }
        if (cnt > 1) {  // rare: tie at the boundary
            int worst_id = -2;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int ii = id[i];
                const bool take = (v[i] == thr) && (ii < 0 || (worst_id != -1 && ii > worst_id));
                // an empty slot (id -1) is always the preferred victim
                if (take) { pos = i; worst_id = ii; }
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (i == pos) ? val : v[i];
        id[pos] = idx;
        float m = v[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) m = fminf(m, v[i]);
        thr = m;
    }
    // selection sort of the k live entries into (score desc, id asc); returns through out arrays
    template <typename F>
    __device__ __forceinline__ void drain_sorted(F&& emit) {
        for (int j = 0; j < k; ++j) {
            float bv = -CUDART_INF_F;
            int bi = 0x7fffffff, bp = -1;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int ii = id[i];
                const bool live = (i < k) && (ii >= 0);
                const bool better = live && (v[i] > bv || (v[i] == bv && ii < bi));
                bv = better ? v[i] : bv;
                bi = better ? ii : bi;
                bp = better ? i : bp;
            }
            if (bp < 0) { emit(j, 0.f, -1); continue; }
            emit(j, bv, bi);
            id[bp] = -1;
        }
    }
};
