// Exact FP32 flat search on CUDA cores for small query batches.
//
// Faiss switches algorithm on the query count (distances.cpp: distance_compute_blas_threshold = 20):
// below it, every (query, database) pair is a direct FP32 dot product / sum of squared differences
// with no |x|^2 + |y|^2 - 2<x,y> expansion.  The reference's online query is exactly that case
// (engine.py:55, nq = 1, k = 20).  Two phases: (1) one warp per (query, db row) pair streams the
// row with 128-bit loads; (2) per-query slices are reduced to sorted top-k lists and merged.
#include "common.cuh"
#include "topk_list.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int64_t kMinSliceLen = 16384;  // scores per phase-2 block (grows so that slices <= 256)

template <bool L2>
__global__ void pair_scores_kernel(const float* __restrict__ q, int64_t nq, const float* __restrict__ db,
                                   int64_t nb, int d, float* __restrict__ scores) {
    extern __shared__ float qs[];  // this block's query row
    const int64_t qi = blockIdx.y;
    for (int c = threadIdx.x; c < d; c += blockDim.x) qs[c] = q[qi * d + c];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(db) & 15) == 0);
    for (int64_t r = warp; r < nb; r += nwarps) {
        const float* row = db + r * (int64_t)d;
        float acc = 0.f;
        if (vec) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            const float4* q4 = reinterpret_cast<const float4*>(qs);
            for (int c = lane; c < d / 4; c += 32) {
                const float4 y = __ldg(row4 + c);
                const float4 x = q4[c];
                if (L2) {
                    const float a = x.x - y.x, b = x.y - y.y, e = x.z - y.z, f = x.w - y.w;
                    acc = fmaf(a, a, acc); acc = fmaf(b, b, acc); acc = fmaf(e, e, acc); acc = fmaf(f, f, acc);
                } else {
                    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
                    acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
                }
            }
        } else {
            for (int c = lane; c < d; c += 32) {
                const float y = __ldg(row + c), x = qs[c];
                if (L2) { const float a = x - y; acc = fmaf(a, a, acc); }
                else acc = fmaf(x, y, acc);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) scores[qi * nb + r] = acc;
    }
}

// block (slice, query): top-k of scores[query, slice*kSliceLen ...) -> parts[slice, query, :]
template <bool LARGEST, int KMAX>
__global__ void slice_topk_kernel(const float* __restrict__ scores, int64_t nq, int64_t nb, int topk,
                                  int64_t slice_len, int64_t id_base, float* __restrict__ pv, int64_t* __restrict__ pi) {
    __shared__ float s_v[kWarps];
    __shared__ int s_i[kWarps];
    __shared__ int s_t[kWarps];
    __shared__ int s_win;
    const int64_t qi = blockIdx.y;
    const int64_t lo = (int64_t)blockIdx.x * slice_len;
    const int64_t hi = min(nb, lo + slice_len);
    const float* row = scores + qi * nb;
    TopKList<KMAX> list;
    list.init(topk);
    // ascending ids per thread => strict insert keeps the lower id among equal scores
    for (int64_t j = lo + threadIdx.x; j < hi; j += kThreads) {
        const float s = LARGEST ? row[j] : -row[j];
        if (s > list.thr) list.insert(s, (int)(j - lo));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int head = 0;
    float* ov = pv + ((int64_t)blockIdx.x * nq + qi) * topk;
    int64_t* oi = pi + ((int64_t)blockIdx.x * nq + qi) * topk;
    for (int j = 0; j < topk; ++j) {
        float v = (head < topk && list.id[head] >= 0) ? list.v[head] : -CUDART_INF_F;
        int id = (head < topk && list.id[head] >= 0) ? list.id[head] : 0x7fffffff;
        int t = threadIdx.x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int id2 = __shfl_xor_sync(0xffffffffu, id, o);
            const int t2 = __shfl_xor_sync(0xffffffffu, t, o);
            if (v2 > v || (v2 == v && id2 < id)) { v = v2; id = id2; t = t2; }
        }
        if (lane == 0) { s_v[warp] = v; s_i[warp] = id; s_t[warp] = t; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float bv = s_v[0]; int bi = s_i[0], bt = s_t[0];
            for (int w2 = 1; w2 < kWarps; ++w2)
                if (s_v[w2] > bv || (s_v[w2] == bv && s_i[w2] < bi)) { bv = s_v[w2]; bi = s_i[w2]; bt = s_t[w2]; }
            if (bi != 0x7fffffff) {
                ov[j] = LARGEST ? bv : -bv;
                oi[j] = id_base + lo + bi;
                s_win = bt;
            } else {
                ov[j] = LARGEST ? -3.402823466e+38f : 3.402823466e+38f;
                oi[j] = -1;
                s_win = -1;
            }
        }
        __syncthreads();
        if (s_win == (int)threadIdx.x) head++;
        __syncthreads();
    }
}

// k > 128: the selected entries of a pass are masked (+-inf is never selected) so the next pass returns the next
// best 128 in the same canonical (score, id) order
__global__ void scores_mask_kernel(float* __restrict__ scores, int64_t nq, int64_t nb, const int64_t* __restrict__ idx,
                                   int kk, int64_t id_base, float fill) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * kk) return;
    const int64_t id = idx[t];
    if (id >= 0) scores[(t / kk) * nb + (id - id_base)] = fill;
}

}  // namespace

static int64_t exact_slice_len(int64_t nb) {
    return std::max<int64_t>(kMinSliceLen, ceil_div64(ceil_div64(nb, 256), 256) * 256);
}
static int64_t exact_num_slices(int64_t nb) { return std::max<int64_t>(1, ceil_div64(nb, exact_slice_len(nb))); }

// per-query top-k of a [nq, nb] score matrix: sorted slice lists, then the canonical (score, id) merge.  Scores of
// +inf (L2) / -inf (IP) are never selected, so short rows are padded with id -1.
static int select_from_scores(ise_ctx* ctx, const float* scores, int64_t nq, int64_t nb, int metric, int topk,
                              int64_t id_base, float* out_val, int64_t* out_idx, int64_t* pi, float* pv,
                              cudaStream_t st) {
    const int64_t slices = exact_num_slices(nb);
    const int64_t slen = exact_slice_len(nb);
    const bool single = slices == 1;
    float* tv = single ? out_val : pv;
    int64_t* ti = single ? out_idx : pi;
    dim3 g2((unsigned)slices, (unsigned)nq);
    const bool largest = metric == ISE_METRIC_IP;
    if (topk <= 32) {
        if (largest) slice_topk_kernel<true, 32><<<g2, kThreads, 0, st>>>(scores, nq, nb, topk, slen, id_base, tv, ti);
        else slice_topk_kernel<false, 32><<<g2, kThreads, 0, st>>>(scores, nq, nb, topk, slen, id_base, tv, ti);
    } else {
        if (largest) slice_topk_kernel<true, 128><<<g2, kThreads, 0, st>>>(scores, nq, nb, topk, slen, id_base, tv, ti);
        else slice_topk_kernel<false, 128><<<g2, kThreads, 0, st>>>(scores, nq, nb, topk, slen, id_base, tv, ti);
    }
    ISE_LAUNCH_CHECK();
    if (single) return 0;
    return ise_topk_merge(ctx, pv, pi, (int)slices, nq, topk, metric, out_val, out_idx, (void*)st);
}

ISE_EXPORT size_t ise_flat_search_exact_workspace_bytes(ise_ctx* ctx, int64_t nq, int64_t nb, int topk) {
    if (!ctx || nq <= 0 || nb <= 0 || topk <= 0) return 0;
    const size_t scores = (size_t)nq * (size_t)nb * sizeof(float);
    const size_t parts = (size_t)exact_num_slices(nb) * (size_t)nq * (size_t)topk * (sizeof(float) + sizeof(int64_t));
    return ((scores + 255) & ~size_t(255)) + parts + 256;
}

ISE_EXPORT int ise_flat_search_exact(ise_ctx* ctx, const float* q, int64_t nq, const float* db, int64_t nb, int d,
                                     int metric, int topk, int64_t id_base, float* out_val, int64_t* out_idx,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    ISE_CHECK_ARG(ctx != nullptr);
    ISE_CHECK_ARG(metric == ISE_METRIC_IP || metric == ISE_METRIC_L2);
    ISE_CHECK_ARG(nq >= 0 && nb > 0 && d > 0 && topk >= 1 && topk <= 128 && nq <= 65535);
    ISE_CHECK_ARG((size_t)d * sizeof(float) <= 48 * 1024);
    if (nq == 0) return 0;
    ISE_CHECK_ARG(q && db && out_val && out_idx && workspace);
    const size_t need = ise_flat_search_exact_workspace_bytes(ctx, nq, nb, topk);
    if (workspace_bytes < need) ISE_FAIL("workspace too small: need " + std::to_string(need));
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t slices = exact_num_slices(nb);
    float* scores = reinterpret_cast<float*>(workspace);
    const size_t scores_bytes = (((size_t)nq * nb * sizeof(float)) + 255) & ~size_t(255);
    int64_t* pi = reinterpret_cast<int64_t*>(reinterpret_cast<uint8_t*>(workspace) + scores_bytes);
    float* pv = reinterpret_cast<float*>(pi + (size_t)slices * nq * topk);

    dim3 g1((unsigned)std::min<int64_t>(ceil_div64(nb, kWarps), (int64_t)ctx->sm_count * 8), (unsigned)nq);
    if (metric == ISE_METRIC_L2)
        pair_scores_kernel<true><<<g1, kThreads, d * sizeof(float), st>>>(q, nq, db, nb, d, scores);
    else
        pair_scores_kernel<false><<<g1, kThreads, d * sizeof(float), st>>>(q, nq, db, nb, d, scores);
    ISE_LAUNCH_CHECK();

    return select_from_scores(ctx, scores, nq, nb, metric, topk, id_base, out_val, out_idx, pi, pv, st);
}

ISE_EXPORT size_t ise_scores_topk_workspace_bytes(ise_ctx* ctx, int64_t nq, int64_t nb, int topk) {
    if (!ctx || nq <= 0 || nb <= 0 || topk <= 0) return 0;
    return (size_t)exact_num_slices(nb) * (size_t)nq * (size_t)topk * (sizeof(float) + sizeof(int64_t)) + 256;
}

ISE_EXPORT int ise_scores_topk(ise_ctx* ctx, const float* scores, int64_t nq, int64_t nb, int metric, int topk,
                               int64_t id_base, float* out_val, int64_t* out_idx, void* workspace,
                               size_t workspace_bytes, void* stream) {
    ISE_CHECK_ARG(ctx != nullptr);
    ISE_CHECK_ARG(metric == ISE_METRIC_IP || metric == ISE_METRIC_L2);
    ISE_CHECK_ARG(nq >= 0 && nb > 0 && topk >= 1 && topk <= 128 && nq <= 65535);
    if (nq == 0) return 0;
    ISE_CHECK_ARG(scores && out_val && out_idx && workspace);
    if (workspace_bytes < ise_scores_topk_workspace_bytes(ctx, nq, nb, topk)) ISE_FAIL("workspace too small");
    DeviceGuard guard(ctx->device);
    const int64_t slices = exact_num_slices(nb);
    int64_t* pi = reinterpret_cast<int64_t*>(workspace);
    float* pv = reinterpret_cast<float*>(pi + (size_t)slices * nq * topk);
    return select_from_scores(ctx, scores, nq, nb, metric, topk, id_base, out_val, out_idx, pi, pv, (cudaStream_t)stream);
}

ISE_EXPORT int ise_pair_scores(ise_ctx* ctx, const float* q, int64_t nq, const float* db, int64_t nb, int d, int metric,
                               float* scores, void* stream) {
    ISE_CHECK_ARG(ctx != nullptr);
    ISE_CHECK_ARG(metric == ISE_METRIC_IP || metric == ISE_METRIC_L2);
    ISE_CHECK_ARG(nq >= 0 && nb > 0 && d > 0 && nq <= 65535 && (size_t)d * sizeof(float) <= 48 * 1024);
    if (nq == 0) return 0;
    ISE_CHECK_ARG(q && db && scores);
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 g1((unsigned)std::min<int64_t>(ceil_div64(nb, kWarps), (int64_t)ctx->sm_count * 8), (unsigned)nq);
    if (metric == ISE_METRIC_L2)
        pair_scores_kernel<true><<<g1, kThreads, d * sizeof(float), st>>>(q, nq, db, nb, d, scores);
    else
        pair_scores_kernel<false><<<g1, kThreads, d * sizeof(float), st>>>(q, nq, db, nb, d, scores);
    ISE_LAUNCH_CHECK();
    return 0;
}

ISE_EXPORT int ise_scores_mask(ise_ctx* ctx, float* scores, int64_t nq, int64_t nb, const int64_t* idx, int kk,
                               int64_t id_base, int metric, void* stream) {
    ISE_CHECK_ARG(ctx && nq >= 0 && nb > 0 && kk >= 1);
    ISE_CHECK_ARG(metric == ISE_METRIC_IP || metric == ISE_METRIC_L2);
    if (nq == 0) return 0;
    ISE_CHECK_ARG(scores && idx);
    DeviceGuard guard(ctx->device);
    const int64_t total = nq * kk;
    const float fill = metric == ISE_METRIC_IP ? -INFINITY : INFINITY;
    scores_mask_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(scores, nq, nb, idx, kk, id_base, fill);
    ISE_LAUNCH_CHECK();
    return 0;
}
