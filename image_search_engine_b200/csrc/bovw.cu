// BoVW histogram (+ fused Okapi/BM25 tf weighting): one CTA per image, shared-memory-privatised
// counts, 128-bit id loads, vectorised row stores.  HBM-bound: 8 B per descriptor id read plus one
// write of the [n_img, k] matrix.
// Replaces create_visual_word_histogram (bag_of_visual_words.py:98-106: np.histogram per image into a
// float64 matrix) and OkapiTransformer.transform (utils.py:153-202).
#include <climits>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxSmemBins = 12 * 1024;  // 48 KiB of int32 counters

// np.histogram(a, bins=k) for integer input, restated with the same float64 operations
// (numpy/lib/_histograms_impl.py: _get_outer_edges, linspace edges, f_indices + edge corrections).
struct NumpyBins {
    double first, last, denom, step;
    int k;
    __device__ __forceinline__ void setup(int64_t mn, int64_t mx, int k_) {
        k = k_;
        if (mn == mx) { first = (double)mn - 0.5; last = (double)mx + 0.5; }
        else { first = (double)mn; last = (double)mx; }
        denom = __dsub_rn(last, first);       // _unsigned_subtract(last_edge, first_edge)
        step = __ddiv_rn(denom, (double)k);   // linspace: step = delta / div
    }
    __device__ __forceinline__ double edge(int i) const {  // linspace(first, last, k + 1)[i]
        if (i == k) return last;
        return __dadd_rn(__dmul_rn((double)i, step), first);
    }
    __device__ __forceinline__ int bin(int64_t a) const {
        const double x = (double)a;
        const double f = __dmul_rn(__ddiv_rn(__dsub_rn(x, first), denom), (double)k);
        int idx = (int)f;  // astype(np.intp) truncation
        if (idx == k) idx -= 1;
        if (x < edge(idx)) idx -= 1;
        if (x >= edge(idx + 1) && idx != k - 1) idx += 1;
        return idx;
    }
};

// tf*k1 / (tf*k1 + k2*(1 - b + b*(dl/avgdl))) with numpy's float64 operation order (utils.py:199-200)
__device__ __forceinline__ double okapi_weight(double tf, double k1, double k2, double b, double ratio) {
    const double t = __dmul_rn(tf, k1);
    const double den = __dadd_rn(t, __dmul_rn(k2, __dadd_rn(__dsub_rn(1.0, b), __dmul_rn(b, ratio))));
    return __ddiv_rn(t, den);
}

template <typename OutT, bool SMEM>
__global__ void histogram_kernel(const int64_t* __restrict__ words, const int64_t* __restrict__ off, int64_t n_img,
                                 int k, int mode, OutT* __restrict__ out, int okapi, double k1, double k2, double b,
                                 double avgdl_in) {
    extern __shared__ int s_cnt[];
    __shared__ long long s_mn, s_mx;
    // np.mean(dl) over the batch being transformed (utils.py:196); a caller that feeds the batch in
    // several launches passes the whole batch's mean explicitly
    const double avgdl = avgdl_in >= 0.0 ? avgdl_in : (double)(off[n_img] - off[0]) / (double)n_img;
    for (int64_t img = blockIdx.x; img < n_img; img += gridDim.x) {
        const int64_t lo = off[img], hi = off[img + 1];
        const int64_t cnt = hi - lo;
        OutT* orow = out + img * (int64_t)k;
        if (SMEM) {
            for (int j = threadIdx.x; j < k; j += kThreads) s_cnt[j] = 0;
        }
        if (threadIdx.x == 0) { s_mn = LLONG_MAX; s_mx = LLONG_MIN; }
        __syncthreads();
        NumpyBins nb;
        if (mode == ISE_HIST_NUMPY_COMPAT && cnt > 0) {
            long long mn = LLONG_MAX, mx = LLONG_MIN;
            for (int64_t j = lo + threadIdx.x; j < hi; j += kThreads) {
                const long long w = words[j];
                mn = min(mn, w); mx = max(mx, w);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            if ((threadIdx.x & 31) == 0) { atomicMin(&s_mn, mn); atomicMax(&s_mx, mx); }
            __syncthreads();
            nb.setup(s_mn, s_mx, k);
        }
        for (int64_t j = lo + threadIdx.x; j < hi; j += kThreads) {
            const int64_t w = words[j];
            int bin;
            if (mode == ISE_HIST_NUMPY_COMPAT) bin = nb.bin(w);
            else bin = (w >= 0 && w < k) ? (int)w : -1;
            if (bin >= 0 && bin < k) {
                if (SMEM) atomicAdd(&s_cnt[bin], 1);
                else atomicAdd(orow + bin, (OutT)1);  // row was zeroed by the caller's memset
            }
        }
        __syncthreads();
        const double ratio = __ddiv_rn((double)cnt, avgdl);  // rep / avgdl
        if (SMEM) {
            for (int j = threadIdx.x; j < k; j += kThreads) {
                const int c = s_cnt[j];
                double v = (double)c;
                if (okapi && c != 0) v = okapi_weight(v, k1, k2, b, ratio);
                orow[j] = (OutT)v;
            }
        } else if (okapi) {
            for (int j = threadIdx.x; j < k; j += kThreads) {
                const double c = (double)orow[j];
                if (c != 0.0) orow[j] = (OutT)okapi_weight(c, k1, k2, b, ratio);
            }
        }
        __syncthreads();
    }
}

// dense in-place Okapi: pass 1 row sums (document lengths), pass 2 weights
template <typename T>
__global__ void row_sum_kernel(const T* __restrict__ h, int64_t n_img, int k, double* __restrict__ dl,
                               double* __restrict__ total) {
    __shared__ double s_part[kThreads / 32];
    for (int64_t img = blockIdx.x; img < n_img; img += gridDim.x) {
        const T* row = h + img * (int64_t)k;
        double acc = 0.0;
        for (int j = threadIdx.x; j < k; j += kThreads) acc += (double)row[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < kThreads / 32; ++i) t += s_part[i];
            dl[img] = t;
            atomicAdd(total, t);
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void okapi_dense_kernel(T* __restrict__ h, int64_t n_img, int k, double k1, double k2, double b,
                                   double avgdl_in, const double* __restrict__ dl, const double* __restrict__ total) {
    const double avgdl = avgdl_in >= 0.0 ? avgdl_in : __ddiv_rn(*total, (double)n_img);
    for (int64_t img = blockIdx.x; img < n_img; img += gridDim.x) {
        T* row = h + img * (int64_t)k;
        const double ratio = __ddiv_rn(dl[img], avgdl);
        for (int j = threadIdx.x; j < k; j += kThreads) {
            const double c = (double)row[j];
            if (c != 0.0) row[j] = (T)okapi_weight(c, k1, k2, b, ratio);
        }
    }
}

}  // namespace

ISE_EXPORT int ise_bovw_histogram(ise_ctx* ctx, const int64_t* words, const int64_t* img_offsets, int64_t n_img,
                                  int k, int mode, int out_dtype, void* out, int okapi, double k1, double k2,
                                  double b, double avgdl, void* stream) {
    ISE_CHECK_ARG(ctx && n_img >= 0 && k >= 1);
    ISE_CHECK_ARG(mode == ISE_HIST_NUMPY_COMPAT || mode == ISE_HIST_BINCOUNT);
    ISE_CHECK_ARG(out_dtype == ISE_OUT_F32 || out_dtype == ISE_OUT_F64);
    if (n_img == 0) return 0;
    ISE_CHECK_ARG(img_offsets && out);
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const bool smem = k <= kMaxSmemBins;
    const int grid = (int)std::min<int64_t>(n_img, (int64_t)ctx->sm_count * 8);
    const size_t elt = out_dtype == ISE_OUT_F64 ? 8 : 4;
    if (!smem) ISE_CUDA(cudaMemsetAsync(out, 0, (size_t)n_img * k * elt, st));
    const size_t shm = smem ? (size_t)k * sizeof(int) : 0;
    if (out_dtype == ISE_OUT_F64) {
        if (smem) histogram_kernel<double, true><<<grid, kThreads, shm, st>>>(words, img_offsets, n_img, k, mode,
                                                                              (double*)out, okapi, k1, k2, b, avgdl);
        else histogram_kernel<double, false><<<grid, kThreads, 0, st>>>(words, img_offsets, n_img, k, mode,
                                                                        (double*)out, okapi, k1, k2, b, avgdl);
    } else {
        if (smem) histogram_kernel<float, true><<<grid, kThreads, shm, st>>>(words, img_offsets, n_img, k, mode,
                                                                             (float*)out, okapi, k1, k2, b, avgdl);
        else histogram_kernel<float, false><<<grid, kThreads, 0, st>>>(words, img_offsets, n_img, k, mode,
                                                                       (float*)out, okapi, k1, k2, b, avgdl);
    }
    ISE_LAUNCH_CHECK();
    return 0;
}

ISE_EXPORT int ise_okapi_tf(ise_ctx* ctx, void* h, int out_dtype, int64_t n_img, int k, double k1, double k2,
                            double b, double avgdl, double* dl_workspace, void* stream) {
    ISE_CHECK_ARG(ctx && n_img >= 0 && k >= 1);
    ISE_CHECK_ARG(out_dtype == ISE_OUT_F32 || out_dtype == ISE_OUT_F64);
    if (n_img == 0) return 0;
    ISE_CHECK_ARG(h && dl_workspace);
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    double* dl = dl_workspace;
    double* total = dl_workspace + n_img;
    ISE_CUDA(cudaMemsetAsync(total, 0, sizeof(double), st));
    const int grid = (int)std::min<int64_t>(n_img, (int64_t)ctx->sm_count * 8);
    if (out_dtype == ISE_OUT_F64) {
        row_sum_kernel<double><<<grid, kThreads, 0, st>>>((const double*)h, n_img, k, dl, total);
        ISE_LAUNCH_CHECK();
        okapi_dense_kernel<double><<<grid, kThreads, 0, st>>>((double*)h, n_img, k, k1, k2, b, avgdl, dl, total);
    } else {
        row_sum_kernel<float><<<grid, kThreads, 0, st>>>((const float*)h, n_img, k, dl, total);
        ISE_LAUNCH_CHECK();
        okapi_dense_kernel<float><<<grid, kThreads, 0, st>>>((float*)h, n_img, k, k1, k2, b, avgdl, dl, total);
    }
    ISE_LAUNCH_CHECK();
    return 0;
}
