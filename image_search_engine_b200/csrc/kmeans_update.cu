// k-means centroid update: the device half of faiss Clustering.cpp compute_centroids /
// split_clusters / post_process_centroids (reached from kmeans_faiss.py:41).
//
// HBM-bound scatter-add: one warp per descriptor row, 128-bit row loads, 128-bit vector
// reductions (red.global.add.v4.f32) into the FP32 [k, d] sum matrix that lives in L2.
// Sums are accumulated in FP32 like Faiss; the order differs (atomics vs Faiss's data order),
// which stays inside the 1e-4 relative centroid tolerance of the parity contract.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// cent != nullptr: the objective term of each row is recomputed here in exact FP32 from the row that is
// already in registers and its centroid (L2-resident): <x, c> (spherical / IP) or sum (x - c)^2 (L2),
// instead of trusting a distance that came out of the tensor-core accumulator.
template <typename T>
__global__ void accumulate_kernel(const T* __restrict__ x, int64_t n, int d, int64_t ldx,
                                  const int64_t* __restrict__ assign, const float* __restrict__ dis,
                                  const float* __restrict__ cent, int metric,
                                  float* __restrict__ sums, float* __restrict__ counts, double* __restrict__ obj) {
    __shared__ double s_obj[kWarps];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * kWarps + wib;
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    const bool vec4 = sizeof(T) == 4 && (d % 4 == 0) && (ldx % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(sums) & 15) == 0);
    double my_obj = 0.0;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int64_t a = __ldg(assign + r);
        if (a < 0) continue;
        float* dst = sums + a * (int64_t)d;
        const T* row = x + r * ldx;
        const float* crow = cent ? cent + a * (int64_t)d : nullptr;
        float acc = 0.f;
        if (vec4) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            for (int c = lane; c < d / 4; c += 32) {
                const float4 v = __ldg(row4 + c);
                red_add_v4(dst + 4 * c, v);
                if (crow) {
                    const float4 y = __ldg(reinterpret_cast<const float4*>(crow) + c);
                    if (metric == ISE_METRIC_IP) {
                        acc = fmaf(v.x, y.x, acc); acc = fmaf(v.y, y.y, acc);
                        acc = fmaf(v.z, y.z, acc); acc = fmaf(v.w, y.w, acc);
                    } else {
                        const float e0 = v.x - y.x, e1 = v.y - y.y, e2 = v.z - y.z, e3 = v.w - y.w;
                        acc = fmaf(e0, e0, acc); acc = fmaf(e1, e1, acc);
                        acc = fmaf(e2, e2, acc); acc = fmaf(e3, e3, acc);
                    }
                }
            }
        } else {
            for (int c = lane; c < d; c += 32) {
                const float v = (float)row[c];
                atomicAdd(dst + c, v);
                if (crow) {
                    const float y = __ldg(crow + c);
                    if (metric == ISE_METRIC_IP) acc = fmaf(v, y, acc);
                    else { const float e = v - y; acc = fmaf(e, e, acc); }
                }
            }
        }
        if (crow) acc = warp_sum(acc);
        if (lane == 0) {
            atomicAdd(counts + a, 1.0f);
            if (crow) my_obj += (double)acc;
            else if (dis) my_obj += (double)__ldg(dis + r);
        }
    }
    if (lane == 0) s_obj[wib] = my_obj;
    __syncthreads();
    if (threadIdx.x == 0 && obj) {
        double t = 0.0;
        for (int i = 0; i < kWarps; ++i) t += s_obj[i];
        if (t != 0.0) atomicAdd(obj, t);
    }
}

__global__ void mean_kernel(const float* __restrict__ sums, const float* __restrict__ counts, int64_t k, int d,
                            float* __restrict__ centroids, int32_t* __restrict__ n_empty) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    int empties = 0;
    for (int64_t c = warp; c < k; c += nwarps) {
        const float h = counts[c];
        if (h == 0.f) {
            for (int j = lane; j < d; j += 32) centroids[c * d + j] = 0.f;
            empties++;
        } else {
            const float norm = 1.f / h;  // float norm = 1 / hassign[ci]; c[j] *= norm
            for (int j = lane; j < d; j += 32) centroids[c * d + j] = sums[c * d + j] * norm;
        }
    }
    if (lane == 0 && empties) atomicAdd(n_empty, empties);
}

// split_clusters: the (ci, cj) plan is ordered and later steps may read rows written by earlier ones,
// so one block walks the plan sequentially and parallelises over the dimension.
__global__ void apply_splits_kernel(float* __restrict__ centroids, int d, const int32_t* __restrict__ pairs,
                                    int nsplit) {
    const float up = 1.f + 1.f / 1024.f, dn = 1.f - 1.f / 1024.f;  // EPS = 1/1024
    for (int s = 0; s < nsplit; ++s) {
        float* ci = centroids + (int64_t)pairs[2 * s] * d;
        float* cj = centroids + (int64_t)pairs[2 * s + 1] * d;
        for (int j = threadIdx.x; j < d; j += blockDim.x) {
            const float v = cj[j];
            if ((j & 1) == 0) { ci[j] = v * up; cj[j] = v * dn; }
            else { ci[j] = v * dn; cj[j] = v * up; }
        }
        __syncthreads();
    }
}

}  // namespace

ISE_EXPORT int ise_kmeans_accumulate(ise_ctx* ctx, const void* x, int dtype, int64_t n, int d, int64_t ldx,
                                     const int64_t* assign, const float* dis, const float* centroids, int metric,
                                     float* sums, float* counts, double* obj, void* stream) {
    ISE_CHECK_ARG(ctx && n >= 0 && d > 0 && ldx >= d);
    ISE_CHECK_ARG(dtype == ISE_DTYPE_F32 || dtype == ISE_DTYPE_U8);
    if (n == 0) return 0;
    ISE_CHECK_ARG(x && assign && sums && counts);
    ISE_CHECK_ARG(metric == ISE_METRIC_IP || metric == ISE_METRIC_L2);
    if (centroids) ISE_CHECK_ARG((reinterpret_cast<uintptr_t>(centroids) & 15) == 0);
    DeviceGuard guard(ctx->device);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n, kWarps), (int64_t)ctx->sm_count * 8));
    if (dtype == ISE_DTYPE_F32)
        accumulate_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const float*)x, n, d, ldx, assign, dis,
                                                                             centroids, metric, sums, counts, obj);
    else
        accumulate_kernel<uint8_t><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const uint8_t*)x, n, d, ldx, assign,
                                                                               dis, centroids, metric, sums, counts, obj);
    ISE_LAUNCH_CHECK();
    return 0;
}

ISE_EXPORT int ise_kmeans_mean(ise_ctx* ctx, const float* sums, const float* counts, int64_t k, int d,
                               float* centroids, int32_t* n_empty, void* stream) {
    ISE_CHECK_ARG(ctx && sums && counts && centroids && n_empty && k > 0 && d > 0);
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ISE_CUDA(cudaMemsetAsync(n_empty, 0, sizeof(int32_t), st));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(k, kWarps), (int64_t)ctx->sm_count * 8));
    mean_kernel<<<grid, kThreads, 0, st>>>(sums, counts, k, d, centroids, n_empty);
    ISE_LAUNCH_CHECK();
    return 0;
}

ISE_EXPORT int ise_kmeans_apply_splits(ise_ctx* ctx, float* centroids, int64_t k, int d, const int32_t* pairs,
                                       int32_t nsplit, void* stream) {
    ISE_CHECK_ARG(ctx && centroids && k > 0 && d > 0 && nsplit >= 0);
    if (nsplit == 0) return 0;
    ISE_CHECK_ARG(pairs != nullptr);
    DeviceGuard guard(ctx->device);
    apply_splits_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(centroids, d, pairs, nsplit);
    ISE_LAUNCH_CHECK();
    return 0;
}
