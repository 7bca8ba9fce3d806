// IVFPQ ("cell-probe" index type, utils.py:311-325: IndexIVFPQ(IndexFlatL2(d), d, 8, 16, 8), nprobe = 5).
// Off every default path of the reference; built from the pieces of the hot path: the coarse quantizer and the
// 16 sub-quantizers are trained by the k-means kernels, vectors are encoded by the fused top-1 assign, and only the
// two steps below are specific: residuals, and the asymmetric-distance scan of the probed inverted lists.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void residual_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int d,
                                const float* __restrict__ cent, const int64_t* __restrict__ assign,
                                float* __restrict__ out) {
    const int64_t total = n * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        const int c = (int)(i - r * d);
        const int64_t a = assign[r];
        out[i] = x[r * ldx + c] - (a >= 0 ? cent[a * d + c] : 0.f);     // compute_residual: x - reconstruct(key)
    }
}

// One CTA per (query, probe): look-up table tab[m][code] = |(q - c_list)_m - pq[m][code]|^2 in shared memory, then
// every code of the list costs M table reads: dis = sum_m tab[m][code_m], summed in m order like Faiss's scanner.
// dist is [nq, ntotal] in list-sorted order, pre-filled with +inf (entries of lists that are not probed).
__global__ void ivfpq_scan_kernel(const float* __restrict__ q, int d, const float* __restrict__ coarse,
                                  const int64_t* __restrict__ probes, int nprobe, const float* __restrict__ pq,
                                  int M, int ksub, const uint8_t* __restrict__ codes,
                                  const int64_t* __restrict__ list_off, int64_t ntotal, float* __restrict__ dist) {
    extern __shared__ float sm[];
    float* res = sm;                 // [d]
    float* tab = sm + d;             // [M * ksub]
    const int64_t qi = blockIdx.y;
    const int64_t key = probes[qi * nprobe + blockIdx.x];
    if (key < 0) return;
    const int64_t lo = list_off[key], hi = list_off[key + 1];
    if (lo == hi) return;
    const int dsub = d / M;
    for (int c = threadIdx.x; c < d; c += kThreads) res[c] = q[qi * d + c] - coarse[key * d + c];
    __syncthreads();
    for (int e = threadIdx.x; e < M * ksub; e += kThreads) {
        const int m = e / ksub;
        const float* cj = pq + (int64_t)e * dsub;
        const float* rj = res + m * dsub;
        float acc = 0.f;
        for (int j = 0; j < dsub; ++j) {
            const float t = rj[j] - cj[j];
            acc = fmaf(t, t, acc);
        }
        tab[e] = acc;
    }
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) {
        const uint8_t* code = codes + i * M;
        float dis = 0.f;
        for (int m = 0; m < M; ++m) dis += tab[m * ksub + code[m]];
        dist[qi * ntotal + i] = dis;
    }
}

}  // namespace

ISE_EXPORT int ise_ivfpq_residual(ise_ctx* ctx, const float* x, int64_t ldx, int64_t n, int d, const float* centroids,
                                  const int64_t* assign, float* out, void* stream) {
    ISE_CHECK_ARG(ctx && n >= 0 && d > 0 && ldx >= d);
    if (n == 0) return 0;
    ISE_CHECK_ARG(x && centroids && assign && out);
    DeviceGuard guard(ctx->device);
    const int grid = (int)std::min<int64_t>(ceil_div64(n * d, kThreads), (int64_t)ctx->sm_count * 8);
    residual_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, ldx, n, d, centroids, assign, out);
    ISE_LAUNCH_CHECK();
    return 0;
}

ISE_EXPORT int ise_ivfpq_scan(ise_ctx* ctx, const float* q, int64_t nq, int d, const float* coarse_centroids,
                              int64_t nlist, const int64_t* probes, int nprobe, const float* pq_centroids, int M,
                              int ksub, const uint8_t* codes, const int64_t* list_offsets, int64_t ntotal,
                              float* dist, void* stream) {
    ISE_CHECK_ARG(ctx && nq >= 0 && d > 0 && nlist > 0 && nprobe >= 1 && M >= 1 && d % M == 0 && ksub >= 1 && ksub <= 256);
    ISE_CHECK_ARG(nq <= 65535);
    if (nq == 0 || ntotal == 0) return 0;
    ISE_CHECK_ARG(q && coarse_centroids && probes && pq_centroids && codes && list_offsets && dist);
    const size_t shm = ((size_t)d + (size_t)M * ksub) * sizeof(float);
    ISE_CHECK_ARG(shm <= 200 * 1024);
    DeviceGuard guard(ctx->device);
    ISE_CUDA(cudaFuncSetAttribute(ivfpq_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    dim3 grid((unsigned)nprobe, (unsigned)nq);
    ivfpq_scan_kernel<<<grid, kThreads, shm, (cudaStream_t)stream>>>(q, d, coarse_centroids, probes, nprobe, pq_centroids, M,
                                                                    ksub, codes, list_offsets, ntotal, dist);
    ISE_LAUNCH_CHECK();
    return 0;
}
