// Operand preparation: FP32 / uint8 rows -> scaled FP16 hi/lo planes + exact FP32 row norms,
// and faiss.normalize_L2.  HBM-bound streaming kernels: one warp per row, 128-bit loads.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;

__global__ void absmax_f32_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ldx, float* meta) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    float m = 0.f;
    const bool vec = (d % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int64_t r = warp; r < n; r += nwarps) {
        const float* row = x + r * ldx;
        if (vec) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            for (int c = lane; c < d / 4; c += 32) {
                float4 v = __ldg(row4 + c);
                m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            }
        } else {
            for (int c = lane; c < d; c += 32) m = fmaxf(m, fabsf(__ldg(row + c)));
        }
    }
    m = warp_max(m);
    // non-negative finite floats order like their bit patterns; NaN/Inf are rejected by the host API
    if (lane == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(meta + META_ABSMAX), __float_as_int(m));
}

__device__ __forceinline__ float scale_from_absmax(float absmax) {
    if (!(absmax > 0.f)) return 1.f;
    int e = ilogbf(absmax);  // absmax in [2^e, 2^(e+1))
    int s = 13 - e;          // absmax * 2^s in [2^13, 2^14): inside fp16 range with headroom
    s = max(-100, min(100, s));
    return ldexpf(1.f, s);
}

template <typename T>
__device__ __forceinline__ float load_as_f32(const T* p);
template <>
__device__ __forceinline__ float load_as_f32<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_as_f32<uint8_t>(const uint8_t* p) { return (float)__ldg(p); }

template <typename T>
__global__ void prepare_planes_kernel(const T* __restrict__ x, int64_t n, int d, int64_t ldx,
                                      __half* __restrict__ hi, __half* __restrict__ lo, int64_t ldp,
                                      float* __restrict__ norms, float* meta, bool fixed_unit_scale) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const float scale = fixed_unit_scale ? 1.f : scale_from_absmax(meta[META_ABSMAX]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        meta[META_SCALE] = scale;
        meta[META_INV_SCALE] = 1.f / scale;
        if (fixed_unit_scale) meta[META_ABSMAX] = 255.f;
    }
    bool any_lo = false;
    float max_ss = 0.f;
    const int dp = (int)ldp;
    for (int64_t r = warp; r < n; r += nwarps) {
        const T* row = x + r * ldx;
        __half* hrow = hi + r * ldp;
        __half* lrow = lo ? lo + r * ldp : nullptr;
        float ss = 0.f;
        // two columns per lane per step so the plane stores are 32-bit
        for (int c = lane * 2; c < dp; c += 64) {
            float v0 = (c < d) ? load_as_f32<T>(row + c) : 0.f;
            float v1 = (c + 1 < d) ? load_as_f32<T>(row + c + 1) : 0.f;
            ss = fmaf(v0, v0, ss);
            ss = fmaf(v1, v1, ss);
            float s0 = v0 * scale, s1 = v1 * scale;
            __half h0 = __float2half_rn(s0), h1 = __float2half_rn(s1);
            __half l0 = __float2half_rn(s0 - __half2float(h0));
            __half l1 = __float2half_rn(s1 - __half2float(h1));
            *reinterpret_cast<__half2*>(hrow + c) = __halves2half2(h0, h1);
            if (lrow) *reinterpret_cast<__half2*>(lrow + c) = __halves2half2(l0, l1);
            any_lo |= (__half2float(l0) != 0.f) | (__half2float(l1) != 0.f);
        }
        ss = warp_sum(ss);
        if (norms && lane == 0) norms[r] = ss;
        max_ss = fmaxf(max_ss, ss);
    }
    // largest row norm^2 (non-negative floats order like ints; one atomic per warp): the coarse-pass
    // error bound of ise_rescore_select needs it
    if (lane == 0 && max_ss > 0.f) atomicMax(reinterpret_cast<int*>(meta + META_MAX_NORM_SQ), __float_as_int(max_ss));
    if (__any_sync(0xffffffffu, any_lo) && lane == 0) meta[META_LO_NONZERO] = 1.f;
}

__global__ void normalize_l2_kernel(float* __restrict__ x, int64_t n, int d) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int64_t r = warp; r < n; r += nwarps) {
        float* row = x + r * (int64_t)d;
        float ss = 0.f;
        if (vec) {
            float4* row4 = reinterpret_cast<float4*>(row);
            for (int c = lane; c < d / 4; c += 32) {
                float4 v = row4[c];
                ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss);
                ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
            }
        } else {
            for (int c = lane; c < d; c += 32) { float v = row[c]; ss = fmaf(v, v, ss); }
        }
        ss = warp_sum(ss);
        if (ss > 0.f) {
            const float inv = 1.0f / sqrtf(ss);  // fvec_renorm_L2: inv_nr = 1.0 / sqrtf(nr)
            if (vec) {
                float4* row4 = reinterpret_cast<float4*>(row);
                for (int c = lane; c < d / 4; c += 32) {
                    float4 v = row4[c];
                    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                    row4[c] = v;
                }
            } else {
                for (int c = lane; c < d; c += 32) row[c] *= inv;
            }
        }
    }
}

int grid_for_rows(const ise_ctx* ctx, int64_t n) {
    int64_t blocks = ceil_div64(n, kWarpsPerBlock);
    int64_t cap = (int64_t)ctx->sm_count * 8;  // 8 resident 256-thread CTAs per SM
    return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

}  // namespace

ISE_EXPORT int ise_prepare_planes(ise_ctx* ctx, const void* x, int dtype, int64_t n, int d, int64_t ldx,
                                  void* hi, void* lo, int64_t ldp, float* norms, float* meta, void* stream) {
    ISE_CHECK_ARG(ctx && meta && hi);
    ISE_CHECK_ARG(n >= 0 && d > 0 && ldx >= d && ldp >= d && ldp % 8 == 0);
    ISE_CHECK_ARG(dtype == ISE_DTYPE_F32 || dtype == ISE_DTYPE_U8);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ISE_CUDA(cudaMemsetAsync(meta, 0, META_FLOATS * sizeof(float), st));
    if (n == 0) return 0;
    ISE_CHECK_ARG(x != nullptr);
    const int grid = grid_for_rows(ctx, n);
    if (dtype == ISE_DTYPE_F32) {
        absmax_f32_kernel<<<grid, kThreads, 0, st>>>((const float*)x, n, d, ldx, meta);
        ISE_LAUNCH_CHECK();
        prepare_planes_kernel<float><<<grid, kThreads, 0, st>>>((const float*)x, n, d, ldx, (__half*)hi,
                                                                (__half*)lo, ldp, norms, meta, false);
    } else {
        prepare_planes_kernel<uint8_t><<<grid, kThreads, 0, st>>>((const uint8_t*)x, n, d, ldx, (__half*)hi,
                                                                  (__half*)lo, ldp, norms, meta, true);
    }
    ISE_LAUNCH_CHECK();
    return 0;
}

ISE_EXPORT int ise_normalize_l2(ise_ctx* ctx, float* x, int64_t n, int d, void* stream) {
    ISE_CHECK_ARG(ctx && d > 0 && n >= 0);
    if (n == 0) return 0;
    ISE_CHECK_ARG(x != nullptr);
    DeviceGuard g(ctx->device);
    normalize_l2_kernel<<<grid_for_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(x, n, d);
    ISE_LAUNCH_CHECK();
    return 0;
}
