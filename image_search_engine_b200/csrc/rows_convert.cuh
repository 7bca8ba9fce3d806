// Single-pass conversion of float32 ROWS into per-row-scaled FP16 hi / lo planes (shared by prepare.cu's stand-alone
// kernel and by the converter warps inside gemm_select's fused assign).
#pragma once

#include "common.cuh"

// split one scaled value into its FP16 hi / lo parts
__device__ __forceinline__ void split_f16(float s, __half& h, __half& l) {
    h = __float2half_rn(s);
    l = __float2half_rn(s - __half2float(h));
}

// sums of R per-lane values over the warp with R + log2(32 / R) - 1 shuffles instead of 5 R: every step halves the
// number of values a lane still carries (the upper half of the lanes keeps the upper half of the rows).  Returns the
// total of row `multi_row<R>(lane)`; the lanes with (lane & (32 / R - 1)) == 0 are the designated writers.
template <int R>
__device__ __forceinline__ float multi_sum(float (&a)[R], int lane) {
    int width = 16;
#pragma unroll
    for (int n = R; n > 1; n >>= 1, width >>= 1) {
        const bool up = (lane & width) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? a[i] : a[i + n / 2];
            const float keep = up ? a[i + n / 2] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, width);
        }
    }
    for (; width > 0; width >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], width);
    return a[0];
}
template <int R>
__device__ __forceinline__ int multi_row(int lane) {
    int row = 0, width = 16;
#pragma unroll
    for (int n = R; n > 1; n >>= 1, width >>= 1) row += (lane & width) ? n / 2 : 0;
    return row;
}

// The first version of this kernel ran at 0.56-0.60 of HBM peak and ncu showed why: 146 warp instructions per row
// (ilogbf / ldexpf / a division for the scale, a float NaN test per element, a multiply + three compares per element for
// the exactness test) -- issue-bound, not memory-bound.  Everything that is per-row-uniform is integer bit arithmetic
// now, and the per-element work is four integer min/max/or + one FMA:
//   * |x| as an unsigned bit pattern orders like the value and keeps NaN / Inf visible (>= 0x7F800000), so one integer
//     max gives both the row's absolute maximum (-> its power-of-two scale: exponent arithmetic) and the non-finite flag;
//   * an element is exact in one FP16 plane iff its low 13 mantissa bits are zero (scaling by a power of two does not
//     change them) and it does not fall below FP16's normal range once scaled: min over (|x| bits - 1) against one
//     per-row threshold (zero wraps to 0xFFFFFFFF and never counts).

struct RowStats {            // per-lane accumulators, committed to meta[] once at the end
    bool any_lo = false;
    float max_ss = 0.f;
    float max_res = 0.f;     // largest |x - hi / scale|^2 / |x|^2 of a row that needed its lo plane
    unsigned max_abs_bits = 0u;
};

// meta[LO_NONZERO] doubles as the largest RELATIVE squared hi-plane residual |x - hi / scale|^2 / |x|^2 of a row:
// 0 = the lo plane is all zeros; any other value says it is not, and a value other than the plain flag 1.0 lets the
// coarse-pass error bound use the MEASURED residual instead of the worst case (2^-11)^2 (CoarseBound, common.cuh).
__device__ __forceinline__ void publish_lo_residual(float* meta, bool any_lo, float max_rel_sq, int lane) {
    if (!any_lo || lane != 0) return;
    const float v = fmaxf(max_rel_sq * 1.0002f, 1.17549435e-38f);     // never reads as "no lo plane"
    atomicMax(reinterpret_cast<int*>(meta + META_LO_NONZERO), __float_as_int(v));
}

// the lane that holds the total of row i after multi_sum<R>
template <int R>
__device__ __forceinline__ int multi_lane_of_row(int i) {
    int l = 0, width = 16;
#pragma unroll
    for (int n = R; n > 1; n >>= 1, width >>= 1) {
        if (i >= n / 2) { l += width; i -= n / 2; }
    }
    return l;
}

// One warp converts the ROWS rows [r0, r0 + ROWS) (those below n): see prepare_rows_f32_kernel for the method.
template <int ROWS, int NV>
__device__ __forceinline__ void convert_row_group(const float* __restrict__ x, int64_t n, int d4, int dp4, int64_t ldx,
                                                  __half* __restrict__ hi, __half* __restrict__ lo, int64_t ldp,
                                                  float* __restrict__ norms, float* __restrict__ row_inv,
                                                  uint8_t* __restrict__ lo_skipped, int64_t r0, int lane, RowStats& st) {
    const int my_row = multi_row<ROWS>(lane);
    const bool writer = (lane & (32 / ROWS - 1)) == 0;
    float4 v[ROWS][NV];
#pragma unroll
    for (int i = 0; i < ROWS; ++i)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = lane + 32 * j;
            v[i][j] = (r0 + i < n && c < d4) ? __ldg(reinterpret_cast<const float4*>(x + (r0 + i) * ldx) + c)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    float ss[ROWS];
    unsigned am[ROWS], mant[ROWS], minm1[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
        float s = 0.f;
        unsigned a = 0u, m = 0u, mn = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float4 t = v[i][j];
            const unsigned b0 = __float_as_uint(t.x), b1 = __float_as_uint(t.y), b2 = __float_as_uint(t.z),
                           b3 = __float_as_uint(t.w);
            const unsigned a0 = b0 & 0x7FFFFFFFu, a1 = b1 & 0x7FFFFFFFu, a2 = b2 & 0x7FFFFFFFu, a3 = b3 & 0x7FFFFFFFu;
            s = fmaf(t.x, t.x, s); s = fmaf(t.y, t.y, s); s = fmaf(t.z, t.z, s); s = fmaf(t.w, t.w, s);
            a = max(max(a, a0), max(a1, max(a2, a3)));
            m |= b0 | b1 | b2 | b3;
            mn = min(min(mn, a0 - 1u), min(a1 - 1u, min(a2 - 1u, a3 - 1u)));
        }
        ss[i] = s;
        mant[i] = m;
        minm1[i] = mn;
        am[i] = __reduce_max_sync(0xffffffffu, a);
    }
    const float row_ss = multi_sum<ROWS>(ss, lane);                    // total of row r0 + my_row
    if (writer && r0 + my_row < n) {
        if (norms) norms[r0 + my_row] = row_ss;
        st.max_ss = fmaxf(st.max_ss, row_ss);
    }
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
        const int64_t r = r0 + i;
        if (r >= n) continue;                               // warp-uniform
        st.max_abs_bits = max(st.max_abs_bits, am[i]);      // >= 0x7F800000 <=> the row holds a NaN or an Inf
        // scale = 2^sh puts the row maximum in [2^13, 2^14): sh = 13 - (biased exponent - 127), clamped like
        // scale_from_absmax; an all-zero row keeps scale 1
        int sh = 140 - (int)(am[i] >> 23);
        sh = max(-100, min(100, sh));
        if (am[i] == 0u) sh = 0;
        const float scale = __uint_as_float((unsigned)(127 + sh) << 23);
        const unsigned small_thr = (unsigned)(127 - 14 - sh) << 23;       // bits of 2^-14 / scale
        const bool inexact = ((mant[i] & 0x1FFFu) != 0u) | (minm1[i] < small_thr - 1u);
        const bool row_lo = __any_sync(0xffffffffu, inexact);
        if (!row_lo) {
            // the common case for descriptors (integer-valued SIFT, ORB as float): one packed conversion per two
            // elements, no residual
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int c = lane + 32 * j;
                if (c < dp4) {
                    const float4 t = v[i][j];
                    const __half2 ha = __floats2half2_rn(t.x * scale, t.y * scale);
                    const __half2 hb = __floats2half2_rn(t.z * scale, t.w * scale);
                    reinterpret_cast<uint2*>(hi + r * ldp)[c] =
                        make_uint2(*reinterpret_cast<const uint32_t*>(&ha), *reinterpret_cast<const uint32_t*>(&hb));
                }
            }
        } else {
            float rs = 0.f;                                 // squared residual of the hi plane (scaled units)
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int c = lane + 32 * j;
                if (c < dp4) {                              // pad columns [d, ldp) come out as zeros
                    const float4 t = v[i][j];
                    __half h0, h1, h2, h3, l0, l1, l2, l3;
                    split_f16(t.x * scale, h0, l0); split_f16(t.y * scale, h1, l1);
                    split_f16(t.z * scale, h2, l2); split_f16(t.w * scale, h3, l3);
                    const float e0 = t.x * scale - __half2float(h0), e1 = t.y * scale - __half2float(h1);
                    const float e2 = t.z * scale - __half2float(h2), e3 = t.w * scale - __half2float(h3);
                    rs = fmaf(e0, e0, rs); rs = fmaf(e1, e1, rs); rs = fmaf(e2, e2, rs); rs = fmaf(e3, e3, rs);
                    const __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3);
                    const __half2 la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
                    reinterpret_cast<uint2*>(hi + r * ldp)[c] =
                        make_uint2(*reinterpret_cast<const uint32_t*>(&ha), *reinterpret_cast<const uint32_t*>(&hb));
                    if (lo)
                        reinterpret_cast<uint2*>(lo + r * ldp)[c] =
                            make_uint2(*reinterpret_cast<const uint32_t*>(&la), *reinterpret_cast<const uint32_t*>(&lb));
                }
            }
            // relative to the row's squared norm (held by the lane that owns row i after multi_sum)
            rs = warp_sum(rs);
            const float ss_i = __shfl_sync(0xffffffffu, row_ss, multi_lane_of_row<ROWS>(i));
            const float inv_s = __uint_as_float((unsigned)(127 - sh) << 23);
            if (ss_i > 0.f) st.max_res = fmaxf(st.max_res, rs * inv_s * inv_s / ss_i);
        }
        if (lane == 0) {
            row_inv[r] = __uint_as_float((unsigned)(127 - sh) << 23);
            if (lo_skipped) lo_skipped[r] = (lo && !row_lo) ? 1 : 0;
        }
        st.any_lo |= row_lo;
    }
}

// whole warp; lane 0 publishes the warp's statistics
__device__ __forceinline__ void commit_row_stats(RowStats& st, float* meta, int lane) {
    st.max_ss = warp_max(st.max_ss);
    if (lane == 0) {
        if (st.max_ss > 0.f) atomicMax(reinterpret_cast<int*>(meta + META_MAX_NORM_SQ), __float_as_int(st.max_ss));
        if (st.max_abs_bits >= 0x7F800000u) meta[META_NONFINITE] = 1.f;
        else if (st.max_abs_bits != 0u) atomicMax(reinterpret_cast<int*>(meta + META_ABSMAX), (int)st.max_abs_bits);
    }
    publish_lo_residual(meta, __any_sync(0xffffffffu, st.any_lo), warp_max(st.max_res), lane);
}
