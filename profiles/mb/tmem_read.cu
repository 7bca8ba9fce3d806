// TMEM read-out microbenchmark (B200): how many bytes per cycle per SM can epilogue warps pull out of
// tensor memory with tcgen05.ld, by load width, loads in flight and warps per lane quadrant?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_read tmem_read.cu ; run: ./tmem_read
#include <cstdio>
#include <cuda_runtime.h>
#include "tmem_ld_gen.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int N> __device__ __forceinline__ float tree_max(const uint32_t (&r)[N]) {
    float t[N];
#pragma unroll
    for (int j = 0; j < N; ++j) t[j] = __uint_as_float(r[j]);
#pragma unroll
    for (int s = N / 2; s > 0; s >>= 1)
#pragma unroll
        for (int j = 0; j < s; ++j) t[j] = fmaxf(t[j], t[j + s]);
    return t[0];
}
template <int N> __device__ __forceinline__ void ld(uint32_t a, uint32_t (&r)[N]) {
    if constexpr (N == 16) tmem_ld_x16(a, r);
    if constexpr (N == 32) tmem_ld_x32(a, r);
    if constexpr (N == 64) tmem_ld_x64(a, r);
    if constexpr (N == 128) tmem_ld_x128(a, r);
}

// MODE 0: ld, wait, consume (serial).  MODE 1: two loads in flight (ping-pong register sets).
// WPQ = warps per lane quadrant (each scans 256 / WPQ columns of the 128 x 256 fp32 tile per iteration)
template <int N, int MODE, int WPQ>
__global__ void __launch_bounds__(128 * WPQ, 1) k(int iters, float* out, long long* cyc) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3, part = warp >> 2;
    constexpr int COLS = 256 / WPQ;
    const uint32_t base = tbase + ((uint32_t)(q * 32) << 16) + part * COLS;
    float acc = -1e30f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t a = base + (it & 1) * 256;
        if (MODE == 0) {
#pragma unroll 1
            for (int c = 0; c < COLS; c += N) {
                uint32_t r[N];
                ld<N>(a + c, r);
                ld_wait();
                acc = fmaxf(acc, tree_max<N>(r));
            }
        } else {
            uint32_t r0[N], r1[N];
            ld<N>(a, r0);
#pragma unroll 1
            for (int c = 0; c < COLS; c += 2 * N) {
                ld<N>(a + c + N, r1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");   // waits for both (no partial wait exists)
                acc = fmaxf(acc, tree_max<N>(r0));
                if (c + 2 * N < COLS) ld<N>(a + c + 2 * N, r0);
                acc = fmaxf(acc, tree_max<N>(r1));
                if (c + 2 * N < COLS) ld_wait();
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}

template <int N, int MODE, int WPQ> void run(const char* name) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<N, MODE, WPQ><<<148, 128 * WPQ>>>(10, out, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<N, MODE, WPQ><<<148, 128 * WPQ>>>(iters, out, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %8.1f cycles / 128x256 fp32 tile  = %6.1f B/cycle/SM   (%.3f ms, %s)\n", name, (double)c / iters,
           131072.0 * iters / (double)c, ms, cudaGetErrorString(err));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<16, 0, 1>("x16  serial, 4 warps");
    run<32, 0, 1>("x32  serial, 4 warps (current epilogue)");
    run<64, 0, 1>("x64  serial, 4 warps");
    run<128, 0, 1>("x128 serial, 4 warps");
    run<32, 1, 1>("x32  2 in flight, 4 warps");
    run<64, 1, 1>("x64  2 in flight, 4 warps");
    run<32, 0, 2>("x32  serial, 8 warps (2 per quadrant)");
    run<64, 0, 2>("x64  serial, 8 warps (2 per quadrant)");
    run<32, 1, 2>("x32  2 in flight, 8 warps");
    run<32, 0, 4>("x32  serial, 16 warps (4 per quadrant)");
    return 0;
}
