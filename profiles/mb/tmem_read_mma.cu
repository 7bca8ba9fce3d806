// TMEM read-out under MMA load (B200): epilogue warps tcgen05.ld one 128 x 256 fp32 accumulator while one
// thread keeps issuing tcgen05.mma (M128 N256 K16, kind::f16) into the OTHER accumulator, as in gemm_select.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../image_search_engine_b200/csrc -o tmem_read_mma tmem_read_mma.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
#include "tmem_ld_gen.cuh"

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
    if constexpr (N == 32) tmem_ld_x32(a, r);
    if constexpr (N == 64) tmem_ld_x64(a, r);
}

// WPQ epilogue warps per quadrant; warp 4*WPQ = MMA issuer.  do_mma / do_read select the load.
template <int N, int WPQ>
__global__ void __launch_bounds__(128 * WPQ + 32, 1) k(int tiles, int mma_per_tile, int do_mma, int do_read, float* out,
                                                       long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tbase;
    __shared__ uint64_t bar[2];
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
    if (threadIdx.x == 0) { ptx::mbar_init(&bar[0], 1); ptx::mbar_init(&bar[1], 1); done = 0; ptx::fence_barrier_init(); }
    ptx::fence_proxy_async();
    if (warp == 0) ptx::tmem_alloc(&tbase, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    float acc = -1e30f;
    int nread = 0;
    const long long t0 = clock64();
    if (warp == 4 * WPQ) {
        if (lane == 0 && do_mma) {
            constexpr uint32_t idesc = ptx::make_idesc_f16_f32(128, 256);
            const uint64_t da = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
            const uint64_t db = ptx::make_smem_desc_sw128(ptx::smem_u32(smem) + 16384);
            for (int t = 0; t < tiles; ++t) {
                for (int i = 0; i < mma_per_tile; ++i)
                    ptx::umma_f16_ss(tbase + 256, da + (uint64_t)((i & 3) * 2), db + (uint64_t)((i & 3) * 2), idesc, i != 0);
                ptx::umma_commit(&bar[t & 1]);
                if (t >= 1) ptx::mbar_wait(&bar[(t - 1) & 1], ((t - 1) >> 1) & 1);   // two tiles in flight
            }
            ptx::mbar_wait(&bar[(tiles - 1) & 1], ((tiles - 1) >> 1) & 1);
            done = 1;
        }
    } else if (do_read) {
        const int q = warp & 3, part = warp >> 2;
        constexpr int COLS = 256 / WPQ;
        const uint32_t a = tbase + ((uint32_t)(q * 32) << 16) + part * COLS;
        for (int t = 0; do_mma ? !done : t < tiles; ++t) {
            nread++;
#pragma unroll 1
            for (int c = 0; c < COLS; c += N) {
                uint32_t r[N];
                ld<N>(a + c, r);
                ptx::tmem_ld_wait();
                acc = fmaxf(acc, tree_max<N>(r));
            }
        }
    }
    const long long t1 = clock64();
    if (lane == 0) { cyc[blockIdx.x * 32 + warp] = t1 - t0; if (warp == 0) cyc[blockIdx.x * 32 + 31] = nread; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tbase, 512);
}

template <int N, int WPQ> void run(const char* name, int do_mma, int do_read, int mma_per_tile) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 32 * 8);
    cudaMemset(cyc, 0, 148 * 32 * 8);
    auto kern = k<N, WPQ>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200);
    const int tiles = 2000;
    kern<<<148, 128 * WPQ + 32, 51200>>>(10, mma_per_tile, do_mma, do_read, out, cyc);
    kern<<<148, 128 * WPQ + 32, 51200>>>(tiles, mma_per_tile, do_mma, do_read, out, cyc);
    cudaError_t err = cudaDeviceSynchronize();
    long long c[32]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%-40s mma/tile %2d: read %8.1f cyc/tile (%6.1f B/clk)   mma %8.1f cyc/tile (%5.1f cyc/MMA)  %s\n", name,
           mma_per_tile, do_read ? (double)c[0] / c[31] : 0.0, do_read ? 131072.0 * c[31] / (double)c[0] : 0.0,
           do_mma ? (double)c[4 * WPQ] / tiles : 0.0, do_mma ? (double)c[4 * WPQ] / tiles / mma_per_tile : 0.0,
           cudaGetErrorString(err));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int mpt : {8, 16}) {
        run<32, 1>("MMA only", 1, 0, mpt);
        run<32, 1>("x32 4 warps, read only", 0, 1, mpt);
        run<32, 1>("x32 4 warps + MMA", 1, 1, mpt);
        run<64, 1>("x64 4 warps + MMA", 1, 1, mpt);
        run<32, 2>("x32 8 warps + MMA", 1, 1, mpt);
        run<64, 2>("x64 8 warps + MMA", 1, 1, mpt);
    }
    return 0;
}
