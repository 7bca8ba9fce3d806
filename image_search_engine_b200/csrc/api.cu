// Context, error plumbing and the host-side sequential-RNG helpers of libise.
#include <random>
#include <unordered_map>
#include <vector>

#include "common.cuh"

static thread_local std::string g_last_error;

void ise_set_error(const std::string& msg) { g_last_error = msg; }

ISE_EXPORT int ise_version(void) { return ISE_VERSION; }

ISE_EXPORT const char* ise_last_error(void) { return g_last_error.c_str(); }

ISE_EXPORT int ise_ctx_create(int device, ise_ctx** out) {
    ISE_CHECK_ARG(out != nullptr);
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        ISE_FAIL(std::string("no CUDA device visible (there is no CPU fallback): ") + cudaGetErrorString(e));
    ISE_CHECK_ARG(device >= 0 && device < ndev);
    cudaDeviceProp prop;
    ISE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        ISE_FAIL("libise is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) +
                 std::to_string(prop.minor));
    DeviceGuard g(device);
    ise_ctx* ctx = new ise_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    ctx->encode_tiled = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ctx->encode_tiled, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ctx->encode_tiled == nullptr) {
        delete ctx;
        ISE_FAIL("cuTensorMapEncodeTiled not available from the driver");
    }
    ctx->sync_buf = nullptr;
    if (cudaMalloc(&ctx->sync_buf, kSyncInts * sizeof(int32_t)) != cudaSuccess) ctx->sync_buf = nullptr;   // optional
    *out = ctx;
    return 0;
}

ISE_EXPORT void ise_ctx_destroy(ise_ctx* ctx) {
    if (!ctx) return;
    if (ctx->sync_buf) {
        DeviceGuard g(ctx->device);
        cudaFree(ctx->sync_buf);
    }
    delete ctx;
}

ISE_EXPORT int ise_ctx_sm_count(const ise_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

// faiss/utils/random.cpp rand_perm: perm = iota; for i in [0, n-1): swap(perm[i], perm[i + mt() % (n - i)]).
// Entry i is final after step i, so the first m entries need m steps; displaced slots are kept sparse.
ISE_EXPORT int ise_rand_perm_prefix(int64_t n, int64_t seed, int64_t m, int64_t* out) {
    ISE_CHECK_ARG(n >= 0 && m >= 0 && m <= n && (out != nullptr || m == 0));
    std::mt19937 mt((unsigned int)seed);
    int64_t steps = std::min<int64_t>(m, std::max<int64_t>(n - 1, 0));
    if (m * 4 >= n) {
        std::vector<int64_t> perm((size_t)n);
        for (int64_t i = 0; i < n; i++) perm[(size_t)i] = i;
        for (int64_t i = 0; i < steps; i++) {
            int64_t j = i + (int64_t)(mt() % (uint64_t)(n - i));
            std::swap(perm[(size_t)i], perm[(size_t)j]);
        }
        for (int64_t i = 0; i < m; i++) out[i] = perm[(size_t)i];
        return 0;
    }
    std::unordered_map<int64_t, int64_t> moved;
    moved.reserve((size_t)steps * 2);
    auto get = [&](int64_t p) {
        auto it = moved.find(p);
        return it == moved.end() ? p : it->second;
    };
    for (int64_t i = 0; i < steps; i++) {
        int64_t j = i + (int64_t)(mt() % (uint64_t)(n - i));
        int64_t vi = get(i), vj = get(j);
        moved[j] = vi;
        out[i] = vj;
    }
    for (int64_t i = steps; i < m; i++) out[i] = get(i);
    return 0;
}

// faiss/Clustering.cpp split_clusters(): the RNG stream is consumed one draw per probed donor, so the
// plan is inherently sequential; it only runs when an iteration produced empty clusters.
namespace {
// std::mt19937-compatible generator whose tempered outputs are produced 624 at a time into a flat buffer, so that
// the donor scan below is a tight compare loop over an array instead of a call (with its refill check) per draw.
struct BlockMT19937 {
    static constexpr int N = 624, M = 397;
    uint32_t st[N];
    uint32_t out[N];
    int pos = N;
    explicit BlockMT19937(uint32_t seed) {
        st[0] = seed;
        for (int i = 1; i < N; ++i) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
    }
    static uint32_t twist(uint32_t u, uint32_t v) {
        const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
        return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
    }
    void refill() {
        for (int i = 0; i < N - M; ++i) st[i] = st[i + M] ^ twist(st[i], st[i + 1]);
        for (int i = N - M; i < N - 1; ++i) st[i] = st[i + M - N] ^ twist(st[i], st[i + 1]);
        st[N - 1] = st[M - 1] ^ twist(st[N - 1], st[0]);
        for (int i = 0; i < N; ++i) {
            uint32_t y = st[i];
            y ^= y >> 11;
            y ^= (y << 7) & 0x9d2c5680u;
            y ^= (y << 15) & 0xefc60000u;
            y ^= y >> 18;
            out[i] = y;
        }
        pos = 0;
    }
};
}  // namespace

ISE_EXPORT int ise_split_plan(float* hassign, int64_t k, int64_t n, int32_t* pairs, int32_t* nsplit) {
    ISE_CHECK_ARG(hassign != nullptr && pairs != nullptr && nsplit != nullptr && k > 0 && n > k);
    BlockMT19937 mt(1234);
    int32_t ns = 0;
    const float denom = (float)(n - k);
    // Same draws, same decisions as Faiss's loop `r = mt() / float(mt.max()); if (r < p) break;`.  At k = 65536 a
    // split probes ~n / (mean cluster size) = 66 k donors and an iteration that empties ~1.7 k clusters spends its
    // time here, on the host, so the test is reduced to one integer compare per draw: x -> float(x) / M is
    // monotone, hence `r < p`  <=>  x < T(p) with T(p) the smallest draw whose quotient is not below p (found
    // by bisection, once per centroid and again for the two entries a split changes).
    const float rng_max = 4294967295.0f;      // float(std::mt19937::max())
    auto threshold = [&](float h) -> uint64_t {
        const float p = (float)((h - 1.0) / denom);
        uint64_t lo = 0, hi = (uint64_t)1 << 32;            // smallest x in [0, 2^32] with !(x / M < p)
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if ((float)(uint32_t)mid / rng_max < p) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    std::vector<uint64_t> thr((size_t)k);
    for (int64_t c = 0; c < k; c++) thr[(size_t)c] = threshold(hassign[c]);
    for (int64_t ci = 0; ci < k; ci++) {
        if (hassign[ci] != 0) continue;
        int64_t cj = 0;
        for (bool found = false; !found;) {
            if (mt.pos == BlockMT19937::N) mt.refill();
            while (mt.pos < BlockMT19937::N) {
                if ((uint64_t)mt.out[mt.pos++] < thr[(size_t)cj]) { found = true; break; }
                if (++cj == k) cj = 0;
            }
        }
        pairs[2 * ns] = (int32_t)ci;
        pairs[2 * ns + 1] = (int32_t)cj;
        hassign[ci] = hassign[cj] / 2;
        hassign[cj] -= hassign[ci];
        thr[(size_t)ci] = threshold(hassign[ci]);
        thr[(size_t)cj] = threshold(hassign[cj]);
        ns++;
    }
    *nsplit = ns;
    return 0;
}
