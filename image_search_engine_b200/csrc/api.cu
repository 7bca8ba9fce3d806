// Context, error plumbing and the host-side sequential-RNG helpers of libise.
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <random>
#include <thread>
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
#if defined(__x86_64__)
    // the same recurrence, eight words at a time: every vector step reads st[i + 1 .. i + 8] before it writes
    // st[i .. i + 7], and the second loop's st[i + M - N] were written 227 words earlier -- no hazard inside a vector
    __attribute__((target("avx2"))) void refill_avx2() {
        const __m256i upper = _mm256_set1_epi32((int)0x80000000u), lower = _mm256_set1_epi32(0x7fffffff);
        const __m256i matrix = _mm256_set1_epi32((int)0x9908b0dfu), one = _mm256_set1_epi32(1), zero = _mm256_setzero_si256();
#define ISE_TWIST8(u, v)                                                                                             \
    _mm256_xor_si256(_mm256_srli_epi32(_mm256_or_si256(_mm256_and_si256((u), upper), _mm256_and_si256((v), lower)), 1), \
                     _mm256_and_si256(_mm256_sub_epi32(zero, _mm256_and_si256((v), one)), matrix))
        int i = 0;
        for (; i + 8 <= N - M; i += 8) {
            const __m256i u = _mm256_loadu_si256((const __m256i*)(st + i)), v = _mm256_loadu_si256((const __m256i*)(st + i + 1));
            const __m256i m = _mm256_loadu_si256((const __m256i*)(st + i + M));
            _mm256_storeu_si256((__m256i*)(st + i), _mm256_xor_si256(m, ISE_TWIST8(u, v)));
        }
        for (; i < N - M; ++i) st[i] = st[i + M] ^ twist(st[i], st[i + 1]);
        for (; i + 8 <= N - 1; i += 8) {
            const __m256i u = _mm256_loadu_si256((const __m256i*)(st + i)), v = _mm256_loadu_si256((const __m256i*)(st + i + 1));
            const __m256i m = _mm256_loadu_si256((const __m256i*)(st + i + M - N));
            _mm256_storeu_si256((__m256i*)(st + i), _mm256_xor_si256(m, ISE_TWIST8(u, v)));
        }
        for (; i < N - 1; ++i) st[i] = st[i + M - N] ^ twist(st[i], st[i + 1]);
        st[N - 1] = st[M - 1] ^ twist(st[N - 1], st[0]);
        const __m256i c1 = _mm256_set1_epi32((int)0x9d2c5680u), c2 = _mm256_set1_epi32((int)0xefc60000u);
        for (i = 0; i + 8 <= N; i += 8) {
            __m256i y = _mm256_loadu_si256((const __m256i*)(st + i));
            y = _mm256_xor_si256(y, _mm256_srli_epi32(y, 11));
            y = _mm256_xor_si256(y, _mm256_and_si256(_mm256_slli_epi32(y, 7), c1));
            y = _mm256_xor_si256(y, _mm256_and_si256(_mm256_slli_epi32(y, 15), c2));
            y = _mm256_xor_si256(y, _mm256_srli_epi32(y, 18));
            _mm256_storeu_si256((__m256i*)(out + i), y);
        }
#undef ISE_TWIST8
        pos = 0;
    }
#endif
    void refill() {
#if defined(__x86_64__)
        static const bool have_avx2 = __builtin_cpu_supports("avx2");
        if (have_avx2) { refill_avx2(); return; }
#endif
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

// ---- sparse index of the split_clusters RNG stream ---------------------------------------------------------------
// split_clusters seeds a FRESH mt19937(1234) on every call, so the stream of draws is the same for every iteration of
// every training run in the process.  A probed donor cj is accepted when r = float(x) / 2^32 < p = (h[cj] - 1) / (n - k);
// with a large codebook p is tiny for every cluster (sum_j p_j = 1), so only "small" draws can ever accept: a draw
// x >= 2^24 gives r >= 2^-8 and is rejected whenever p_max <= 2^-8.  The stream is therefore generated ONCE per
// process and only the positions / values of its small draws are kept (1 in 256: 12 bytes per 256 draws); a plan then
// walks ~k / 256 candidates per split instead of ~k draws, and later iterations reuse the index (k = 65536, 1.7 k
// splits: 111 M draws -> 0.43 M candidate checks).  Same draws, same order, same float predicate as Faiss's loop.
namespace {
struct SplitStreamIndex {
    static constexpr uint32_t kSmall = 1u << 24;
    static constexpr int64_t kStep = 4 << 20;             // draws generated per extension step
    std::mutex mu;
    BlockMT19937 mt{1234};
    int64_t generated = 0;                                // draws consumed from mt so far
    std::vector<int64_t> pos;                             // stream positions of the small draws, ascending
    std::vector<uint32_t> val;
    std::thread warm;
    bool warm_running = false;                            // guarded by mu
    int64_t warm_target = 0;                              // guarded by mu
    std::atomic<bool> stop{false};
    std::atomic<int> waiters{0};                          // plans waiting for mu: the warm-up thread steps aside for them
    // caller holds mu
    void extend_to(int64_t target) {
        while (generated < target) {
            if (mt.pos == BlockMT19937::N) mt.refill();
            const int take = (int)std::min<int64_t>(BlockMT19937::N - mt.pos, target - generated);
            const uint32_t* o = mt.out + mt.pos;
            int i = 0;
            // 1 draw in 256 is small: test 8 at a time on the OR of their top bytes, look closer only on a hit
            for (; i + 8 <= take; i += 8) {
                const uint32_t any = ((o[i] >> 24) == 0) | ((o[i + 1] >> 24) == 0) | ((o[i + 2] >> 24) == 0) | ((o[i + 3] >> 24) == 0) |
                                     ((o[i + 4] >> 24) == 0) | ((o[i + 5] >> 24) == 0) | ((o[i + 6] >> 24) == 0) | ((o[i + 7] >> 24) == 0);
                if (any)
                    for (int j = i; j < i + 8; ++j)
                        if (o[j] < kSmall) { pos.push_back(generated + j); val.push_back(o[j]); }
            }
            for (; i < take; ++i)
                if (o[i] < kSmall) { pos.push_back(generated + i); val.push_back(o[i]); }
            mt.pos += take;
            generated += take;
        }
    }
    ~SplitStreamIndex() {
        stop.store(true);
        if (warm.joinable()) warm.join();
    }
};
SplitStreamIndex g_split_stream;
}  // namespace

// Non-blocking: generate the first n_draws of the stream in a background thread (k-means training calls this when it
// starts on a large codebook, so the index exists by the time an iteration needs a plan).  A later call with a larger
// target raises the running thread's target, or starts a new thread when the previous one has finished.
ISE_EXPORT int ise_split_plan_warm(int64_t n_draws) {
    ISE_CHECK_ARG(n_draws >= 0);
    SplitStreamIndex& ix = g_split_stream;
    std::lock_guard<std::mutex> lk(ix.mu);
    if (ix.generated >= n_draws) return 0;
    ix.warm_target = std::max(ix.warm_target, n_draws);
    if (ix.warm_running) return 0;
    if (ix.warm.joinable()) ix.warm.join();      // finished earlier (it cleared warm_running under mu and needs mu no more)
    ix.warm_running = true;
    ix.warm = std::thread([]() {
        SplitStreamIndex& s = g_split_stream;
        for (;;) {
            while (s.waiters.load() > 0 && !s.stop.load()) std::this_thread::yield();   // std::mutex is not fair
            std::lock_guard<std::mutex> g(s.mu);
            if (s.stop.load() || s.generated >= s.warm_target) {
                s.warm_running = false;
                return;
            }
            s.extend_to(std::min<int64_t>(s.warm_target, s.generated + SplitStreamIndex::kStep));
        }
    });
    return 0;
}

// the original form: one integer compare per draw against per-centroid thresholds (any donor probabilities)
static int split_plan_dense(float* hassign, int64_t k, int64_t n, int32_t* pairs, int32_t* nsplit) {
    BlockMT19937 mt(1234);
    int32_t ns = 0;
    const float denom = (float)(n - k);
    // Same draws, same decisions as Faiss's loop `r = mt() / float(mt.max()); if (r < p) break;`: x -> float(x) / M is
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

ISE_EXPORT int ise_split_plan(float* hassign, int64_t k, int64_t n, int32_t* pairs, int32_t* nsplit) {
    ISE_CHECK_ARG(hassign != nullptr && pairs != nullptr && nsplit != nullptr && k > 0 && n > k);
    const float denom = (float)(n - k);
    const float rng_max = 4294967295.0f;      // float(std::mt19937::max()) == 2^32
    float h_max = 0.f;
    bool any_empty = false;
    for (int64_t c = 0; c < k; c++) {
        h_max = std::max(h_max, hassign[c]);
        any_empty |= hassign[c] == 0;
    }
    if (!any_empty) { *nsplit = 0; return 0; }
    // donor probabilities only shrink while the plan runs (a split halves the donor), so p_max is known up front
    const float p_max = (float)((h_max - 1.0) / denom);
    if (!(p_max <= 0.00390625f) || getenv("ISE_SPLIT_PLAN_DENSE"))     // 2^-8: a draw >= 2^24 could accept
        return split_plan_dense(hassign, k, n, pairs, nsplit);
    if (!(h_max > 1.f)) ISE_FAIL("no cluster has more than one point: split_clusters would never terminate");
    SplitStreamIndex& ix = g_split_stream;
    ix.waiters.fetch_add(1);
    std::lock_guard<std::mutex> lk(ix.mu);
    ix.waiters.fetch_sub(1);
    int32_t ns = 0;
    int64_t g = 0;                 // stream position of the next draw
    size_t cur = 0;                // first index entry with pos >= g
    for (int64_t ci = 0; ci < k; ci++) {
        if (hassign[ci] != 0) continue;
        int64_t cj = -1;
        while (cj < 0) {
            if (cur == ix.pos.size()) {             // out of candidates: generate more of the stream
                ix.extend_to(ix.generated + SplitStreamIndex::kStep);
                continue;
            }
            const int64_t P = ix.pos[cur];
            const uint32_t x = ix.val[cur];
            ++cur;
            const int64_t c = (P - g) % k;          // the donor this draw is compared with
            const float p = (float)((hassign[c] - 1.0) / denom);
            if ((float)x / rng_max < p) {           // Faiss: r = mt() / float(mt.max()); if (r < p) break;
                cj = c;
                g = P + 1;
            }
        }
        pairs[2 * ns] = (int32_t)ci;
        pairs[2 * ns + 1] = (int32_t)cj;
        hassign[ci] = hassign[cj] / 2;
        hassign[cj] -= hassign[ci];
        ns++;
    }
    *nsplit = ns;
    return 0;
}

// ---- ragged ingestion: list of per-image descriptor arrays -> one packed (pinned) matrix ------------------------------
// The reference's input contract is a Python list with one (n_i, d) array per image (descriptors.py:104-139); its
// run_clustering does one single-threaded np.concatenate (bag_of_visual_words.py:128).  This is the same copy,
// spread over host threads, straight into the caller's pinned staging buffer, optionally narrowing float32 -> uint8
// on the way when every value is an integer in [0, 255] (OpenCV SIFT / ORB-as-float descriptors are): a quarter of
// the host -> device bytes, and the device path widens uint8 for free.
namespace {
// returns false as soon as a value is not exactly representable as uint8
inline bool narrow_f32_to_u8_scalar(const float* src, uint8_t* dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        const float v = src[i];
        if (!(v >= 0.f && v <= 255.f)) return false;
        const int q = (int)v;
        if ((float)q != v) return false;
        dst[i] = (uint8_t)q;
    }
    return true;
}

#if defined(__x86_64__)
// 32 floats per step: truncate, compare back (exact integers only), range-check via the packed saturations
__attribute__((target("avx2"))) bool narrow_f32_to_u8_avx2(const float* src, uint8_t* dst, int64_t n) {
    int64_t i = 0;
    const __m256i perm = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    for (; i + 32 <= n; i += 32) {
        const __m256 a = _mm256_loadu_ps(src + i), b = _mm256_loadu_ps(src + i + 8);
        const __m256 c = _mm256_loadu_ps(src + i + 16), e = _mm256_loadu_ps(src + i + 24);
        const __m256i qa = _mm256_cvttps_epi32(a), qb = _mm256_cvttps_epi32(b);
        const __m256i qc = _mm256_cvttps_epi32(c), qe = _mm256_cvttps_epi32(e);
        // exact integer <=> converting back gives the same float (NaN / out-of-range values fail the compare)
        __m256 ok = _mm256_and_ps(_mm256_and_ps(_mm256_cmp_ps(_mm256_cvtepi32_ps(qa), a, _CMP_EQ_OQ),
                                                _mm256_cmp_ps(_mm256_cvtepi32_ps(qb), b, _CMP_EQ_OQ)),
                                  _mm256_and_ps(_mm256_cmp_ps(_mm256_cvtepi32_ps(qc), c, _CMP_EQ_OQ),
                                                _mm256_cmp_ps(_mm256_cvtepi32_ps(qe), e, _CMP_EQ_OQ)));
        // [0, 255] <=> no bit above the low byte
        const __m256i hi_bits = _mm256_or_si256(_mm256_or_si256(qa, qb), _mm256_or_si256(qc, qe));
        const __m256i in_range = _mm256_cmpeq_epi32(_mm256_srli_epi32(hi_bits, 8), _mm256_setzero_si256());
        if (_mm256_movemask_ps(_mm256_and_ps(ok, _mm256_castsi256_ps(in_range))) != 0xFF) return false;
        const __m256i w0 = _mm256_packus_epi32(qa, qb), w1 = _mm256_packus_epi32(qc, qe);   // per 128-bit lane
        const __m256i bytes = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(w0, w1), perm);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), bytes);
    }
    return narrow_f32_to_u8_scalar(src + i, dst + i, n - i);
}
#endif

inline bool narrow_f32_to_u8(const float* src, uint8_t* dst, int64_t n) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return narrow_f32_to_u8_avx2(src, dst, n);
#endif
    return narrow_f32_to_u8_scalar(src, dst, n);
}
}  // namespace

namespace {
// Thread t's share (of nt) of images [i0, i1): contiguous image ranges balanced by row count.
void pack_share(const void* const* srcs, const int64_t* offsets, int64_t i0, int64_t i1, int d, int src_dtype,
                int dst_dtype, void* dst_base, int t, int nt, std::atomic<int>& ok) {
    const size_t src_elt = src_dtype == ISE_DTYPE_F32 ? 4 : 1, dst_elt = dst_dtype == ISE_DTYPE_F32 ? 4 : 1;
    const int64_t total_rows = offsets[i1] - offsets[i0];
    const int64_t r_lo = offsets[i0] + total_rows * t / nt, r_hi = offsets[i0] + total_rows * (t + 1) / nt;
    int64_t a = std::lower_bound(offsets + i0, offsets + i1, r_lo) - offsets;
    int64_t b = std::lower_bound(offsets + i0, offsets + i1, r_hi) - offsets;
    if (t == nt - 1) b = i1;
    for (int64_t i = a; i < b && ok.load(std::memory_order_relaxed); ++i) {
        const int64_t rows = offsets[i + 1] - offsets[i];
        if (rows <= 0) continue;
        uint8_t* dst = (uint8_t*)dst_base + (size_t)offsets[i] * d * dst_elt;
        if (src_dtype == dst_dtype) memcpy(dst, srcs[i], (size_t)rows * d * src_elt);
        else if (!narrow_f32_to_u8((const float*)srcs[i], dst, rows * d)) ok.store(0);
    }
}

// One unit (image / row block) into its place of the packed matrix.
inline void pack_unit(const void* const* srcs, const int64_t* offsets, int64_t i, int d, int src_dtype, int dst_dtype,
                      void* dst_base, std::atomic<int>& ok) {
    const int64_t rows = offsets[i + 1] - offsets[i];
    if (rows <= 0 || !ok.load(std::memory_order_relaxed)) return;
    const size_t src_elt = src_dtype == ISE_DTYPE_F32 ? 4 : 1, dst_elt = dst_dtype == ISE_DTYPE_F32 ? 4 : 1;
    uint8_t* dst = (uint8_t*)dst_base + (size_t)offsets[i] * d * dst_elt;
    if (src_dtype == dst_dtype) memcpy(dst, srcs[i], (size_t)rows * d * src_elt);
    else if (!narrow_f32_to_u8((const float*)srcs[i], dst, rows * d)) ok.store(0);
}

// One asynchronous packing job: nt threads walk the chunks IN ORDER and take the units of a chunk from a shared counter
// (a thread that starts late or shares its core does not hold a chunk back), so chunk 0 is complete after ~1/n_chunks of
// the work and the caller can send it while the rest is still being packed.
struct PackJob {
    std::vector<std::thread> threads;
    std::vector<int64_t> cuts;
    std::unique_ptr<std::atomic<int64_t>[]> next;     // per chunk: next unit to hand out (relative)
    std::unique_ptr<std::atomic<int64_t>[]> done;     // per chunk: units completed
    std::unique_ptr<std::atomic<int>[]> owner;        // per chunk: 0 = nobody yet, 1 = the workers, 2 = the caller took it
    std::atomic<int> ok{1};
    std::mutex mu;
    std::condition_variable cv;
    int nt = 1;
    bool finished(int c) const {
        return owner[c].load() == 2 || done[c].load() >= cuts[c + 1] - cuts[c];
    }
};
}  // namespace

ISE_EXPORT int ise_pack_begin(const void* const* srcs, const int64_t* offsets, const int64_t* image_cuts, int n_chunks,
                              int d, int src_dtype, int dst_dtype, void* dst_base, int nthreads, void** job_out) {
    ISE_CHECK_ARG(srcs && offsets && image_cuts && dst_base && job_out && d > 0 && n_chunks > 0);
    ISE_CHECK_ARG(src_dtype == ISE_DTYPE_F32 || src_dtype == ISE_DTYPE_U8);
    ISE_CHECK_ARG(dst_dtype == ISE_DTYPE_F32 || dst_dtype == ISE_DTYPE_U8);
    ISE_CHECK_ARG(!(src_dtype == ISE_DTYPE_U8 && dst_dtype == ISE_DTYPE_F32));
    for (int c = 0; c < n_chunks; ++c) ISE_CHECK_ARG(image_cuts[c] >= 0 && image_cuts[c + 1] >= image_cuts[c]);
    PackJob* job = new PackJob();
    job->cuts.assign(image_cuts, image_cuts + n_chunks + 1);
    job->next.reset(new std::atomic<int64_t>[n_chunks]);
    job->done.reset(new std::atomic<int64_t>[n_chunks]);
    job->owner.reset(new std::atomic<int>[n_chunks]);
    for (int c = 0; c < n_chunks; ++c) { job->next[c].store(0); job->done[c].store(0); job->owner[c].store(0); }
    job->nt = std::max(1, std::min<int>(nthreads, 64));
    const int nt = job->nt;
    for (int t = 0; t < nt; ++t)
        job->threads.emplace_back([=]() {
            for (int c = 0; c < n_chunks; ++c) {
                int who = 0;
                job->owner[c].compare_exchange_strong(who, 1);      // the first worker to arrive takes the chunk ...
                if (job->owner[c].load() != 1) continue;            // ... unless the caller claimed it (ise_pack_claim)
                const int64_t u0 = job->cuts[c], n_units = job->cuts[c + 1] - u0;
                for (;;) {
                    const int64_t u = job->next[c].fetch_add(1);
                    if (u >= n_units) break;
                    pack_unit(srcs, offsets, u0 + u, d, src_dtype, dst_dtype, dst_base, job->ok);
                    if (job->done[c].fetch_add(1) + 1 == n_units) {
                        std::lock_guard<std::mutex> lk(job->mu);
                        job->cv.notify_all();
                    }
                }
            }
        });
    *job_out = job;
    return 0;
}

ISE_EXPORT int ise_pack_wait(void* job_, int chunk, int* ok_out) {
    PackJob* job = (PackJob*)job_;
    ISE_CHECK_ARG(job && ok_out && chunk >= 0 && chunk + 1 < (int)job->cuts.size());
    std::unique_lock<std::mutex> lk(job->mu);
    // (re-checked every 200 us: a chunk without units completes without a notification)
    while (!job->finished(chunk)) job->cv.wait_for(lk, std::chrono::microseconds(200));
    *ok_out = job->ok.load();
    return 0;
}

ISE_EXPORT int ise_pack_poll(void* job_, int chunk, int* done_out, int* ok_out) {
    PackJob* job = (PackJob*)job_;
    ISE_CHECK_ARG(job && done_out && ok_out && chunk >= 0 && chunk + 1 < (int)job->cuts.size());
    const bool done = job->owner[chunk].load() != 2 && job->done[chunk].load() >= job->cuts[chunk + 1] - job->cuts[chunk];
    *ok_out = job->ok.load();      // read AFTER the count: ok == 1 now means no unit of the chunk was skipped
    *done_out = done ? 1 : 0;
    return 0;
}

ISE_EXPORT int ise_pack_claim(void* job_, int chunk, int* claimed_out) {
    PackJob* job = (PackJob*)job_;
    ISE_CHECK_ARG(job && claimed_out && chunk >= 0 && chunk + 1 < (int)job->cuts.size());
    int who = 0;
    *claimed_out = job->owner[chunk].compare_exchange_strong(who, 2) ? 1 : 0;
    return 0;
}

ISE_EXPORT int ise_pack_end(void* job_) {
    PackJob* job = (PackJob*)job_;
    if (!job) return 0;
    for (auto& t : job->threads) t.join();
    delete job;
    return 0;
}

ISE_EXPORT int ise_pack_rows(const void* const* srcs, const int64_t* offsets, int64_t i0, int64_t i1, int d,
                             int src_dtype, int dst_dtype, void* dst_base, int nthreads, int* ok_out) {
    ISE_CHECK_ARG(srcs && offsets && dst_base && ok_out && d > 0 && i0 >= 0 && i1 >= i0);
    ISE_CHECK_ARG(src_dtype == ISE_DTYPE_F32 || src_dtype == ISE_DTYPE_U8);
    ISE_CHECK_ARG(dst_dtype == ISE_DTYPE_F32 || dst_dtype == ISE_DTYPE_U8);
    ISE_CHECK_ARG(!(src_dtype == ISE_DTYPE_U8 && dst_dtype == ISE_DTYPE_F32));     // widening happens on the device
    *ok_out = 1;
    if (i1 == i0) return 0;
    const int64_t total_rows = offsets[i1] - offsets[i0];
    int nt = std::max(1, std::min<int>(nthreads, 64));
    if (total_rows * d < (int64_t)1 << 16) nt = 1;
    std::atomic<int> ok{1};
    auto work = [&](int t) { pack_share(srcs, offsets, i0, i1, d, src_dtype, dst_dtype, dst_base, t, nt, ok); };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        th.reserve(nt);
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    *ok_out = ok.load();
    return 0;
}
