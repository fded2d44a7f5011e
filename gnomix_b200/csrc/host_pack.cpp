// host_pack.cpp -- host side of the packed H2D path of gnx_infer_host.
//
// The haplotype matrix the reference hands to Base.predict_proba is int8 with values
// {0,1,2} (src/utils.py:150-153), i.e. 2 bits of information per byte.  PCIe moves
// ~55 GB/s, the host cores read memory several times faster, so the host-buffer
// pipeline packs every 64 SNPs of a row into two 64-bit planes (bit 0, bit 1 of the
// value) while the previous chunk is in flight and ships a quarter of the bytes; a
// device kernel (pack.cu) restores the int8 tile K1's TMA loads expect.  Rows holding
// any value outside 0..3 are reported so the caller can ship that chunk unpacked.
//
// Packed row layout: group g (SNPs 64g .. 64g+63) = { u64 plane0, u64 plane1 }, bit i
// of planeK = bit K of X[row][64g + i]; SNPs >= C pack as 0.
#include <immintrin.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "host_pack.h"

namespace gnx {

// ------------------------------------------------------------------ row kernels
static inline void pack_tail(const int8_t* x, int64_t n, uint64_t* out, unsigned* bad) {
    uint64_t p0 = 0, p1 = 0;
    unsigned b = 0;
    for (int64_t i = 0; i < n; i++) {
        const unsigned v = (uint8_t)x[i];
        b |= v;
        p0 |= (uint64_t)(v & 1u) << i;
        p1 |= (uint64_t)((v >> 1) & 1u) << i;
    }
    out[0] = p0;
    out[1] = p1;
    *bad |= b;
}

static unsigned pack_row_scalar(const int8_t* x, int64_t C, uint64_t* out, int64_t groups) {
    unsigned bad = 0;
    const int64_t full = C / 64;
    for (int64_t g = 0; g < full; g++) pack_tail(x + 64 * g, 64, out + 2 * g, &bad);
    int64_t g = full;
    if (g < groups && C > 64 * full) {
        pack_tail(x + 64 * full, C - 64 * full, out + 2 * g, &bad);
        g++;
    }
    for (; g < groups; g++) out[2 * g] = out[2 * g + 1] = 0;
    return bad & 0xFCu;
}

__attribute__((target("avx2"))) static unsigned pack_row_avx2(const int8_t* x, int64_t C, uint64_t* out, int64_t groups) {
    const int64_t full = C / 64;
    __m256i acc = _mm256_setzero_si256();
    for (int64_t g = 0; g < full; g++) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(x + 64 * g));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(x + 64 * g + 32));
        acc = _mm256_or_si256(acc, _mm256_or_si256(a, b));
        const uint32_t a0 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(a, 7));
        const uint32_t b0 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(b, 7));
        const uint32_t a1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(a, 6));
        const uint32_t b1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(b, 6));
        out[2 * g] = (uint64_t)a0 | ((uint64_t)b0 << 32);
        out[2 * g + 1] = (uint64_t)a1 | ((uint64_t)b1 << 32);
    }
    const __m256i hi = _mm256_and_si256(acc, _mm256_set1_epi8((char)0xFC));
    unsigned bad = _mm256_testz_si256(hi, hi) ? 0u : 0xFCu;
    int64_t g = full;
    if (g < groups && C > 64 * full) {
        pack_tail(x + 64 * full, C - 64 * full, out + 2 * g, &bad);
        g++;
    }
    for (; g < groups; g++) out[2 * g] = out[2 * g + 1] = 0;
    return bad & 0xFCu;
}

__attribute__((target("avx512f,avx512bw"))) static unsigned pack_row_avx512(const int8_t* x, int64_t C, uint64_t* out,
                                                                           int64_t groups) {
    const int64_t full = C / 64;
    const __m512i one = _mm512_set1_epi8(1), two = _mm512_set1_epi8(2);
    __m512i acc = _mm512_setzero_si512();
    // (measured on the B200 box's 16 host threads: 118 GB/s; streaming stores changed nothing, software
    // prefetch cost 10 %)
    int64_t g = 0;
    for (; g + 2 <= full; g += 2) {
        const __m512i a = _mm512_loadu_si512(x + 64 * g);
        const __m512i b = _mm512_loadu_si512(x + 64 * g + 64);
        acc = _mm512_or_si512(acc, _mm512_or_si512(a, b));
        out[2 * g] = _mm512_test_epi8_mask(a, one);
        out[2 * g + 1] = _mm512_test_epi8_mask(a, two);
        out[2 * g + 2] = _mm512_test_epi8_mask(b, one);
        out[2 * g + 3] = _mm512_test_epi8_mask(b, two);
    }
    for (; g < full; g++) {
        const __m512i a = _mm512_loadu_si512(x + 64 * g);
        acc = _mm512_or_si512(acc, a);
        out[2 * g] = _mm512_test_epi8_mask(a, one);
        out[2 * g + 1] = _mm512_test_epi8_mask(a, two);
    }
    unsigned bad = _mm512_test_epi8_mask(acc, _mm512_set1_epi8((char)0xFC)) ? 0xFCu : 0u;
    if (g < groups && C > 64 * full) {
        const int64_t n = C - 64 * full;
        const __mmask64 k = (~0ULL) >> (64 - n);
        const __m512i a = _mm512_maskz_loadu_epi8(k, x + 64 * full);
        if (_mm512_test_epi8_mask(a, _mm512_set1_epi8((char)0xFC))) bad = 0xFCu;
        out[2 * g] = _mm512_test_epi8_mask(a, one);
        out[2 * g + 1] = _mm512_test_epi8_mask(a, two);
        g++;
    }
    for (; g < groups; g++) out[2 * g] = out[2 * g + 1] = 0;
    return bad;
}

typedef unsigned (*pack_row_fn)(const int8_t*, int64_t, uint64_t*, int64_t);

static pack_row_fn choose_impl(int* which) {
    const char* force = getenv("GNX_HOST_PACK_ISA");  // "scalar" | "avx2" | "avx512" (tests)
    __builtin_cpu_init();
    const bool has512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
    const bool has2 = __builtin_cpu_supports("avx2");
    if (force && !strcmp(force, "scalar")) { *which = 0; return pack_row_scalar; }
    if (force && !strcmp(force, "avx2") && has2) { *which = 1; return pack_row_avx2; }
    if (has512) { *which = 2; return pack_row_avx512; }
    if (has2) { *which = 1; return pack_row_avx2; }
    *which = 0;
    return pack_row_scalar;
}

unsigned pack_row_best(const int8_t* x, int64_t C, uint64_t* out, int64_t groups) {
    static int which = 0;
    static const pack_row_fn fn = choose_impl(&which);
    return fn(x, C, out, groups);
}

// ------------------------------------------------------------------ worker pool
// Persistent threads; run(n_items, fn) hands out item indices from an atomic counter and
// returns when all are done.  One job at a time (gnx_infer_host holds the workspace lock).
class Pool {
  public:
    explicit Pool(int n) : n_(n) {
        for (int i = 0; i < n_; i++) th_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return n_; }
    void run(int64_t items, const std::function<void(int64_t)>& fn) {
        {
            std::lock_guard<std::mutex> l(mu_);
            fn_ = &fn;
            items_ = items;
            next_.store(0, std::memory_order_relaxed);
            pending_ = n_;
            gen_++;
        }
        cv_.notify_all();
        work();  // the caller works too
        std::unique_lock<std::mutex> l(mu_);
        done_.wait(l, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void work() {
        for (;;) {
            const int64_t i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= items_) break;
            (*fn_)(i);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            work();
            {
                std::lock_guard<std::mutex> l(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int64_t)>* fn_ = nullptr;
    int64_t items_ = 0;
    std::atomic<int64_t> next_{0};
    int pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

int host_threads_default() {
    int n = 0;
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = CPU_COUNT(&set);
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    if (n <= 0) n = 1;
    if (const char* e = getenv("GNX_HOST_THREADS")) {
        const int v = atoi(e);
        if (v > 0) return std::min(v, 128);
    }
    // one process per GPU (torchrun): the ranks of a node share its cores
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) {
        const int v = atoi(e);
        if (v > 1) n = std::max(1, n / v);
    }
    return std::min(n, 128);
}

static Pool* g_pool = nullptr;
static pid_t g_pool_pid = 0;
static std::mutex g_pool_mu;

static Pool* pool(int threads) {
    // workers besides the calling thread
    const int want = std::max(0, threads - 1);
    if (g_pool && g_pool_pid != getpid()) g_pool = nullptr;  // forked child: the parent's workers do not exist here
    if (!g_pool || g_pool->size() != want) {
        delete g_pool;
        g_pool = new Pool(want);
        g_pool_pid = getpid();
    }
    return g_pool;
}

void parallel_for(int64_t items, int threads, const std::function<void(int64_t)>& fn) {
    if (threads <= 0) threads = host_threads_default();
    if (threads == 1 || items <= 1) {
        for (int64_t i = 0; i < items; i++) fn(i);
        return;
    }
    std::lock_guard<std::mutex> l(g_pool_mu);
    pool(threads)->run(items, fn);
}

int pack_rows(const int8_t* X, int64_t n, int64_t ldX, int64_t C, uint64_t* out, int64_t out_pitch_words, int threads,
              int* isa) {
    int which = 0;
    const pack_row_fn fn = choose_impl(&which);
    if (isa) *isa = which;
    const int64_t groups = out_pitch_words / 2;
    if (threads <= 0) threads = host_threads_default();
    std::atomic<unsigned> bad{0};
    // blocks of rows sized to ~1 MB of input so that the atomic counter stays cold
    const int64_t rows_per = std::max<int64_t>(1, (int64_t(1) << 20) / std::max<int64_t>(C, 1));
    const int64_t blocks = (n + rows_per - 1) / rows_per;
    const std::function<void(int64_t)> job = [&](int64_t b) {
        unsigned local = 0;
        const int64_t r1 = std::min(n, (b + 1) * rows_per);
        for (int64_t r = b * rows_per; r < r1; r++) local |= fn(X + r * ldX, C, out + r * out_pitch_words, groups);
        if (local) bad.fetch_or(local, std::memory_order_relaxed);
    };
    parallel_for(blocks, threads, job);
    return bad.load() ? 1 : 0;
}

}  // namespace gnx

extern "C" int gnx_pack_rows_host(const int8_t* X, int64_t n, int64_t ldX, int64_t C, uint64_t* out, int64_t out_pitch_words,
                                  int threads, int* out_of_range) {
    if (n < 0 || C < 0 || ldX < C || out_pitch_words < 2 * ((C + 63) / 64) || (n > 0 && (!X || !out))) return 2;
    const int bad = gnx::pack_rows(X, n, ldX, C, out, out_pitch_words, threads, nullptr);
    if (out_of_range) *out_of_range = bad;
    return 0;
}

extern "C" int gnx_host_threads(void) { return gnx::host_threads_default(); }
