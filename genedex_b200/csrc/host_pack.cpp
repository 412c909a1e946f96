// host_pack.cpp -- see host_pack.h
#include "host_pack.h"

#include <immintrin.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <exception>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <thread>

namespace gdx {

// ---- thread pool --------------------------------------------------------------------------------------
namespace {
struct Job {
    const std::function<void(uint64_t)> *fn;
    uint64_t pieces;
    std::atomic<uint64_t> next{0};
    std::atomic<uint64_t> done{0};
    std::atomic<bool> failed{false};
    std::exception_ptr error;  // the first exception a piece threw (rethrown on the thread that posted the job)
};
}  // namespace

struct HostPool::Impl {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv, done_cv;
    std::deque<std::shared_ptr<Job>> queue;  // jobs that may still have unclaimed pieces
    bool stop = false;
    std::shared_mutex life;  // shared: a job is posted and waited for; exclusive: the workers are being replaced

    // claims and runs pieces of `job` until none is left
    void work(const std::shared_ptr<Job> &job) {
        for (;;) {
            const uint64_t k = job->next.fetch_add(1, std::memory_order_relaxed);
            if (k >= job->pieces) break;
            try {
                (*job->fn)(k);
            } catch (...) {
                if (!job->failed.exchange(true)) job->error = std::current_exception();
            }
            if (job->done.fetch_add(1, std::memory_order_acq_rel) + 1 == job->pieces) {
                std::lock_guard<std::mutex> lk(mu);
                done_cv.notify_all();
            }
        }
    }
    void loop() {
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] {
                    while (!queue.empty() && queue.front()->next.load(std::memory_order_relaxed) >= queue.front()->pieces)
                        queue.pop_front();  // fully claimed: nothing left to hand out
                    return stop || !queue.empty();
                });
                if (stop) return;
                job = queue.front();
            }
            work(job);
        }
    }
};

static unsigned usable_cpus() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0) {
        const int c = CPU_COUNT(&set);
        if (c > 0) return (unsigned)c;
    }
    const unsigned h = std::thread::hardware_concurrency();
    return h ? h : 1;
}

HostPool::HostPool() : impl_(new Impl()) { resize(0); }

// total = threads that work on one staging job incl. the caller; 0 = GDX_HOST_THREADS, else the CPUs the
// calling thread may run on (new workers inherit its affinity), at most 32.  Waits for running jobs.
unsigned HostPool::resize(unsigned total) {
    std::unique_lock<std::shared_mutex> life(impl_->life);
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop = true;
    }
    impl_->cv.notify_all();
    for (auto &t : impl_->workers) t.join();
    impl_->workers.clear();
    impl_->stop = false;
    if (total == 0) {
        // one CPU is left to the rest of the process (driver threads, the host language's own threads): with every
        // CPU taken, whichever pool thread gets preempted holds up its piece -- and the whole job -- for a scheduler
        // time slice (measured: p90 of a 128 MiB packing job 21 ms instead of 1.9 ms, tools/host_pack_bench.py latency)
        const unsigned cpus = usable_cpus();
        total = std::min(32u, cpus > 2 ? cpus - 1 : cpus);
        if (const char *e = getenv("GDX_HOST_THREADS"))
            if (atoi(e) > 0) total = (unsigned)atoi(e);
    }
    for (unsigned i = 1; i < total; ++i) impl_->workers.emplace_back([this] { impl_->loop(); });
    return threads();
}

HostPool::~HostPool() {
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop = true;
    }
    impl_->cv.notify_all();
    for (auto &t : impl_->workers) t.join();
    delete impl_;
}

thread_local bool HostPool::t_caller_helps = true;

HostPool &HostPool::get() {
    static HostPool pool;
    return pool;
}

unsigned HostPool::threads() const { return (unsigned)impl_->workers.size() + 1; }

void HostPool::parallel_for(uint64_t pieces, const std::function<void(uint64_t)> &fn) {
    if (pieces == 0) return;
    if (pieces == 1) {
        fn(0);
        return;
    }
    std::shared_lock<std::shared_mutex> life(impl_->life);
    if (impl_->workers.empty() || (!t_caller_helps && pieces <= 2)) {
        for (uint64_t k = 0; k < pieces; ++k) fn(k);
        return;
    }
    auto job = std::make_shared<Job>();
    job->fn = &fn;
    job->pieces = pieces;
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->queue.push_back(job);
    }
    impl_->cv.notify_all();
    if (t_caller_helps) impl_->work(job);
    {
        std::unique_lock<std::mutex> lk(impl_->mu);
        impl_->done_cv.wait(lk, [&] { return job->done.load(std::memory_order_acquire) == job->pieces; });
    }
    if (job->failed.load(std::memory_order_acquire)) std::rethrow_exception(job->error);
}

void HostPool::copy(void *dst, const void *src, uint64_t bytes) {
    constexpr uint64_t kPiece = 1ull << 20;
    if (bytes <= 2 * kPiece) {
        memcpy(dst, src, bytes);
        return;
    }
    const uint64_t pieces = (bytes + kPiece - 1) / kPiece;
    parallel_for(pieces, [&](uint64_t k) {
        const uint64_t off = k * kPiece, nb = std::min(kPiece, bytes - off);
        memcpy((uint8_t *)dst + off, (const uint8_t *)src + off, nb);
    });
}

void HostPool::widen_u32(uint64_t *dst, const uint32_t *src, uint64_t n) {
    constexpr uint64_t kPiece = 1ull << 18;  // elements
    const uint64_t pieces = (n + kPiece - 1) / kPiece;
    parallel_for(pieces, [&](uint64_t k) {
        const uint64_t b = k * kPiece, e = std::min(n, b + kPiece);
        for (uint64_t i = b; i < e; ++i) dst[i] = src[i];
    });
}

// ---- 2-bit packer ---------------------------------------------------------------------------------------
void build_pack_table(const uint8_t io_to_dense[256], uint32_t num_searchable, PackTable &t) {
    memset(&t, 0, sizeof t);
    t.usable = num_searchable >= 1 && num_searchable <= 4;
    for (int x = 0; x < 256; ++x) {
        const uint32_t d = io_to_dense[x];
        t.code[x] = (t.usable && d >= 1 && d <= num_searchable) ? (uint8_t)(d - 1) : 0xff;
    }
    if (!t.usable) return;
    // nibble-split classification: try "class = code" first, then "class = byte value" (<= 8 valid bytes)
    for (int attempt = 0; attempt < 2 && !t.simd_ok; ++attempt) {
        int cls[256];
        int nclasses = 0;
        bool fits = true;
        for (int x = 0; x < 256; ++x) {
            cls[x] = -1;
            if (t.code[x] == 0xff) continue;
            cls[x] = attempt == 0 ? t.code[x] : nclasses;
            ++nclasses;
            if (attempt == 1 && nclasses > 8) fits = false;
        }
        if (!fits) break;
        memset(t.lo_class, 0, 16);
        memset(t.hi_class, 0, 16);
        memset(t.code_lo, 0, 16);
        for (int x = 0; x < 256; ++x)
            if (cls[x] >= 0) {
                t.lo_class[x & 15] |= (uint8_t)(1u << cls[x]);
                t.hi_class[x >> 4] |= (uint8_t)(1u << cls[x]);
                t.code_lo[x & 15] = t.code[x];
            }
        bool ok = true;
        for (int x = 0; x < 256 && ok; ++x) {
            const bool valid = (t.lo_class[x & 15] & t.hi_class[x >> 4]) != 0;
            if (valid != (t.code[x] != 0xff)) ok = false;
            else if (valid && t.code_lo[x & 15] != t.code[x]) ok = false;
        }
        t.simd_ok = ok;
    }
}

static void pack2_scalar(const PackTable &t, const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t pos0,
                         std::vector<uint64_t> &exc) {
    uint64_t i = 0;
    for (; i + 4 <= n; i += 4) {
        const uint8_t a = t.code[src[i]], b = t.code[src[i + 1]], c = t.code[src[i + 2]], d = t.code[src[i + 3]];
        if ((a | b | c | d) & 0x80) {
            uint8_t v = 0;
            for (int k = 0; k < 4; ++k) {
                const uint8_t ck = t.code[src[i + k]];
                if (ck == 0xff) exc.push_back(pos0 + i + k);
                else v |= (uint8_t)(ck << (2 * k));
            }
            dst[i >> 2] = v;
        } else {
            dst[i >> 2] = (uint8_t)(a | (b << 2) | (c << 4) | (d << 6));
        }
    }
    if (i < n) {
        uint8_t v = 0;
        for (int k = 0; i + k < n; ++k) {
            const uint8_t ck = t.code[src[i + k]];
            if (ck == 0xff) exc.push_back(pos0 + i + k);
            else v |= (uint8_t)(ck << (2 * k));
        }
        dst[i >> 2] = v;
    }
}

__attribute__((target("avx2"))) static void pack2_avx2(const PackTable &t, const uint8_t *src, uint64_t n, uint8_t *dst,
                                                       uint64_t pos0, std::vector<uint64_t> &exc) {
    const __m256i lo_tab = _mm256_broadcastsi128_si256(_mm_loadu_si128((const __m128i *)t.lo_class));
    const __m256i hi_tab = _mm256_broadcastsi128_si256(_mm_loadu_si128((const __m128i *)t.hi_class));
    const __m256i code_tab = _mm256_broadcastsi128_si256(_mm_loadu_si128((const __m128i *)t.code_lo));
    const __m256i nib = _mm256_set1_epi8(0x0f), zero = _mm256_setzero_si256();
    const __m256i mul_1_4 = _mm256_set1_epi16(0x0401);       // bytes (1, 4): c0 + 4 c1 per 16-bit lane
    const __m256i mul_1_16 = _mm256_set1_epi32(0x00100001);  // words (1, 16): + 16 (c2 + 4 c3) per 32-bit lane
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    uint64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i x = _mm256_loadu_si256((const __m256i *)(src + i));
        const __m256i lo = _mm256_and_si256(x, nib);
        const __m256i hi = _mm256_and_si256(_mm256_srli_epi16(x, 4), nib);
        const __m256i cls = _mm256_and_si256(_mm256_shuffle_epi8(lo_tab, lo), _mm256_shuffle_epi8(hi_tab, hi));
        const __m256i inval = _mm256_cmpeq_epi8(cls, zero);
        __m256i code = _mm256_shuffle_epi8(code_tab, lo);
        const uint32_t m = (uint32_t)_mm256_movemask_epi8(inval);
        if (m) {
            code = _mm256_andnot_si256(inval, code);
            for (uint32_t mm = m; mm; mm &= mm - 1) exc.push_back(pos0 + i + (uint64_t)__builtin_ctz(mm));
        }
        const __m256i p16 = _mm256_maddubs_epi16(code, mul_1_4);
        const __m256i p32 = _mm256_madd_epi16(p16, mul_1_16);
        const __m256i g = _mm256_shuffle_epi8(p32, gather);
        const uint32_t o0 = (uint32_t)_mm256_extract_epi32(g, 0), o1 = (uint32_t)_mm256_extract_epi32(g, 4);
        const uint64_t o = (uint64_t)o0 | ((uint64_t)o1 << 32);
        memcpy(dst + (i >> 2), &o, 8);
    }
    if (i < n) pack2_scalar(t, src + i, n - i, dst + (i >> 2), pos0 + i, exc);
}

static std::atomic<int> g_pack_prefetch{getenv("GDX_PACK_PREFETCH") ? atoi(getenv("GDX_PACK_PREFETCH")) : 4096};
static std::atomic<int> g_pack_stream{getenv("GDX_PACK_STREAM") ? (atoi(getenv("GDX_PACK_STREAM")) != 0) : 1};
void set_pack_tuning(int prefetch_bytes, int stream) {
    g_pack_prefetch.store(prefetch_bytes < 0 ? 0 : prefetch_bytes);
    g_pack_stream.store(stream != 0);
}

// the same 64 bytes at a time: in-lane nibble shuffles, a mask register for the validity, vpmovdb for the gather
__attribute__((target("avx512f,avx512bw"))) static void pack2_avx512(const PackTable &t, const uint8_t *src, uint64_t n,
                                                                     uint8_t *dst, uint64_t pos0, std::vector<uint64_t> &exc) {
    const __m512i lo_tab = _mm512_broadcast_i32x4(_mm_loadu_si128((const __m128i *)t.lo_class));
    const __m512i hi_tab = _mm512_broadcast_i32x4(_mm_loadu_si128((const __m128i *)t.hi_class));
    const __m512i code_tab = _mm512_broadcast_i32x4(_mm_loadu_si128((const __m128i *)t.code_lo));
    const __m512i nib = _mm512_set1_epi8(0x0f);
    const __m512i mul_1_4 = _mm512_set1_epi16(0x0401), mul_1_16 = _mm512_set1_epi32(0x00100001);
    // A packer thread streams from DRAM, so it is bound by how many cache-line fills one core keeps in flight:
    // a software prefetch 4 KB ahead and non-temporal stores of the packed words (they are read next by the DMA
    // engine, not by a CPU: no need to pull their lines into a cache) lift the pool from 85 to 108 GB/s and a 30 M-query
    // gdx_count_many from 19.7 to 16.4 ms on a bench host (alternating runs inside one process,
    // profiles/r2_pack_tuning_ab.txt).  GDX_PACK_PREFETCH=<bytes> (0 = none), GDX_PACK_STREAM=0 and
    // gdx_host_pack_tuning() are the A/B switches.
    const uint64_t prefetch = (uint64_t)g_pack_prefetch.load(std::memory_order_relaxed);
    const bool stream = g_pack_stream.load(std::memory_order_relaxed) != 0;
    const bool nt = stream && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
    uint64_t i = 0;
    for (; i + 64 <= n; i += 64) {
        if (prefetch) _mm_prefetch((const char *)(src + i + prefetch), _MM_HINT_T0);
        const __m512i x = _mm512_loadu_si512((const void *)(src + i));
        const __m512i lo = _mm512_and_si512(x, nib);
        const __m512i hi = _mm512_and_si512(_mm512_srli_epi16(x, 4), nib);
        const __m512i cls = _mm512_and_si512(_mm512_shuffle_epi8(lo_tab, lo), _mm512_shuffle_epi8(hi_tab, hi));
        const __mmask64 valid = _mm512_test_epi8_mask(cls, cls);
        __m512i code = _mm512_shuffle_epi8(code_tab, lo);
        if (valid != ~(__mmask64)0) {
            code = _mm512_maskz_mov_epi8(valid, code);
            for (uint64_t mm = ~(uint64_t)valid; mm; mm &= mm - 1) exc.push_back(pos0 + i + (uint64_t)__builtin_ctzll(mm));
        }
        const __m512i p32 = _mm512_madd_epi16(_mm512_maddubs_epi16(code, mul_1_4), mul_1_16);
        if (nt) _mm_stream_si128((__m128i *)(dst + (i >> 2)), _mm512_cvtepi32_epi8(p32));
        else _mm_storeu_si128((__m128i *)(dst + (i >> 2)), _mm512_cvtepi32_epi8(p32));
    }
    if (nt) _mm_sfence();
    if (i < n) pack2_scalar(t, src + i, n - i, dst + (i >> 2), pos0 + i, exc);
}

// AVX-512 VBMI: the whole 256-entry table in four registers, two vpermi2b + one blend per 64 bytes (works for
// any alphabet, no nibble split needed), 128 bytes per iteration.  Table value 0x80 = no code.
#define GDX_VBMI __attribute__((target("avx512f,avx512bw,avx512vbmi"), always_inline)) static inline
GDX_VBMI __m512i vbmi_lookup(__m512i x, __m512i t0, __m512i t1, __m512i t2, __m512i t3) {
    const __m512i lo = _mm512_permutex2var_epi8(t0, x, t1);  // index bits 0..6 select among 128 entries
    const __m512i hi = _mm512_permutex2var_epi8(t2, x, t3);
    return _mm512_mask_blend_epi8(_mm512_movepi8_mask(x), lo, hi);  // bit 7 of the byte picks the upper half
}
GDX_VBMI __m128i vbmi_squeeze(__m512i code) {
    const __m512i mul_1_4 = _mm512_set1_epi16(0x0401), mul_1_16 = _mm512_set1_epi32(0x00100001);
    return _mm512_cvtepi32_epi8(_mm512_madd_epi16(_mm512_maddubs_epi16(code, mul_1_4), mul_1_16));
}
__attribute__((target("avx512f,avx512bw,avx512vbmi"))) static void pack2_vbmi(const PackTable &t, const uint8_t *src, uint64_t n,
                                                                              uint8_t *dst, uint64_t pos0,
                                                                              std::vector<uint64_t> &exc) {
    alignas(64) uint8_t tab[256];
    for (int x = 0; x < 256; ++x) tab[x] = t.code[x] == 0xff ? 0x80 : t.code[x];
    const __m512i t0 = _mm512_load_si512(tab), t1 = _mm512_load_si512(tab + 64), t2 = _mm512_load_si512(tab + 128),
                  t3 = _mm512_load_si512(tab + 192);
    uint64_t i = 0;
    for (; i + 128 <= n; i += 128) {
        __m512i c0 = vbmi_lookup(_mm512_loadu_si512((const void *)(src + i)), t0, t1, t2, t3);
        __m512i c1 = vbmi_lookup(_mm512_loadu_si512((const void *)(src + i + 64)), t0, t1, t2, t3);
        const __mmask64 bad0 = _mm512_movepi8_mask(c0), bad1 = _mm512_movepi8_mask(c1);
        if (bad0 | bad1) {
            c0 = _mm512_maskz_mov_epi8(~bad0, c0);
            c1 = _mm512_maskz_mov_epi8(~bad1, c1);
            for (uint64_t mm = bad0; mm; mm &= mm - 1) exc.push_back(pos0 + i + (uint64_t)__builtin_ctzll(mm));
            for (uint64_t mm = bad1; mm; mm &= mm - 1) exc.push_back(pos0 + i + 64 + (uint64_t)__builtin_ctzll(mm));
        }
        _mm_storeu_si128((__m128i *)(dst + (i >> 2)), vbmi_squeeze(c0));
        _mm_storeu_si128((__m128i *)(dst + (i >> 2) + 16), vbmi_squeeze(c1));
    }
    if (i < n) pack2_scalar(t, src + i, n - i, dst + (i >> 2), pos0 + i, exc);
}

void pack2_serial(const PackTable &t, const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t pos0,
                  std::vector<uint64_t> &exceptions) {
    // GDX_PACK_ISA = scalar | avx2 | avx512 | vbmi caps the instruction set (tests, A/B runs); GDX_PACK_SCALAR=1 = scalar
    static const int level = [] {
        int lvl = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
        if (lvl == 2 && __builtin_cpu_supports("avx512vbmi")) lvl = 3;
        if (const char *e = getenv("GDX_PACK_ISA")) {
            const int cap = !strcmp(e, "scalar") ? 0 : (!strcmp(e, "avx2") ? 1 : (!strcmp(e, "avx512") ? 2 : 3));
            lvl = std::min(lvl, cap);
        }
        if (getenv("GDX_PACK_SCALAR") && atoi(getenv("GDX_PACK_SCALAR"))) lvl = 0;
        return lvl;
    }();
    // the nibble-split AVX-512 loop is the fastest where the alphabet allows it (15.6 vs 14.6 GB/s per thread for the
    // full-table VBMI loop, which in turn serves every other alphabet at SIMD speed); GDX_PACK_ISA=vbmi forces VBMI
    static const bool force_vbmi = getenv("GDX_PACK_ISA") && !strcmp(getenv("GDX_PACK_ISA"), "vbmi");
    if (level == 3 && (force_vbmi || !t.simd_ok)) pack2_vbmi(t, src, n, dst, pos0, exceptions);
    else if (t.simd_ok && level >= 2) pack2_avx512(t, src, n, dst, pos0, exceptions);
    else if (t.simd_ok && level == 1) pack2_avx2(t, src, n, dst, pos0, exceptions);
    else pack2_scalar(t, src, n, dst, pos0, exceptions);
}

void pack2_parallel(const PackTable &t, const uint8_t *src, uint64_t n, uint8_t *dst, std::vector<uint64_t> &exceptions) {
    constexpr uint64_t kPiece = 256ull << 10;  // source bytes per piece, a multiple of 4: pieces own whole output bytes
    exceptions.clear();
    if (n <= kPiece) {
        pack2_serial(t, src, n, dst, 0, exceptions);
        return;
    }
    const uint64_t pieces = (n + kPiece - 1) / kPiece;
    std::vector<std::vector<uint64_t>> per_piece(pieces);
    HostPool::get().parallel_for(pieces, [&](uint64_t k) {
        const uint64_t off = k * kPiece, nb = std::min(kPiece, n - off);
        pack2_serial(t, src + off, nb, dst + (off >> 2), off, per_piece[k]);
    });
    for (auto &v : per_piece) exceptions.insert(exceptions.end(), v.begin(), v.end());
}

}  // namespace gdx
