// api.cu -- C ABI of genedex_b200 (include/genedex_b200.h): index handles, device image
// construction from host parts, chunked H2D / kernel / D2H pipelines, locate CSR plumbing.
#include <cuda_runtime.h>
#include <type_traits>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/genedex_b200.h"
#include "device_build.h"
#include "device_index.h"
#include "host_build.h"
#include "kernels.cuh"

using namespace gdx;

// ---- thread-local error / stats state ----------------------------------------------------------------
namespace {

thread_local std::string t_error;
thread_local uint64_t t_error_query = 0;
thread_local gdx_stats t_stats = {};

gdx_status fail(gdx_status st, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
    return st;
}

#define CUDA_TRY(expr)                                                                           \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return fail(_e == cudaErrorMemoryAllocation ? GDX_ERR_OOM : GDX_ERR_CUDA,            \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,        \
                        __LINE__);                                                               \
    } while (0)

#define GDX_TRY(expr)                  \
    do {                               \
        gdx_status _s = (expr);        \
        if (_s != GDX_OK) return _s;   \
    } while (0)

uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }
uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// Random 32 B record reads: ask the L2 to fetch single sectors from DRAM instead of sector pairs
// (cudaLimitMaxL2FetchGranularity is a per-context hint).  GDX_L2_FETCH_GRANULARITY=0 leaves the
// driver default, 32/64/128 sets it (measured: no effect on B200, profiles/r1_ab_l2_fetch_granularity.txt).
// Applied once per device, with that device current.
void apply_device_settings(int dev) {
    static std::mutex mu;
    static bool done[64] = {};
    if (dev < 0 || dev >= 64) return;
    std::lock_guard<std::mutex> lk(mu);
    if (done[dev]) return;
    done[dev] = true;
    const char *env = getenv("GDX_L2_FETCH_GRANULARITY");
    const int want = env ? atoi(env) : 32;
    if (want > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)want);
    // the device-resident entry points take their scratch from the stream-ordered pool: keep freed
    // blocks cached across synchronisation points instead of returning them to the driver
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 16ull << 20);  // room for the superblock tables
    cudaGetLastError();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        if (dev >= 0 && dev != prev) {
            if (cudaSetDevice(dev) != cudaSuccess) return;
        }
        apply_device_settings(dev >= 0 ? dev : prev);
        ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// grow-only device buffer
struct DBuf {
    void *p = nullptr;
    uint64_t cap = 0;
    cudaError_t reserve(uint64_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        uint64_t want = align_up(bytes + bytes / 8, 256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, want = align_up(bytes, 256));
            if (e != cudaSuccess) return e;
        }
        cap = want;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

// grow-only pinned host buffer (staging for callers that pass ordinary pageable memory)
struct HBuf {
    void *p = nullptr;
    uint64_t cap = 0;
    cudaError_t reserve(uint64_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const uint64_t want = align_up(bytes + bytes / 8, 4096);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// true for cudaMallocHost / cudaHostRegister'ed / managed memory, false for ordinary (pageable) host memory
bool is_pinned(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

// A few host threads that copy between pageable caller memory and the pinned staging buffers: one thread
// moves ~10 GB/s, a PCIe Gen5 link 52 GB/s.  One job at a time; the calling thread works too.
class HostPool {
public:
    static HostPool &get() {
        static HostPool pool;
        return pool;
    }
    void copy(void *dst, const void *src, uint64_t bytes) {
        constexpr uint64_t kPiece = 2ull << 20;
        if (bytes <= 2 * kPiece || workers_.empty()) {
            memcpy(dst, src, bytes);
            return;
        }
        std::lock_guard<std::mutex> one_job(job_mu_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = (uint8_t *)dst;
            src_ = (const uint8_t *)src;
            bytes_ = bytes;
            pieces_ = (bytes + kPiece - 1) / kPiece;
            next_.store(0);
            done_ = 0;
            ++generation_;
        }
        cv_.notify_all();
        work(kPiece);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return done_ == pieces_; });
    }

private:
    HostPool() {
        unsigned n = std::thread::hardware_concurrency();
        n = n > 2 ? std::min(7u, n / 2) : 0;  // plus the caller
        for (unsigned i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void work(uint64_t piece) {
        for (;;) {
            const uint64_t k = next_.fetch_add(1);
            if (k >= pieces_) break;
            const uint64_t off = k * piece, nb = std::min(piece, bytes_ - off);
            memcpy(dst_ + off, src_ + off, nb);
            std::lock_guard<std::mutex> lk(mu_);
            if (++done_ == pieces_) done_cv_.notify_all();
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            work(2ull << 20);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, job_mu_;
    std::condition_variable cv_, done_cv_;
    uint8_t *dst_ = nullptr;
    const uint8_t *src_ = nullptr;
    uint64_t bytes_ = 0, pieces_ = 0, done_ = 0, generation_ = 0;
    std::atomic<uint64_t> next_{0};
    bool stop_ = false;
};

constexpr uint64_t kStageMinBytes = 4ull << 20;  // smaller transfers go through the driver's own staging

constexpr int kSlots = 3;
// Pipeline chunking (defaults 4 / 32 / 4 MB, profiles/r1_ab_pipeline_chunks.txt): the byte budget of
// successive chunks doubles from kChunkFirst to kChunkMax
// (fast pipeline fill, then few large launches: less launch overhead and deeper shared suffixes for the
// per-chunk query sort) and the batch ends with a chunk of about kChunkTail (short drain).
uint64_t env_mb(const char *name, uint64_t dflt) {
    const char *e = getenv(name);
    return (uint64_t)(e && atoi(e) > 0 ? atoi(e) : dflt) << 20;
}
const uint64_t kChunkFirst = env_mb("GDX_CHUNK_FIRST_MB", 4);
const uint64_t kChunkMax = env_mb("GDX_CHUNK_MAX_MB", 32);
const uint64_t kChunkTail = env_mb("GDX_CHUNK_TAIL_MB", 4);
constexpr uint64_t kChunkMaxQueries = 64ull << 20;

struct Slot {
    cudaStream_t stream = nullptr;
    DBuf bytes, offsets, out_a, out_b, sort;
    // pipelined locate: per-chunk CSR + rows + hits
    DBuf counts, local_off, scan_tmp, rows, hits, big;
    uint64_t *d_words = nullptr;  // device: [0] number of wide intervals, [1] worklist cursor
    uint64_t *h_words = nullptr;  // pinned: [0] hits of the chunk, [1] number of wide intervals
    cudaEvent_t ev_total = nullptr;
    // staging for pageable caller buffers
    HBuf h_in, h_off, h_out_a, h_out_b;
    cudaEvent_t ev_h2d = nullptr, ev_out = nullptr;
    bool h2d_pending = false;
    std::vector<cudaEvent_t> ev;  // pairs (begin, end) around the kernels of the current call
    size_t ev_used = 0;
};

struct Small {  // pinned + device words shared by one call
    uint64_t *h = nullptr;  // pinned host, 16 words
    uint64_t *d = nullptr;  // device, 16 words: [0..2] err per slot, [4] steps, [5] walk steps,
                            //                   [6] big count, [7] big cursor, [8] total
};

struct Workspace {
    Slot slot[kSlots];
    Small small;
    DBuf starts, ends, counts, hit_offsets, rows, big_list, hits, scan_tmp, symbols;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    bool init() {
        for (auto &s : slot) {
            if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return false;
            if (cudaMalloc(&s.d_words, 4 * sizeof(uint64_t)) != cudaSuccess) return false;
            if (cudaMallocHost(&s.h_words, 4 * sizeof(uint64_t)) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&s.ev_total, cudaEventDisableTiming) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming) != cudaSuccess) return false;
        }
        if (cudaMallocHost(&small.h, 16 * sizeof(uint64_t)) != cudaSuccess) return false;
        if (cudaMalloc(&small.d, 16 * sizeof(uint64_t)) != cudaSuccess) return false;
        if (cudaEventCreate(&ev_a) != cudaSuccess || cudaEventCreate(&ev_b) != cudaSuccess) return false;
        return true;
    }
    void destroy() {
        for (auto &s : slot) {
            if (s.stream) cudaStreamDestroy(s.stream);
            s.bytes.release();
            s.offsets.release();
            s.out_a.release();
            s.out_b.release();
            s.sort.release();
            for (DBuf *b : {&s.counts, &s.local_off, &s.scan_tmp, &s.rows, &s.hits, &s.big}) b->release();
            if (s.d_words) cudaFree(s.d_words);
            if (s.h_words) cudaFreeHost(s.h_words);
            if (s.ev_total) cudaEventDestroy(s.ev_total);
            if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
            if (s.ev_out) cudaEventDestroy(s.ev_out);
            for (HBuf *b : {&s.h_in, &s.h_off, &s.h_out_a, &s.h_out_b}) b->release();
            for (auto e : s.ev) cudaEventDestroy(e);
        }
        if (small.h) cudaFreeHost(small.h);
        if (small.d) cudaFree(small.d);
        for (DBuf *b : {&starts, &ends, &counts, &hit_offsets, &rows, &big_list, &hits, &scan_tmp, &symbols})
            b->release();
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
    }
    cudaEvent_t next_event(Slot &s) {
        if (s.ev_used == s.ev.size()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;  // recording on nullptr fails loudly later
            s.ev.push_back(e);
        }
        return s.ev[s.ev_used++];
    }
};

struct PinnedHits {
    void *p = nullptr;
    uint64_t cap = 0;
    bool in_use = false;
};

}  // namespace

struct gdx_index {
    ImageHeader h;
    void *image = nullptr;
    bool own_image = false;
    int device = 0;
    DevIndex dev;
    void *dense_sa = nullptr;      // accelerator outside the image (gdx_index_set_dense_suffix_array)
    uint64_t dense_sa_bytes = 0;
    bool no_dense_sa = false;      // GDX_FLAG_NO_DENSE_SUFFIX_ARRAY
    void *seed_lut = nullptr;      // accelerator outside the image (gdx_index_set_seed_table_depth)
    uint64_t seed_lut_bytes = 0;
    bool no_seed_table = false;    // GDX_FLAG_NO_SEED_TABLE
    mutable std::mutex mu;
    mutable std::vector<Workspace *> free_ws;
    mutable std::vector<PinnedHits> pinned;
};

namespace {

Workspace *acquire_ws(const gdx_index *idx) {
    {
        std::lock_guard<std::mutex> lk(idx->mu);
        if (!idx->free_ws.empty()) {
            Workspace *w = idx->free_ws.back();
            idx->free_ws.pop_back();
            return w;
        }
    }
    Workspace *w = new Workspace();
    if (!w->init()) {
        w->destroy();
        delete w;
        return nullptr;
    }
    // The superblock table (count[c] + per-superblock ranks, ~1.5 MB for a 3.1 Gbp text) is read by
    // every LF step: keep it in the persisting part of L2 for the kernels on the workspace's streams.
    // Best effort (GDX_L2_PERSIST=0 disables it).
    static const bool persist = !(getenv("GDX_L2_PERSIST") && atoi(getenv("GDX_L2_PERSIST")) == 0);
    const uint64_t sbc_bytes = idx->h.n_superblocks * idx->h.layout.noff * 8;
    if (persist && sbc_bytes && sbc_bytes <= (16ull << 20)) {
        cudaStreamAttrValue attr = {};
        attr.accessPolicyWindow.base_ptr = const_cast<uint64_t *>(idx->dev.sbc);
        attr.accessPolicyWindow.num_bytes = sbc_bytes;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        for (auto &sl : w->slot) cudaStreamSetAttribute(sl.stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError();
    }
    return w;
}
void release_ws(const gdx_index *idx, Workspace *w) {
    std::lock_guard<std::mutex> lk(idx->mu);
    idx->free_ws.push_back(w);
}
struct WsLease {
    const gdx_index *idx;
    Workspace *w;
    explicit WsLease(const gdx_index *i) : idx(i), w(acquire_ws(i)) {}
    ~WsLease() {
        if (w) release_ws(idx, w);
    }
};

// ---- layout dispatch -----------------------------------------------------------------------------------
template <class F>
gdx_status dispatch_layout(const RankLayout &L, F &&f) {
    if (L.kind == kLayoutK32) return f(K32{});
    switch (L.planes) {
    case 1: return f(KG<1>{});
    case 2: return f(KG<2>{});
    case 3: return f(KG<3>{});
    case 4: return f(KG<4>{});
    case 5: return f(KG<5>{});
    case 6: return f(KG<6>{});
    case 7: return f(KG<7>{});
    case 8: return f(KG<8>{});
    }
    return fail(GDX_ERR_UNSUPPORTED, "unsupported number of bit planes %u", L.planes);
}

// ---- device image construction -------------------------------------------------------------------------
struct ImageSources {
    const gdx_alphabet *alphabet;
    uint32_t storage, sampling_rate, lookup_depth;
    uint64_t n;
    const uint64_t *count;             // host, sigma + 1
    const uint64_t *sentinels;         // host
    uint64_t ntexts;
    const uint64_t *border_rows;       // host, sorted ascending together with border_pos
    const uint64_t *border_pos;        // host
    uint64_t n_border;
    const uint8_t *d_bwt;              // device, n bytes
    const uint8_t *d_text = nullptr;   // device, n bytes dense text (optional: enables the text section)
    // exactly one of the sample sources
    const uint64_t *h_samples64 = nullptr;  // host
    const void *d_samples = nullptr;        // device, element width given by d_samples_wide
    bool d_samples_wide = false;
    // optional sampled inverse suffix array (one of them, only together with d_text)
    const uint64_t *h_isa64 = nullptr;      // host
    const uint32_t *d_isa32 = nullptr;      // device
};

gdx_status validate_alphabet(const gdx_alphabet &a) {
    if (a.num_dense_symbols < 2 || a.num_dense_symbols > 256)
        return fail(GDX_ERR_BAD_ARG, "alphabet size must be in [2,256] incl. the sentinel (alphabet.rs:166-174)");
    if (a.num_searchable_dense_symbols < 1 || a.num_searchable_dense_symbols > a.num_dense_symbols - 1)
        return fail(GDX_ERR_BAD_ARG, "there must be at least one searchable symbol (alphabet.rs:186-189)");
    for (int i = 0; i < 256; ++i)
        if (a.io_to_dense[i] >= a.num_dense_symbols)
            return fail(GDX_ERR_BAD_ARG, "io_to_dense[%d] = %u is not a dense symbol", i, a.io_to_dense[i]);
    return GDX_OK;
}

gdx_status plan_header(const ImageSources &src, ImageHeader &h) {
    memset(&h, 0, sizeof h);
    h.magic = kImageMagic;
    h.version = GDX_ABI_VERSION;
    h.n = src.n;
    h.ntexts = src.ntexts;
    h.n_border = src.n_border;
    h.sigma = src.alphabet->num_dense_symbols;
    h.ns = src.alphabet->num_searchable_dense_symbols;
    h.storage = src.storage;
    h.sampling_rate = src.sampling_rate;
    h.lookup_depth = src.lookup_depth;
    // 64-bit samples / lookup entries are only needed beyond 2^32 - 1 symbols; GDX_FORCE_WIDE=1 selects them
    // for any text so that this path can be tested without a 4.3 G symbol index
    const char *fw = getenv("GDX_FORCE_WIDE");
    h.wide = (src.n > 0xffffffffull || (fw && atoi(fw) != 0)) ? 1 : 0;
    h.layout = choose_layout(h.sigma);
    memcpy(h.io_to_dense, src.alphabet->io_to_dense, 256);
    if (src.lookup_depth > kMaxLookupDepth)
        return fail(GDX_ERR_UNSUPPORTED, "lookup table depth %u > %u", src.lookup_depth, kMaxLookupDepth);
    const uint64_t P = 1ull << h.layout.log2_pos;
    h.n_records = div_up(src.n + 1, P);
    h.n_superblocks = div_up(src.n + 1, 1ull << kSuperblockLog2);
    h.n_samples = div_up(src.n, src.sampling_rate);
    uint64_t entries = 0, pw = 1;
    for (uint32_t d = 0; d <= kMaxLookupDepth; ++d) {
        h.lut_level_off[d] = entries;
        h.lut_pow[d] = pw;
        if (d <= src.lookup_depth) {
            entries += pw;
            if (entries > (1ull << 36) || pw > (1ull << 36))
                return fail(GDX_ERR_UNSUPPORTED, "lookup tables of depth %u are too large", src.lookup_depth);
            pw *= h.ns;
        }
    }
    const uint64_t esz = h.wide ? 8 : 4;
    uint64_t off = 0;
    auto place = [&](uint64_t bytes) {
        uint64_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    h.off_records = place(h.n_records * h.layout.stride);
    // pack kernels write totals for whole superblock CTAs
    h.off_sbc = place(h.n_superblocks * h.layout.noff * 8);
    h.off_samples = place(h.n_samples * esz);
    h.off_lookup = place(entries * 2 * esz);
    h.off_border_rows = place(h.n_border * 8);
    h.off_border_pos = place(h.n_border * 8);
    h.off_sentinels = place(h.ntexts * 8);
    h.off_count = place((uint64_t)(h.sigma + 1) * 8);
    h.text_bits = 0;
    h.off_text = off;
    if (src.d_text) {
        h.text_bits = h.sigma <= 16 ? 4 : 8;
        // + 16: the comparison reads whole 64-bit words, one past the last (compare_with_text)
        h.off_text = place((h.text_bits == 4 ? (src.n + 1) / 2 : src.n) + 16);
    }
    h.has_isa = 0;
    h.off_isa = off;
    if (h.text_bits && (src.h_isa64 || src.d_isa32)) {
        h.has_isa = 1;
        h.off_isa = place(h.n_samples * esz);
    }
    h.image_bytes = off;
    return GDX_OK;
}

// run-time overrides of kernel parameters (measurements only)
gdx_status init_policies(gdx_index *idx) {
    if (const char *vm = getenv("GDX_VERIFY_MIN"))
        if (atoi(vm) > 0) idx->dev.verify_min_remaining = (uint32_t)atoi(vm);
    return GDX_OK;
}

// ---- dense suffix array accelerator (include/genedex_b200.h: gdx_index_set_dense_suffix_array) ----------
gdx_status build_dense_sa(gdx_index *idx) {
    if (idx->dense_sa || idx->h.n == 0) return GDX_OK;
    const uint64_t n = idx->h.n, bytes = n * (idx->h.wide ? 8 : 4);
    void *d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(GDX_ERR_OOM, "dense suffix array: %llu bytes of device memory not available", (unsigned long long)bytes);
    }
    gdx_status st = dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
        k_densify<decltype(L)><<<(unsigned)div_up(n, 256), 256>>>(idx->dev, d, n);
        return GDX_OK;
    });
    cudaError_t e = st == GDX_OK ? cudaDeviceSynchronize() : cudaSuccess;
    if (st != GDX_OK || e != cudaSuccess) {
        cudaFree(d);
        return st != GDX_OK ? st : fail(GDX_ERR_CUDA, "dense suffix array: %s", cudaGetErrorString(e));
    }
    idx->dense_sa = d;
    idx->dense_sa_bytes = bytes;
    idx->dev.samples = d;  // resolve_row / k_locate_walk now see a suffix array with sampling rate 1
    idx->dev.sampling_rate = 1;
    idx->dev.sampling_shift = 0;
    // resolving a row is now one load: the text comparison pays off from 4 remaining symbols on (measured on the
    // protein config, 12-symbol queries: 2.58 -> 1.23 ms per 10 M queries; no effect on 50-symbol DNA queries)
    if (!getenv("GDX_VERIFY_MIN")) idx->dev.verify_min_remaining = 4;
    return GDX_OK;
}

void drop_dense_sa(gdx_index *idx) {
    if (!idx->dense_sa) return;
    const DevIndex fresh = make_dev_index(idx->h, idx->image);
    idx->dev.samples = fresh.samples;
    idx->dev.sampling_rate = fresh.sampling_rate;
    idx->dev.sampling_shift = fresh.sampling_shift;
    if (!getenv("GDX_VERIFY_MIN")) idx->dev.verify_min_remaining = fresh.verify_min_remaining;
    cudaFree(idx->dense_sa);
    idx->dense_sa = nullptr;
    idx->dense_sa_bytes = 0;
}

// ---- seed table accelerator (include/genedex_b200.h: gdx_index_set_seed_table_depth) -------------------
void drop_seed_table(gdx_index *idx) {
    if (!idx->seed_lut) return;
    idx->dev.seed_lookup = nullptr;
    idx->dev.seed_depth = 0;
    cudaFree(idx->seed_lut);
    idx->seed_lut = nullptr;
    idx->seed_lut_bytes = 0;
}

// entries of level d, 0 if ns^d exceeds 2^36 (grid and memory limits)
uint64_t seed_entries(uint32_t ns, uint32_t d) {
    uint64_t e = 1;
    for (uint32_t i = 0; i < d; ++i) {
        e *= ns;
        if (e > (1ull << 36)) return 0;
    }
    return e;
}

gdx_status build_seed_table(gdx_index *idx, uint32_t depth) {
    drop_seed_table(idx);
    const ImageHeader &h = idx->h;
    // a level no deeper than the configured table would change nothing for the better and would move the
    // eager translation of the configured suffix (lookup_table.rs:154-157): not built
    if (depth <= h.lookup_depth || h.n == 0 || h.ns == 0) return GDX_OK;
    const uint64_t esz = h.wide ? 16 : 8, last = seed_entries(h.ns, depth), prev = seed_entries(h.ns, depth - 1);
    if (last == 0 || depth > 40) return fail(GDX_ERR_UNSUPPORTED, "seed table of depth %u is too large", depth);
    // levels alternate between the final buffer and a temporary one of the size of level depth - 1
    void *fin = nullptr, *tmp = nullptr;
    if (cudaMalloc(&fin, last * esz) != cudaSuccess || cudaMalloc(&tmp, prev * esz) != cudaSuccess) {
        cudaGetLastError();
        if (fin) cudaFree(fin);
        return fail(GDX_ERR_OOM, "seed table of depth %u: %llu bytes of device memory not available", depth,
                    (unsigned long long)((last + prev) * esz));
    }
    void *cur = (depth % 2 == 0) ? fin : tmp;  // level 0 lives where level `depth` will not collide: parity of depth
    cudaError_t e;
    if (h.wide) {
        const uint64_t e0[2] = {0, h.n};
        e = cudaMemcpy(cur, e0, sizeof e0, cudaMemcpyHostToDevice);
    } else {
        const uint32_t e0[2] = {0, (uint32_t)h.n};
        e = cudaMemcpy(cur, e0, sizeof e0, cudaMemcpyHostToDevice);
    }
    gdx_status st = GDX_OK;
    for (uint32_t d = 1; d <= depth && e == cudaSuccess && st == GDX_OK; ++d) {
        void *nxt = cur == fin ? tmp : fin;
        const uint64_t entries = seed_entries(h.ns, d);
        st = dispatch_layout(h.layout, [&](auto L) -> gdx_status {
            k_lut_extend<decltype(L)><<<(unsigned)div_up(entries, 256), 256>>>(idx->dev, cur, nxt, entries);
            return GDX_OK;
        });
        e = cudaGetLastError();
        cur = nxt;
    }
    if (e == cudaSuccess && st == GDX_OK) e = cudaDeviceSynchronize();
    cudaFree(tmp);
    if (st != GDX_OK || e != cudaSuccess || cur != fin) {
        cudaFree(fin);
        return st != GDX_OK ? st : fail(GDX_ERR_CUDA, "seed table: %s", cudaGetErrorString(e));
    }
    idx->seed_lut = fin;
    idx->seed_lut_bytes = last * esz;
    idx->dev.seed_lookup = fin;
    idx->dev.seed_depth = depth;
    return GDX_OK;
}

void auto_seed_table(gdx_index *idx) {
    if (!idx || idx->no_seed_table || idx->h.n == 0 || idx->h.ns < 2) return;
    const char *e = getenv("GDX_SEED_TABLE");
    uint32_t depth = 0;
    if (e) {
        if (atoi(e) <= 0) return;
        depth = (uint32_t)atoi(e);
    } else {
        while (seed_entries(idx->h.ns, depth + 1) && seed_entries(idx->h.ns, depth + 1) <= idx->h.n) ++depth;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return;
        const uint64_t esz = idx->h.wide ? 16 : 8;
        while (depth > 0 && (seed_entries(idx->h.ns, depth) + seed_entries(idx->h.ns, depth - 1)) * esz > free_b / 4) --depth;
    }
    if (depth <= idx->h.lookup_depth) return;  // the configured table is at least as deep
    if (build_seed_table(idx, depth) != GDX_OK) t_error.clear();  // optional: not an error of the call
}

// best effort after every way of creating a replica; the caller holds a DeviceGuard and has freed its temporaries
void auto_dense_sa(gdx_index *idx) {
    if (!idx || idx->no_dense_sa || idx->h.n == 0) return;
    const char *e = getenv("GDX_DENSE_SA");
    if (e && atoi(e) == 0) return;
    if (!(e && atoi(e) != 0)) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return;
        if (idx->h.n * (idx->h.wide ? 8ull : 4ull) > free_b / 4) return;
    }
    if (build_dense_sa(idx) != GDX_OK) t_error.clear();  // optional: not an error of the call
}

gdx_status build_image(const ImageSources &src, int device, gdx_index **out) {
    std::unique_ptr<gdx_index> idx(new gdx_index());
    GDX_TRY(plan_header(src, idx->h));
    ImageHeader &h = idx->h;
    idx->device = device;
    CUDA_TRY(cudaMalloc(&idx->image, h.image_bytes));
    idx->own_image = true;
    auto cleanup = [&](gdx_status st) {
        cudaFree(idx->image);
        idx->image = nullptr;
        return st;
    };
#define IMG_TRY(expr)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return cleanup(fail(GDX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                     \
                                cudaGetErrorString(_e), __FILE__, __LINE__));                     \
    } while (0)
    uint8_t *base = (uint8_t *)idx->image;
    IMG_TRY(cudaMemset(base, 0, h.image_bytes));
    IMG_TRY(cudaMemcpy(base + h.off_sentinels, src.sentinels, h.ntexts * 8, cudaMemcpyHostToDevice));
    if (h.n_border) {
        IMG_TRY(cudaMemcpy(base + h.off_border_rows, src.border_rows, h.n_border * 8, cudaMemcpyHostToDevice));
        IMG_TRY(cudaMemcpy(base + h.off_border_pos, src.border_pos, h.n_border * 8, cudaMemcpyHostToDevice));
    }
    IMG_TRY(cudaMemcpy(base + h.off_count, src.count, (uint64_t)(h.sigma + 1) * 8, cudaMemcpyHostToDevice));

    // rank records + superblock table
    uint64_t *sbc = (uint64_t *)(base + h.off_sbc);
    if (h.layout.kind == kLayoutK32) {
        k_pack_k32<<<(unsigned)h.n_superblocks, 1024>>>(src.d_bwt, h.n, h.layout.noff, base + h.off_records,
                                                        h.n_records, sbc);
    } else {
        gdx_status st = dispatch_layout(h.layout, [&](auto L) -> gdx_status {
            using LT = decltype(L);
            if constexpr (!std::is_same<LT, K32>::value) {
                constexpr int B = sizeof(typename LT::Planes) / 16;
                k_pack_kg<B><<<(unsigned)h.n_superblocks, 512>>>(src.d_bwt, h.n, h.sigma, h.layout.stride,
                                                                 base + h.off_records, h.n_records, sbc);
            }
            return GDX_OK;
        });
        if (st != GDX_OK) return cleanup(st);
    }
    IMG_TRY(cudaGetLastError());
    k_sb_scan<<<div_up(h.layout.noff, 64), 64>>>(sbc, h.n_superblocks, h.layout.noff,
                                                 (const uint64_t *)(base + h.off_count));
    IMG_TRY(cudaGetLastError());

    // SA samples
    if (h.n_samples) {
        void *dst = base + h.off_samples;
        if (src.h_samples64) {
            if (h.wide) {
                IMG_TRY(cudaMemcpy(dst, src.h_samples64, h.n_samples * 8, cudaMemcpyHostToDevice));
            } else {
                std::vector<uint32_t> narrow(h.n_samples);
                for (uint64_t i = 0; i < h.n_samples; ++i) narrow[i] = (uint32_t)src.h_samples64[i];
                IMG_TRY(cudaMemcpy(dst, narrow.data(), h.n_samples * 4, cudaMemcpyHostToDevice));
            }
        } else if (src.d_samples) {
            const unsigned g = (unsigned)div_up(h.n_samples, 256);
            if (src.d_samples_wide == (h.wide != 0))
                IMG_TRY(cudaMemcpy(dst, src.d_samples, h.n_samples * (h.wide ? 8 : 4), cudaMemcpyDeviceToDevice));
            else if (h.wide)
                k_widen_u32<<<g, 256>>>((const uint32_t *)src.d_samples, h.n_samples, (uint64_t *)dst);
            else
                k_narrow_u64<<<g, 256>>>((const uint64_t *)src.d_samples, h.n_samples, (uint32_t *)dst);
            IMG_TRY(cudaGetLastError());
        }
    }

    if (h.has_isa && h.n_samples) {  // same element width as the SA samples
        void *dst = base + h.off_isa;
        if (src.h_isa64) {
            if (h.wide) {
                IMG_TRY(cudaMemcpy(dst, src.h_isa64, h.n_samples * 8, cudaMemcpyHostToDevice));
            } else {
                std::vector<uint32_t> narrow(h.n_samples);
                for (uint64_t i = 0; i < h.n_samples; ++i) narrow[i] = (uint32_t)src.h_isa64[i];
                IMG_TRY(cudaMemcpy(dst, narrow.data(), h.n_samples * 4, cudaMemcpyHostToDevice));
            }
        } else if (h.wide) {
            k_widen_u32<<<(unsigned)div_up(h.n_samples, 256), 256>>>(src.d_isa32, h.n_samples, (uint64_t *)dst);
            IMG_TRY(cudaGetLastError());
        } else {
            IMG_TRY(cudaMemcpy(dst, src.d_isa32, h.n_samples * 4, cudaMemcpyDeviceToDevice));
        }
    }
    if (h.text_bits && h.n) {
        const uint64_t items = h.text_bits == 4 ? (h.n + 1) / 2 : h.n;
        k_pack_text<<<(unsigned)div_up(items, 256), 256>>>(src.d_text, h.n, h.text_bits, base + h.off_text);
        IMG_TRY(cudaGetLastError());
    }

    idx->dev = make_dev_index(h, idx->image);
    {
        gdx_status ps = init_policies(idx.get());
        if (ps != GDX_OK) return cleanup(ps);
    }

    // lookup tables: level 0 = [(0, n)] (lookup_table.rs:205-209), level d from level d-1
    {
        void *lut = base + h.off_lookup;
        if (h.wide) {
            uint64_t e0[2] = {0, h.n};
            IMG_TRY(cudaMemcpy(lut, e0, sizeof e0, cudaMemcpyHostToDevice));
        } else {
            uint32_t e0[2] = {0, (uint32_t)h.n};
            IMG_TRY(cudaMemcpy(lut, e0, sizeof e0, cudaMemcpyHostToDevice));
        }
        for (uint32_t d = 1; d <= h.lookup_depth; ++d) {
            gdx_status st = dispatch_layout(h.layout, [&](auto L) -> gdx_status {
                using LT = decltype(L);
                k_lut_fill<LT><<<(unsigned)div_up(h.lut_pow[d], 256), 256>>>(idx->dev, lut, d);
                return GDX_OK;
            });
            if (st != GDX_OK) return cleanup(st);
            IMG_TRY(cudaGetLastError());
        }
    }
    IMG_TRY(cudaDeviceSynchronize());
#undef IMG_TRY
    *out = idx.release();
    return GDX_OK;
}

gdx_status check_config(const gdx_config &c) {
    if (c.suffix_array_sampling_rate == 0)
        return fail(GDX_ERR_BAD_ARG, "suffix array sampling rate must be > 0 (config.rs:28)");
    if (c.storage > GDX_I64) return fail(GDX_ERR_BAD_ARG, "unknown index storage %u", c.storage);
    if (c.construction > GDX_CONSTRUCT_AUTO) return fail(GDX_ERR_BAD_ARG, "unknown construction mode %u", c.construction);
    return GDX_OK;
}

gdx_status resolve_device(int32_t requested, int *device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(GDX_ERR_CUDA, "no CUDA device available: %s (genedex_b200 has no CPU fallback)",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (requested < 0) {
        CUDA_TRY(cudaGetDevice(device));
    } else {
        if (requested >= count) return fail(GDX_ERR_BAD_ARG, "device %d out of range (%d devices)", requested, count);
        *device = requested;
    }
    return GDX_OK;
}

}  // namespace

// ================================================================================================
// misc
// ================================================================================================
extern "C" uint32_t gdx_abi_version(void) { return GDX_ABI_VERSION; }
extern "C" const char *gdx_last_error_message(void) { return t_error.c_str(); }
extern "C" uint64_t gdx_last_error_query(void) { return t_error_query; }
extern "C" int32_t gdx_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
    return c;
}
extern "C" gdx_status gdx_get_stats(gdx_stats *out) {
    if (!out) return fail(GDX_ERR_BAD_ARG, "out is NULL");
    *out = t_stats;
    return GDX_OK;
}
extern "C" gdx_status gdx_host_alloc(uint64_t bytes, void **out) {
    if (!out) return fail(GDX_ERR_BAD_ARG, "out is NULL");
    CUDA_TRY(cudaMallocHost(out, bytes ? bytes : 1));
    return GDX_OK;
}
extern "C" void gdx_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ================================================================================================
// construction
// ================================================================================================
extern "C" gdx_status gdx_index_build(const uint8_t *texts, const uint64_t *text_offsets, uint64_t num_texts,
                                      const gdx_alphabet *alphabet, const gdx_config *config, gdx_index **out) {
    if (!out) return fail(GDX_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (!text_offsets || !alphabet || !config) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    if (num_texts == 0) return fail(GDX_ERR_BAD_ARG, "there should be at least one text (construction/mod.rs:300)");
    GDX_TRY(validate_alphabet(*alphabet));
    GDX_TRY(check_config(*config));
    int device;
    GDX_TRY(resolve_device(config->device, &device));
    DeviceGuard guard(device);

    ConcatText ct;
    uint64_t bad = 0;
    gdx_status st = concat_texts(texts, text_offsets, num_texts, *alphabet, ct, &bad);
    if (st != GDX_OK) {
        t_error_query = bad;
        return fail(st, "text %llu contains a symbol that is not in the alphabet (alphabet.rs:195-198)",
                    (unsigned long long)bad);
    }
    const uint64_t n = ct.text.size();
    if (n > storage_max(config->storage))
        return fail(GDX_ERR_TEXT_TOO_LONG, "text length %llu exceeds the index storage type (construction/mod.rs:34)",
                    (unsigned long long)n);

    ImageSources src;
    src.alphabet = alphabet;
    src.storage = config->storage;
    src.sampling_rate = config->suffix_array_sampling_rate;
    src.lookup_depth = config->lookup_table_depth;
    src.n = n;
    src.count = ct.count.data();
    src.sentinels = ct.sentinels.data();
    src.ntexts = num_texts;

    // the dense text goes to the device once: input of the device suffix sort and source of the
    // optional text section of the image
    const bool keep_text = (config->flags & GDX_FLAG_NO_TEXT) == 0;
    const bool keep_isa = keep_text && (config->flags & GDX_FLAG_NO_INVERSE_SAMPLES) == 0;
    struct DevText {
        uint8_t *p = nullptr;
        ~DevText() {
            if (p) cudaFree(p);
        }
    } d_text;
    uint32_t construction = config->construction;
    if (construction == GDX_CONSTRUCT_AUTO) {  // device suffix sort when the text and its scratch fit
        size_t free_b = 0, total_b = 0;
        construction = GDX_CONSTRUCT_HOST;
        if (n < 0xffffffffull && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > 36 * n + (1ull << 30))
            construction = GDX_CONSTRUCT_DEVICE;
    }
    if (keep_text || construction == GDX_CONSTRUCT_DEVICE) {
        CUDA_TRY(cudaMalloc(&d_text.p, n ? n : 1));
        CUDA_TRY(cudaMemcpy(d_text.p, ct.text.data(), n, cudaMemcpyHostToDevice));
        if (keep_text) src.d_text = d_text.p;
    }

    if (construction == GDX_CONSTRUCT_DEVICE) {
        DeviceBuildResult r;
        const bool verify = (config->flags & GDX_FLAG_VERIFY_SUFFIX_ARRAY) != 0;
        st = device_build_from_text(nullptr, d_text.p, n, alphabet->num_dense_symbols,
                                    config->suffix_array_sampling_rate, r, t_error, false, verify, keep_isa);
        if (st != GDX_OK) return st;
        if (verify && r.verify_violations) {
            r.release();
            return fail(GDX_ERR_CUDA, "device suffix array failed verification (%llu violations)",
                        (unsigned long long)r.verify_violations);
        }
        src.border_rows = r.border_rows.data();
        src.border_pos = r.border_pos.data();
        src.n_border = r.border_rows.size();
        src.d_bwt = r.d_bwt;
        src.d_samples = r.d_samples;
        src.d_samples_wide = false;
        src.d_isa32 = r.d_isa_samples;
        st = build_image(src, device, out);
        r.release();
        if (st == GDX_OK) {
            (*out)->no_dense_sa = (config->flags & GDX_FLAG_NO_DENSE_SUFFIX_ARRAY) != 0;
            (*out)->no_seed_table = (config->flags & GDX_FLAG_NO_SEED_TABLE) != 0;
            auto_dense_sa(*out);
            auto_seed_table(*out);
        }
        return st;
    }

    HostParts hp;
    host_parts_from_text(ct.text.data(), n, alphabet->num_dense_symbols, config->suffix_array_sampling_rate, hp);
    uint8_t *d_bwt = nullptr;
    CUDA_TRY(cudaMalloc(&d_bwt, n ? n : 1));
    cudaError_t e = cudaMemcpy(d_bwt, hp.bwt.data(), n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(d_bwt);
        return fail(GDX_ERR_CUDA, "BWT upload failed: %s", cudaGetErrorString(e));
    }
    src.border_rows = hp.border_rows.data();
    src.border_pos = hp.border_pos.data();
    src.n_border = hp.border_rows.size();
    src.d_bwt = d_bwt;
    src.h_samples64 = hp.samples.data();
    if (keep_isa) src.h_isa64 = hp.isa_samples.data();
    st = build_image(src, device, out);
    cudaFree(d_bwt);
    if (st == GDX_OK) {
        (*out)->no_dense_sa = (config->flags & GDX_FLAG_NO_DENSE_SUFFIX_ARRAY) != 0;
        (*out)->no_seed_table = (config->flags & GDX_FLAG_NO_SEED_TABLE) != 0;
        auto_dense_sa(*out);
        auto_seed_table(*out);
    }
    return st;
}

static gdx_status sources_from_parts(const gdx_parts *parts, ImageSources &src,
                                     std::vector<uint64_t> &rows, std::vector<uint64_t> &pos) {
    if (!parts) return fail(GDX_ERR_BAD_ARG, "parts is NULL");
    GDX_TRY(validate_alphabet(parts->alphabet));
    if (parts->sampling_rate == 0) return fail(GDX_ERR_BAD_ARG, "sampling rate must be > 0");
    if (!parts->count || !parts->sentinel_indices || parts->num_texts == 0)
        return fail(GDX_ERR_BAD_ARG, "count / sentinel_indices missing");
    if (parts->text_len > storage_max(parts->storage)) return fail(GDX_ERR_TEXT_TOO_LONG, "text too long for storage");
    // sort the border map by row (it is a HashMap in the reference)
    std::vector<std::pair<uint64_t, uint64_t>> b(parts->num_text_borders);
    for (uint64_t i = 0; i < parts->num_text_borders; ++i)
        b[i] = {parts->text_border_rows[i], parts->text_border_positions[i]};
    std::sort(b.begin(), b.end());
    rows.resize(b.size());
    pos.resize(b.size());
    for (size_t i = 0; i < b.size(); ++i) {
        rows[i] = b[i].first;
        pos[i] = b[i].second;
    }
    src.alphabet = &parts->alphabet;
    src.storage = parts->storage;
    src.sampling_rate = parts->sampling_rate;
    src.lookup_depth = parts->lookup_table_depth;
    src.n = parts->text_len;
    src.count = parts->count;
    src.sentinels = parts->sentinel_indices;
    src.ntexts = parts->num_texts;
    src.border_rows = rows.data();
    src.border_pos = pos.data();
    src.n_border = rows.size();
    src.h_samples64 = parts->sampled_suffix_array;
    return GDX_OK;
}

extern "C" gdx_status gdx_index_create_from_bwt(const uint8_t *bwt, const gdx_parts *parts, int32_t device_req,
                                                gdx_index **out) {
    if (!out) return fail(GDX_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (!bwt) return fail(GDX_ERR_BAD_ARG, "bwt is NULL");
    ImageSources src;
    std::vector<uint64_t> rows, pos;
    GDX_TRY(sources_from_parts(parts, src, rows, pos));
    int device;
    GDX_TRY(resolve_device(device_req, &device));
    DeviceGuard guard(device);
    uint8_t *d_bwt = nullptr;
    CUDA_TRY(cudaMalloc(&d_bwt, src.n ? src.n : 1));
    cudaError_t e = cudaMemcpy(d_bwt, bwt, src.n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(d_bwt);
        return fail(GDX_ERR_CUDA, "BWT upload failed: %s", cudaGetErrorString(e));
    }
    src.d_bwt = d_bwt;
    gdx_status st = build_image(src, device, out);
    cudaFree(d_bwt);
    if (st == GDX_OK) {
        auto_dense_sa(*out);
        auto_seed_table(*out);
    }
    return st;
}

extern "C" gdx_status gdx_index_create_from_parts(const gdx_parts *parts, int32_t device_req, gdx_index **out) {
    if (!out) return fail(GDX_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    ImageSources src;
    std::vector<uint64_t> rows, pos;
    GDX_TRY(sources_from_parts(parts, src, rows, pos));
    if (!parts->interleaved_blocks) return fail(GDX_ERR_BAD_ARG, "interleaved_blocks is NULL");
    int device;
    GDX_TRY(resolve_device(device_req, &device));
    DeviceGuard guard(device);
    // reference planes -> dense BWT on the device (condensed.rs:343-362), then the common path
    const uint32_t nplanes = choose_layout(parts->alphabet.num_dense_symbols).planes;
    const uint64_t nwords = div_up(src.n + 1, 64) * nplanes;
    uint64_t *d_blocks = nullptr;
    uint8_t *d_bwt = nullptr;
    CUDA_TRY(cudaMalloc(&d_blocks, nwords * 8));
    cudaError_t e = cudaMalloc(&d_bwt, src.n ? src.n : 1);
    if (e == cudaSuccess) e = cudaMemcpy(d_blocks, parts->interleaved_blocks, nwords * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && src.n) {
        k_planes_to_bwt<<<(unsigned)div_up(src.n, 256), 256>>>(d_blocks, nplanes, src.n, d_bwt);
        e = cudaGetLastError();
    }
    cudaFree(d_blocks);
    if (e != cudaSuccess) {
        if (d_bwt) cudaFree(d_bwt);
        return fail(GDX_ERR_CUDA, "plane upload failed: %s", cudaGetErrorString(e));
    }
    src.d_bwt = d_bwt;
    gdx_status st = build_image(src, device, out);
    cudaFree(d_bwt);
    if (st == GDX_OK) {
        auto_dense_sa(*out);
        auto_seed_table(*out);
    }
    return st;
}

extern "C" gdx_status gdx_concat_texts(const uint8_t *texts, const uint64_t *text_offsets, uint64_t num_texts,
                                       const gdx_alphabet *alphabet, uint8_t *dense_out, uint64_t *sentinels_out,
                                       uint64_t *count_out) {
    if (!text_offsets || !alphabet || !dense_out || !sentinels_out || !count_out || num_texts == 0)
        return fail(GDX_ERR_BAD_ARG, "NULL argument or no texts");
    GDX_TRY(validate_alphabet(*alphabet));
    ConcatText ct;
    uint64_t bad = 0;
    gdx_status st = concat_texts(texts, text_offsets, num_texts, *alphabet, ct, &bad);
    if (st != GDX_OK) {
        t_error_query = bad;
        return fail(st, "text %llu contains a symbol that is not in the alphabet (alphabet.rs:195-198)",
                    (unsigned long long)bad);
    }
    memcpy(dense_out, ct.text.data(), ct.text.size());
    memcpy(sentinels_out, ct.sentinels.data(), ct.sentinels.size() * 8);
    memcpy(count_out, ct.count.data(), ct.count.size() * 8);
    return GDX_OK;
}

extern "C" gdx_status gdx_suffix_array(const uint8_t *dense_text, uint64_t n, uint32_t sigma, uint32_t where,
                                       int32_t device_req, uint64_t *sa_out) {
    if (n == 0) return GDX_OK;
    if (!dense_text || !sa_out) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    if (sigma < 2 || sigma > 256) return fail(GDX_ERR_BAD_ARG, "alphabet size must be in [2,256]");
    for (uint64_t i = 0; i < n; ++i)
        if (dense_text[i] >= sigma) return fail(GDX_ERR_BAD_ARG, "symbol %u at %llu is not dense", dense_text[i], (unsigned long long)i);
    if (where == GDX_CONSTRUCT_HOST) {
        std::vector<int64_t> sa;
        suffix_array_sais(dense_text, n, sigma, sa);
        for (uint64_t i = 0; i < n; ++i) sa_out[i] = (uint64_t)sa[i];
        return GDX_OK;
    }
    int device;
    GDX_TRY(resolve_device(device_req, &device));
    DeviceGuard guard(device);
    DeviceBuildResult r;
    gdx_status st = device_build_from_text(dense_text, nullptr, n, sigma, 1, r, t_error, true, true);
    if (st != GDX_OK) return st;
    std::vector<uint32_t> sa32(n);
    cudaError_t e = cudaMemcpy(sa32.data(), r.d_sa, n * 4, cudaMemcpyDeviceToHost);
    const uint64_t viol = r.verify_violations;
    r.release();
    if (e != cudaSuccess) return fail(GDX_ERR_CUDA, "suffix array download failed: %s", cudaGetErrorString(e));
    for (uint64_t i = 0; i < n; ++i) sa_out[i] = sa32[i];
    if (viol) return fail(GDX_ERR_CUDA, "device suffix array failed verification (%llu violations)", (unsigned long long)viol);
    return GDX_OK;
}

extern "C" gdx_status gdx_index_download_bwt(const gdx_index *idx, uint8_t *bwt_out) {
    if (!idx || !bwt_out) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(idx->device);
    const uint64_t n = idx->h.n, chunk = 256ull << 20;
    uint8_t *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, std::min<uint64_t>(n ? n : 1, chunk)));
    gdx_status st = GDX_OK;
    for (uint64_t b = 0; b < n && st == GDX_OK; b += chunk) {
        const uint64_t e = std::min<uint64_t>(n, b + chunk);
        st = dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
            k_records_to_bwt<decltype(L)><<<(unsigned)div_up(e - b, 256), 256>>>(idx->dev, b, e, d);
            return GDX_OK;
        });
        cudaError_t ce = cudaMemcpy(bwt_out + b, d, e - b, cudaMemcpyDeviceToHost);
        if (st == GDX_OK && ce != cudaSuccess) st = fail(GDX_ERR_CUDA, "BWT download failed: %s", cudaGetErrorString(ce));
    }
    cudaFree(d);
    return st;
}

extern "C" gdx_status gdx_index_download_samples(const gdx_index *idx, uint64_t *samples_out) {
    if (!idx || !samples_out) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(idx->device);
    const uint64_t ns = idx->h.n_samples;
    const uint8_t *src = (const uint8_t *)idx->image + idx->h.off_samples;
    if (idx->h.wide) {
        CUDA_TRY(cudaMemcpy(samples_out, src, ns * 8, cudaMemcpyDeviceToHost));
    } else {
        // narrow samples land in the upper half of the output buffer and are widened in place
        uint32_t *tmp = reinterpret_cast<uint32_t *>(samples_out) + ns;
        CUDA_TRY(cudaMemcpy(tmp, src, ns * 4, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < ns; ++i) samples_out[i] = tmp[i];
    }
    return GDX_OK;
}

extern "C" gdx_status gdx_index_download_text_borders(const gdx_index *idx, uint64_t *rows_out, uint64_t *positions_out) {
    if (!idx || !rows_out || !positions_out) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(idx->device);
    const uint8_t *base = (const uint8_t *)idx->image;
    CUDA_TRY(cudaMemcpy(rows_out, base + idx->h.off_border_rows, idx->h.n_border * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(positions_out, base + idx->h.off_border_pos, idx->h.n_border * 8, cudaMemcpyDeviceToHost));
    return GDX_OK;
}

extern "C" gdx_status gdx_index_get_count(const gdx_index *idx, uint64_t *count_out) {
    if (!idx || !count_out) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(idx->device);
    CUDA_TRY(cudaMemcpy(count_out, (const uint8_t *)idx->image + idx->h.off_count, (uint64_t)(idx->h.sigma + 1) * 8,
                        cudaMemcpyDeviceToHost));
    return GDX_OK;
}

extern "C" void gdx_index_destroy(gdx_index *idx) {
    if (!idx) return;
    DeviceGuard guard(idx->device);
    for (Workspace *w : idx->free_ws) {
        w->destroy();
        delete w;
    }
    for (auto &p : idx->pinned)
        if (p.p) cudaFreeHost(p.p);
    if (idx->dense_sa) cudaFree(idx->dense_sa);
    if (idx->seed_lut) cudaFree(idx->seed_lut);
    if (idx->own_image && idx->image) cudaFree(idx->image);
    delete idx;
}

extern "C" gdx_status gdx_index_get_info(const gdx_index *idx, gdx_index_info *out) {
    if (!idx || !out) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    const ImageHeader &h = idx->h;
    out->text_len = h.n;
    out->num_texts = h.ntexts;
    out->num_dense_symbols = h.sigma;
    out->num_searchable_dense_symbols = h.ns;
    out->storage = h.storage;
    out->sampling_rate = h.sampling_rate;
    out->lookup_table_depth = h.lookup_depth;
    out->rank_layout = h.layout.kind;
    out->rank_record_bytes = h.layout.stride;
    out->rank_positions_per_record = 1u << h.layout.log2_pos;
    out->device = idx->device;
    out->image_bytes = h.image_bytes;
    out->rank_bytes = h.off_samples - h.off_records;
    out->sample_bytes = h.off_lookup - h.off_samples;
    out->lookup_bytes = h.off_border_rows - h.off_lookup;
    out->num_samples = h.n_samples;
    out->num_text_borders = h.n_border;
    out->text_bytes = h.text_bits ? h.off_isa - h.off_text : 0;
    out->inverse_sample_bytes = h.has_isa ? h.image_bytes - h.off_isa : 0;
    out->dense_suffix_array_bytes = idx->dense_sa_bytes;
    out->seed_table_bytes = idx->seed_lut_bytes;
    out->seed_table_depth = idx->dev.seed_depth;
    out->reserved = 0;
    return GDX_OK;
}

// ================================================================================================
// index files
// ================================================================================================
namespace {
constexpr char kFileMagic[8] = {'G', 'D', 'X', 'F', 'I', 'L', 'E', '1'};
struct FilePrefix {
    char magic[8];
    uint64_t header_bytes, image_bytes, user_bytes;
};
struct FileCloser {
    FILE *f;
    ~FileCloser() {
        if (f) fclose(f);
    }
};
}  // namespace

extern "C" gdx_status gdx_index_save_to_file(const gdx_index *idx, const char *path, const void *user_data,
                                             uint64_t user_bytes) {
    if (!idx || !path || (user_bytes && !user_data)) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(idx->device);
    FileCloser fc{fopen(path, "wb")};
    if (!fc.f) return fail(GDX_ERR_BAD_ARG, "cannot open %s for writing", path);
    FilePrefix p;
    memcpy(p.magic, kFileMagic, 8);
    p.header_bytes = sizeof(ImageHeader);
    p.image_bytes = idx->h.image_bytes;
    p.user_bytes = user_bytes;
    if (fwrite(&p, sizeof p, 1, fc.f) != 1 || fwrite(&idx->h, sizeof(ImageHeader), 1, fc.f) != 1 ||
        (user_bytes && fwrite(user_data, user_bytes, 1, fc.f) != 1))
        return fail(GDX_ERR_BAD_ARG, "write to %s failed", path);
    const uint64_t chunk = 64ull << 20;
    void *stage = nullptr;
    CUDA_TRY(cudaMallocHost(&stage, chunk));
    gdx_status st = GDX_OK;
    for (uint64_t off = 0; off < p.image_bytes && st == GDX_OK; off += chunk) {
        const uint64_t nb = std::min<uint64_t>(chunk, p.image_bytes - off);
        cudaError_t e = cudaMemcpy(stage, (const uint8_t *)idx->image + off, nb, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) st = fail(GDX_ERR_CUDA, "image download failed: %s", cudaGetErrorString(e));
        else if (fwrite(stage, nb, 1, fc.f) != 1) st = fail(GDX_ERR_BAD_ARG, "write to %s failed", path);
    }
    cudaFreeHost(stage);
    return st;
}

extern "C" gdx_status gdx_index_load_from_file(const char *path, int32_t device_req, gdx_index **out,
                                               void *user_data_out, uint64_t user_capacity, uint64_t *user_bytes_out) {
    if (!path || !out) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    *out = nullptr;
    FileCloser fc{fopen(path, "rb")};
    if (!fc.f) return fail(GDX_ERR_BAD_ARG, "cannot open %s", path);
    FilePrefix p;
    ImageHeader h;
    if (fread(&p, sizeof p, 1, fc.f) != 1 || memcmp(p.magic, kFileMagic, 8) != 0 || p.header_bytes != sizeof(ImageHeader) ||
        fread(&h, sizeof h, 1, fc.f) != 1 || h.magic != kImageMagic || h.version != GDX_ABI_VERSION ||
        h.image_bytes != p.image_bytes)
        return fail(GDX_ERR_BAD_ARG, "%s is not a genedex_b200 index file of this version", path);
    if (user_bytes_out) *user_bytes_out = p.user_bytes;
    if (p.user_bytes) {
        std::vector<uint8_t> blob(p.user_bytes);
        if (fread(blob.data(), p.user_bytes, 1, fc.f) != 1) return fail(GDX_ERR_BAD_ARG, "%s is truncated", path);
        if (user_data_out) memcpy(user_data_out, blob.data(), std::min<uint64_t>(p.user_bytes, user_capacity));
    }
    int device;
    GDX_TRY(resolve_device(device_req, &device));
    DeviceGuard guard(device);
    void *image = nullptr, *stage = nullptr;
    CUDA_TRY(cudaMalloc(&image, p.image_bytes ? p.image_bytes : 1));
    const uint64_t chunk = 64ull << 20;
    cudaError_t e = cudaMallocHost(&stage, chunk);
    gdx_status st = e == cudaSuccess ? GDX_OK : fail(GDX_ERR_OOM, "pinned staging allocation failed");
    for (uint64_t off = 0; off < p.image_bytes && st == GDX_OK; off += chunk) {
        const uint64_t nb = std::min<uint64_t>(chunk, p.image_bytes - off);
        if (fread(stage, nb, 1, fc.f) != 1) st = fail(GDX_ERR_BAD_ARG, "%s is truncated", path);
        else if ((e = cudaMemcpy((uint8_t *)image + off, stage, nb, cudaMemcpyHostToDevice)) != cudaSuccess)
            st = fail(GDX_ERR_CUDA, "image upload failed: %s", cudaGetErrorString(e));
    }
    if (stage) cudaFreeHost(stage);
    if (st == GDX_OK) st = gdx_index_adopt_image(&h, image, device, 1, out);
    if (st != GDX_OK) cudaFree(image);
    return st;
}

// ================================================================================================
// replication
// ================================================================================================
extern "C" uint64_t gdx_index_header_bytes(void) { return sizeof(ImageHeader); }

extern "C" gdx_status gdx_index_export(const gdx_index *idx, void *header_out, const void **device_image,
                                       uint64_t *image_bytes) {
    if (!idx) return fail(GDX_ERR_BAD_ARG, "idx is NULL");
    if (header_out) memcpy(header_out, &idx->h, sizeof(ImageHeader));
    if (device_image) *device_image = idx->image;
    if (image_bytes) *image_bytes = idx->h.image_bytes;
    return GDX_OK;
}

extern "C" gdx_status gdx_index_adopt_image(const void *header, void *device_image, int32_t device_req,
                                            int32_t own_image, gdx_index **out) {
    if (!header || !device_image || !out) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    *out = nullptr;
    ImageHeader h;
    memcpy(&h, header, sizeof h);
    if (h.magic != kImageMagic || h.version != GDX_ABI_VERSION)
        return fail(GDX_ERR_BAD_ARG, "not a genedex_b200 image header (magic/version mismatch)");
    int device;
    GDX_TRY(resolve_device(device_req, &device));
    gdx_index *idx = new gdx_index();
    idx->h = h;
    idx->image = device_image;
    idx->own_image = own_image != 0;
    idx->device = device;
    idx->dev = make_dev_index(h, device_image);
    {
        DeviceGuard guard(device);
        gdx_status st = init_policies(idx);
        if (st != GDX_OK) {
            delete idx;
            return st;
        }
        auto_dense_sa(idx);
        auto_seed_table(idx);
    }
    *out = idx;
    return GDX_OK;
}

extern "C" gdx_status gdx_index_set_seed_table_depth(gdx_index *idx, int32_t depth) {
    if (!idx) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(idx->device);
    if (depth <= 0) {
        drop_seed_table(idx);
        return GDX_OK;
    }
    return build_seed_table(idx, (uint32_t)depth);
}

extern "C" gdx_status gdx_index_set_dense_suffix_array(gdx_index *idx, int32_t on) {
    if (!idx) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(idx->device);
    if (on) return build_dense_sa(idx);
    drop_dense_sa(idx);
    return GDX_OK;
}

extern "C" gdx_status gdx_index_replicate(const gdx_index *idx, const int32_t *devices, int32_t n_devices,
                                          gdx_index **out_replicas) {
    if (!idx || !devices || !out_replicas || n_devices < 0) return fail(GDX_ERR_BAD_ARG, "bad argument");
    for (int i = 0; i < n_devices; ++i) out_replicas[i] = nullptr;
    for (int i = 0; i < n_devices; ++i) {
        int device;
        GDX_TRY(resolve_device(devices[i], &device));
        DeviceGuard guard(device);
        void *img = nullptr;
        CUDA_TRY(cudaMalloc(&img, idx->h.image_bytes));
        cudaError_t e = cudaMemcpyPeer(img, device, idx->image, idx->device, idx->h.image_bytes);
        if (e != cudaSuccess) {
            cudaFree(img);
            return fail(GDX_ERR_CUDA, "peer copy to device %d failed: %s", device, cudaGetErrorString(e));
        }
        gdx_status st = gdx_index_adopt_image(&idx->h, img, device, 1, &out_replicas[i]);
        if (st != GDX_OK) {
            cudaFree(img);
            return st;
        }
    }
    return GDX_OK;
}

// ================================================================================================
// search
// ================================================================================================
namespace {

bool verify_enabled() {
    static const bool on = [] {
        const char *e = getenv("GDX_VERIFY");
        return !e || atoi(e) != 0;
    }();
    return on;
}

// mode 0: cursors (starts, ends); 1: counts; 2: locate intervals (interval or resolved hit)
template <class L>
void launch_search(const gdx_index *idx, const DevQueries &dq, uint64_t *a, uint64_t *b, int mode,
                   uint64_t qbase, uint64_t *err, unsigned long long *steps, const uint32_t *perm,
                   cudaStream_t stream) {
    if (dq.nq == 0) return;
    const unsigned grid = (unsigned)div_up(dq.nq, 256);
    const bool verify = idx->dev.text && verify_enabled();
    if (verify && mode != 0)
        k_search<L, true, false><<<grid, 256, 0, stream>>>(idx->dev, dq, a, b, mode, qbase, err, steps, perm);
    else if (verify && idx->dev.isa)
        k_search<L, true, true><<<grid, 256, 0, stream>>>(idx->dev, dq, a, b, mode, qbase, err, steps, perm);
    else
        k_search<L, false, false><<<grid, 256, 0, stream>>>(idx->dev, dq, a, b, mode, qbase, err, steps, perm);
}

// ---- suffix sort of a query batch (locality of the first search steps, see k_query_keys) -------------
constexpr uint64_t kSortMinQueries = 1ull << 15;

bool sort_enabled() {
    static const bool on = [] {
        const char *e = getenv("GDX_SORT_QUERIES");
        return !e || atoi(e) != 0;
    }();
    return on;
}

struct SortPlan {
    uint32_t key_bits = 0, key_syms = 0;
    size_t tmp_bytes = 0;
    uint64_t total_bytes = 0;  // keys a/b + idx a/b + cub temp, 256 B aligned pieces
    bool use = false;
    bool bucket = false;       // one-pass bucket grouping instead of the radix sort
};

bool bucket_mode() {
    static const bool on = [] {
        const char *e = getenv("GDX_SORT_MODE");
        return e && strcmp(e, "bucket") == 0;
    }();
    return on;
}

// fixed_len: length of every query of the batch, 0 = variable
SortPlan plan_sort(const gdx_index *idx, uint64_t nq, uint64_t fixed_len) {
    SortPlan p;
    if (!sort_enabled() || nq < kSortMinQueries || nq >= 0xffffffffull) return p;
    // The sort makes neighbouring threads share the first ~log_ns(nq) search steps.  When a lookup level
    // (configured, or the seed table accelerator) already replaces that many steps there is nothing left to
    // share: reading the queries in their own order is then cheaper (coalesced, no sort): 1.39 vs 2.01 ms per
    // 7.5 M queries at depth 13.
    {
        uint32_t d = idx->h.lookup_depth;
        if (idx->dev.seed_lookup && idx->dev.seed_depth > d && (fixed_len == 0 || fixed_len >= idx->dev.seed_depth))
            d = idx->dev.seed_depth;
        const uint64_t entries = seed_entries(idx->h.ns, d);
        if (d > 0 && (entries == 0 || entries >= nq)) return p;
    }
    uint32_t bits = 1;
    while ((1u << bits) < idx->h.ns) ++bits;
    p.key_bits = bits;
    if (bucket_mode() && bits <= kBucketBits) {
        p.key_syms = kBucketBits / bits;
        p.bucket = true;
        p.total_bytes = 2 * align_up(nq * 4, 256) + align_up(kNumBuckets * 4, 256);  // keys, perm, histogram
        p.use = true;
        return p;
    }
    p.key_syms = std::min<uint32_t>(32 / bits, 12);  // DNA: 24-bit keys = 3 radix passes, 4^12 > any batch
    cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
    if (cub::DeviceRadixSort::SortPairs(nullptr, p.tmp_bytes, dk, dv, (int64_t)nq, 0, (int)(p.key_bits * p.key_syms)) !=
        cudaSuccess)
        return p;
    p.total_bytes = 4 * align_up(nq * 4, 256) + align_up(p.tmp_bytes, 256);
    p.use = true;
    return p;
}

// scratch: total_bytes of device memory; returns the permutation (sorted query indices) in *perm
gdx_status sort_queries(const gdx_index *idx, const DevQueries &dq, const SortPlan &p, void *scratch,
                        cudaStream_t stream, const uint32_t **perm) {
    uint8_t *base = (uint8_t *)scratch;
    const uint64_t piece = align_up(dq.nq * 4, 256);
    if (p.bucket) {
        uint32_t *keys = (uint32_t *)base, *out = (uint32_t *)(base + piece), *hist = (uint32_t *)(base + 2 * piece);
        const unsigned grid = (unsigned)div_up(dq.nq, 256);
        CUDA_TRY(cudaMemsetAsync(hist, 0, kNumBuckets * 4, stream));
        k_bucket_count<<<grid, 256, 0, stream>>>(idx->dev, dq, p.key_bits, p.key_syms, keys, hist);
        k_bucket_scan<<<1, 1024, 0, stream>>>(hist);
        k_bucket_scatter<<<grid, 256, 0, stream>>>(keys, dq.nq, hist, out);
        CUDA_TRY(cudaGetLastError());
        *perm = out;
        return GDX_OK;
    }
    uint32_t *ka = (uint32_t *)base, *kb = (uint32_t *)(base + piece), *ia = (uint32_t *)(base + 2 * piece),
             *ib = (uint32_t *)(base + 3 * piece);
    void *tmp = base + 4 * piece;
    k_query_keys<<<(unsigned)div_up(dq.nq, 256), 256, 0, stream>>>(idx->dev, dq, p.key_bits, p.key_syms, ka, ia);
    CUDA_TRY(cudaGetLastError());
    cub::DoubleBuffer<uint32_t> dk(ka, kb), dv(ia, ib);
    size_t tb = p.tmp_bytes;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (int64_t)dq.nq, 0, (int)(p.key_bits * p.key_syms), stream));
    *perm = dv.Current();
    return GDX_OK;
}

gdx_status check_queries(const gdx_queries *q) {
    if (!q) return fail(GDX_ERR_BAD_ARG, "queries is NULL");
    if (q->offsets && q->nq && q->offsets[q->nq] < q->offsets[0])
        return fail(GDX_ERR_BAD_ARG, "queries->offsets must be non-decreasing");
    if (q->nq && !q->offsets && q->fixed_len && !q->bytes) return fail(GDX_ERR_BAD_ARG, "queries->bytes is NULL");
    return GDX_OK;
}

uint64_t query_bytes_end(const gdx_queries *q, uint64_t i) {
    return q->offsets ? q->offsets[i] : i * q->fixed_len;
}

gdx_status acquire_pinned_hits(const gdx_index *idx, uint64_t bytes, void **out, uint64_t *cap_out);
void release_pinned_hits(const gdx_index *idx, void *p);

// State of a pipelined gdx_locate_many: every chunk of the search pipeline continues, on its own
// stream, with counts -> scan -> expand -> walk -> D2H of its hits, so that the locate kernels and the
// hit copies of chunk k overlap the query upload of the later chunks.
struct LocatePipe {
    const gdx_index *idx;
    Workspace *ws;
    uint64_t *hit_offsets;  // caller's array, nq + 1 entries
    void *pinned = nullptr; // library-owned pinned hit buffer (grows)
    uint64_t pinned_cap = 0;
    uint64_t total = 0;     // hits of all finished chunks
    struct Pending {
        int slot;
        uint64_t q0, cq;
        bool valid = false;
    } pending, pending_offsets;
    bool stage_offsets = false;  // the caller's hit_offsets array is pageable

    // hand the previous chunk's CSR offsets from the slot's pinned staging to the caller
    gdx_status flush_offsets() {
        if (!pending_offsets.valid) return GDX_OK;
        pending_offsets.valid = false;
        Slot &ps = ws->slot[pending_offsets.slot];
        CUDA_TRY(cudaEventSynchronize(ps.ev_out));
        HostPool::get().copy(hit_offsets + pending_offsets.q0, ps.h_out_a.p, pending_offsets.cq * 8);
        return GDX_OK;
    }

    // after k_search(mode 2) of a chunk: widths, exclusive scan, totals to the host
    gdx_status stage_counts(Slot &sl, int slot, uint64_t q0, uint64_t cq) {
        CUDA_TRY(sl.counts.reserve((cq + 1) * 8));
        CUDA_TRY(sl.local_off.reserve((cq + 1) * 8));
        CUDA_TRY(cudaMemsetAsync(sl.d_words, 0, 4 * sizeof(uint64_t), sl.stream));
        CUDA_TRY(cudaMemsetAsync(sl.counts.as<uint64_t>() + cq, 0, 8, sl.stream));
        k_interval_counts<<<(unsigned)div_up(cq, 256), 256, 0, sl.stream>>>(
            sl.out_a.as<uint64_t>(), sl.out_b.as<uint64_t>(), cq, sl.counts.as<uint64_t>(),
            reinterpret_cast<unsigned long long *>(sl.d_words));
        size_t tmp = 0;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp, sl.counts.as<uint64_t>(), sl.local_off.as<uint64_t>(),
                                               cq + 1, sl.stream));
        CUDA_TRY(sl.scan_tmp.reserve(tmp));
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(sl.scan_tmp.p, tmp, sl.counts.as<uint64_t>(),
                                               sl.local_off.as<uint64_t>(), cq + 1, sl.stream));
        CUDA_TRY(cudaMemcpyAsync(sl.h_words, sl.local_off.as<uint64_t>() + cq, 8, cudaMemcpyDeviceToHost, sl.stream));
        CUDA_TRY(cudaMemcpyAsync(sl.h_words + 1, sl.d_words, 8, cudaMemcpyDeviceToHost, sl.stream));
        CUDA_TRY(cudaEventRecord(sl.ev_total, sl.stream));
        pending = Pending{slot, q0, cq, true};
        t_stats.kernel_launches += 2;
        return GDX_OK;
    }

    Pending take_pending() {
        Pending p = pending;
        pending.valid = false;
        return p;
    }

    // once the chunk's number of hits is known: expand, walk, copy hits + global offsets out
    gdx_status finish(const Pending &p) {
        if (!p.valid) return GDX_OK;
        Slot &sl = ws->slot[p.slot];
        const uint64_t cq = p.cq, q0 = p.q0;
        CUDA_TRY(cudaEventSynchronize(sl.ev_total));
        const uint64_t n_hits = sl.h_words[0], nbig = sl.h_words[1], base = total;
        if (n_hits) {
            if ((base + n_hits) * sizeof(gdx_hit) > pinned_cap) {  // grow the pinned result buffer (rare)
                for (int s2 = 0; s2 < kSlots; ++s2) CUDA_TRY(cudaStreamSynchronize(ws->slot[s2].stream));
                void *bigger = nullptr;
                uint64_t cap = 0;
                GDX_TRY(acquire_pinned_hits(idx, 2 * (base + n_hits) * sizeof(gdx_hit), &bigger, &cap));
                if (pinned) {
                    memcpy(bigger, pinned, base * sizeof(gdx_hit));
                    release_pinned_hits(idx, pinned);
                }
                pinned = bigger;
                pinned_cap = cap;
            }
            size_t free_b = 0, total_b = 0;
            if (n_hits * 24 > sl.rows.cap + sl.hits.cap) {
                CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
                if (n_hits * 24 > free_b + sl.rows.cap + sl.hits.cap)
                    return fail(GDX_ERR_OOM, "%llu hits of one chunk do not fit into device memory; split the batch",
                                (unsigned long long)n_hits);
            }
            CUDA_TRY(sl.rows.reserve(n_hits * 8));
            CUDA_TRY(sl.hits.reserve(n_hits * 16));
            CUDA_TRY(sl.big.reserve((nbig + 1) * 8));
            k_expand_rows<<<(unsigned)div_up(cq, 256), 256, 0, sl.stream>>>(
                sl.out_a.as<uint64_t>(), sl.out_b.as<uint64_t>(), sl.local_off.as<uint64_t>(), cq,
                sl.rows.as<uint64_t>(), sl.big.as<uint64_t>(), reinterpret_cast<unsigned long long *>(sl.d_words + 1));
            if (nbig) {
                dim3 grid((unsigned)nbig, 32);
                k_expand_big_rows<<<grid, 256, 0, sl.stream>>>(sl.out_a.as<uint64_t>(), sl.out_b.as<uint64_t>(),
                                                               sl.local_off.as<uint64_t>(), sl.big.as<uint64_t>(),
                                                               sl.rows.as<uint64_t>());
            }
            unsigned long long *d_walk = reinterpret_cast<unsigned long long *>(ws->small.d + 5);
            GDX_TRY(dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
                k_locate_walk<decltype(L)><<<(unsigned)div_up(n_hits, 256), 256, 0, sl.stream>>>(
                    idx->dev, sl.rows.as<uint64_t>(), n_hits, sl.hits.as<ulonglong2>(), d_walk);
                return GDX_OK;
            }));
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync((gdx_hit *)pinned + base, sl.hits.p, n_hits * sizeof(gdx_hit),
                                     cudaMemcpyDeviceToHost, sl.stream));
            t_stats.kernel_launches += 2 + (nbig ? 1 : 0);
        }
        k_add_base<<<(unsigned)div_up(cq, 256), 256, 0, sl.stream>>>(sl.local_off.as<uint64_t>(), cq, base);
        CUDA_TRY(cudaGetLastError());
        if (stage_offsets) {
            GDX_TRY(flush_offsets());  // frees the staging of the chunk before
            CUDA_TRY(sl.h_out_a.reserve(cq * 8));
            CUDA_TRY(cudaMemcpyAsync(sl.h_out_a.p, sl.local_off.p, cq * 8, cudaMemcpyDeviceToHost, sl.stream));
            CUDA_TRY(cudaEventRecord(sl.ev_out, sl.stream));
            pending_offsets = Pending{p.slot, q0, cq, true};
        } else {
            CUDA_TRY(cudaMemcpyAsync(hit_offsets + q0, sl.local_off.p, cq * 8, cudaMemcpyDeviceToHost, sl.stream));
        }
        t_stats.kernel_launches += 1;
        total += n_hits;
        return GDX_OK;
    }
};

// Chunked pipeline over kSlots streams: H2D(query bytes) -> k_search -> D2H(results) per chunk.
// If dev_a/dev_b are given the results stay on the device (locate path) and nothing is copied back.
gdx_status search_host(const gdx_index *idx, Workspace *ws, const gdx_queries *qs, uint64_t *out_a,
                       uint64_t *out_b, int mode, uint64_t *dev_a, uint64_t *dev_b, LocatePipe *lp = nullptr) {
    const uint64_t nq = qs->nq;
    for (int s = 0; s < kSlots; ++s) ws->slot[s].ev_used = 0;
    CUDA_TRY(cudaMemset(ws->small.d, 0xff, 4 * sizeof(uint64_t)));
    CUDA_TRY(cudaMemset(ws->small.d + 4, 0, 12 * sizeof(uint64_t)));
    unsigned long long *d_steps = reinterpret_cast<unsigned long long *>(ws->small.d + 4);
    static const bool trace = getenv("GDX_TRACE") && atoi(getenv("GDX_TRACE")) != 0;
    struct TraceRow {
        int k;
        uint64_t nq, bytes;
        cudaEvent_t h2d_begin, k_begin, k_end, d2h_end;
    };
    std::vector<TraceRow> trace_rows;
    const auto t_host0 = std::chrono::steady_clock::now();
    const uint64_t byte_end = query_bytes_end(qs, nq);
    // Ordinary (pageable) caller memory would make every cudaMemcpyAsync a blocking, driver-staged copy at
    // 8-14 GB/s.  Large batches are staged through the slots' own pinned buffers instead, copied by a few
    // host threads (HostPool); pinned caller buffers take the direct path.
    const uint64_t total_in = byte_end - query_bytes_end(qs, 0);
    const bool stage_in = nq && total_in >= kStageMinBytes && !is_pinned(qs->bytes);
    const bool stage_off = nq && qs->offsets && nq * 8 >= kStageMinBytes && !is_pinned(qs->offsets);
    const bool stage_out = nq && !dev_a && !lp && nq * 8 >= kStageMinBytes && !is_pinned(out_a);
    struct PendingOut {
        int slot = 0;
        uint64_t q0 = 0, cq = 0;
        bool valid = false;
    } pend_out;
    auto finish_out = [&](PendingOut &p) -> gdx_status {
        if (!p.valid) return GDX_OK;
        p.valid = false;
        Slot &ps = ws->slot[p.slot];
        CUDA_TRY(cudaEventSynchronize(ps.ev_out));
        HostPool::get().copy(out_a + p.q0, ps.h_out_a.p, p.cq * 8);
        if (mode == 0) HostPool::get().copy(out_b + p.q0, ps.h_out_b.p, p.cq * 8);
        return GDX_OK;
    };
    for (int s2 = 0; s2 < kSlots; ++s2) ws->slot[s2].h2d_pending = false;
    uint64_t budget = kChunkFirst;
    uint64_t q0 = 0;
    int k = 0;
    while (q0 < nq) {
        // chunk [q0, q1): about `budget` query bytes, at least one query
        const uint64_t remaining = byte_end - query_bytes_end(qs, q0);
        uint64_t want = budget;
        if (remaining <= budget) want = remaining > 2 * kChunkTail ? remaining - kChunkTail : remaining;
        uint64_t q1;
        if (qs->offsets) {
            const uint64_t *b = qs->offsets + q0 + 1, *e = qs->offsets + nq + 1;
            const uint64_t *it = std::upper_bound(b, e, qs->offsets[q0] + want);
            q1 = q0 + (uint64_t)(it - b);
            if (q1 == q0) q1 = q0 + 1;
        } else {
            q1 = q0 + (qs->fixed_len ? std::max<uint64_t>(1, want / qs->fixed_len) : kChunkMaxQueries);
        }
        q1 = std::min<uint64_t>(q1, std::min<uint64_t>(nq, q0 + kChunkMaxQueries));
        budget = std::min<uint64_t>(budget * 2, kChunkMax);
        const uint64_t cq = q1 - q0;
        const uint64_t byte0 = query_bytes_end(qs, q0), byte1 = query_bytes_end(qs, q1);
        Slot &sl = ws->slot[k % kSlots];
        // growing a slot buffer frees the old one: only safe once the slot's stream has drained
        const SortPlan sp = plan_sort(idx, cq, qs->offsets ? 0 : qs->fixed_len);
        if (sl.bytes.cap < byte1 - byte0 + 16 || (qs->offsets && sl.offsets.cap < (cq + 1) * 8) ||
            (!dev_a && (sl.out_a.cap < cq * 8 || ((mode == 0 || lp) && sl.out_b.cap < cq * 8))) ||
            (sp.use && sl.sort.cap < sp.total_bytes))
            CUDA_TRY(cudaStreamSynchronize(sl.stream));
        CUDA_TRY(sl.bytes.reserve(byte1 - byte0 + 16));
        cudaEvent_t ev_h2d = nullptr;
        if (trace) {
            cudaEventCreate(&ev_h2d);
            cudaEventRecord(ev_h2d, sl.stream);
        }
        if ((stage_in || stage_off) && sl.h2d_pending) {  // the slot's staging buffers are free again
            CUDA_TRY(cudaEventSynchronize(sl.ev_h2d));
            sl.h2d_pending = false;
        }
        if (byte1 > byte0) {
            const uint8_t *src = qs->bytes + byte0;
            if (stage_in) {
                CUDA_TRY(sl.h_in.reserve(byte1 - byte0));
                HostPool::get().copy(sl.h_in.p, src, byte1 - byte0);
                src = (const uint8_t *)sl.h_in.p;
            }
            CUDA_TRY(cudaMemcpyAsync(sl.bytes.p, src, byte1 - byte0, cudaMemcpyHostToDevice, sl.stream));
        }
        DevQueries dq;
        dq.bytes = sl.bytes.as<uint8_t>();
        dq.offsets = nullptr;
        dq.fixed_len = qs->fixed_len;
        dq.nq = cq;
        dq.base = 0;
        if (qs->offsets) {
            CUDA_TRY(sl.offsets.reserve((cq + 1) * 8));
            const uint64_t *osrc = qs->offsets + q0;
            if (stage_off) {
                CUDA_TRY(sl.h_off.reserve((cq + 1) * 8));
                HostPool::get().copy(sl.h_off.p, osrc, (cq + 1) * 8);
                osrc = (const uint64_t *)sl.h_off.p;
            }
            CUDA_TRY(cudaMemcpyAsync(sl.offsets.p, osrc, (cq + 1) * 8, cudaMemcpyHostToDevice, sl.stream));
            dq.offsets = sl.offsets.as<uint64_t>();
            dq.base = byte0;
        }
        uint64_t *a, *b;
        if (dev_a) {
            a = dev_a + q0;
            b = dev_b ? dev_b + q0 : nullptr;
        } else {
            CUDA_TRY(sl.out_a.reserve(cq * 8));
            a = sl.out_a.as<uint64_t>();
            b = nullptr;
            if (mode == 0 || lp) {
                CUDA_TRY(sl.out_b.reserve(cq * 8));
                b = sl.out_b.as<uint64_t>();
            }
        }
        if (stage_in || stage_off) {
            CUDA_TRY(cudaEventRecord(sl.ev_h2d, sl.stream));
            sl.h2d_pending = true;
        }
        cudaEvent_t e0 = ws->next_event(sl), e1 = ws->next_event(sl);
        CUDA_TRY(cudaEventRecord(e0, sl.stream));
        if (trace) trace_rows.push_back(TraceRow{k, cq, byte1 - byte0, ev_h2d, e0, e1, nullptr});
        const uint32_t *perm = nullptr;
        if (sp.use) {
            CUDA_TRY(sl.sort.reserve(sp.total_bytes));
            GDX_TRY(sort_queries(idx, dq, sp, sl.sort.p, sl.stream, &perm));
            t_stats.kernel_launches += 2;
        }
        gdx_status st = dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
            launch_search<decltype(L)>(idx, dq, a, b, mode, q0, ws->small.d + (k % kSlots), d_steps, perm, sl.stream);
            return GDX_OK;
        });
        GDX_TRY(st);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(e1, sl.stream));
        if (lp) {  // locate continues per chunk; the previous chunk is finished while this one runs
            const LocatePipe::Pending prev = lp->take_pending();
            GDX_TRY(lp->stage_counts(sl, k % kSlots, q0, cq));
            GDX_TRY(lp->finish(prev));
        } else if (stage_out) {  // results go to pinned staging; the previous chunk's are handed over meanwhile
            PendingOut prev = pend_out;
            CUDA_TRY(sl.h_out_a.reserve(cq * 8));
            CUDA_TRY(cudaMemcpyAsync(sl.h_out_a.p, a, cq * 8, cudaMemcpyDeviceToHost, sl.stream));
            if (mode == 0) {
                CUDA_TRY(sl.h_out_b.reserve(cq * 8));
                CUDA_TRY(cudaMemcpyAsync(sl.h_out_b.p, b, cq * 8, cudaMemcpyDeviceToHost, sl.stream));
            }
            CUDA_TRY(cudaEventRecord(sl.ev_out, sl.stream));
            pend_out = PendingOut{k % kSlots, q0, cq, true};
            GDX_TRY(finish_out(prev));
        } else if (!dev_a) {
            CUDA_TRY(cudaMemcpyAsync(out_a + q0, a, cq * 8, cudaMemcpyDeviceToHost, sl.stream));
            if (mode == 0) CUDA_TRY(cudaMemcpyAsync(out_b + q0, b, cq * 8, cudaMemcpyDeviceToHost, sl.stream));
        }
        if (trace) {
            cudaEventCreate(&trace_rows.back().d2h_end);
            cudaEventRecord(trace_rows.back().d2h_end, sl.stream);
        }
        t_stats.kernel_launches += 1;
        q0 = q1;
        ++k;
    }
    if (lp) GDX_TRY(lp->finish(lp->take_pending()));
    if (lp) GDX_TRY(lp->flush_offsets());
    GDX_TRY(finish_out(pend_out));
    const double t_issue = trace ? std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count() : 0;
    for (int s = 0; s < kSlots; ++s) CUDA_TRY(cudaStreamSynchronize(ws->slot[s].stream));
    if (trace && !trace_rows.empty()) {
        const double t_sync = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
        fprintf(stderr, "[gdx trace] host: all chunks issued at %.3f ms, streams drained at %.3f ms\n", t_issue, t_sync);
        cudaEvent_t base = trace_rows[0].h2d_begin;
        for (auto &r : trace_rows) {
            float a = 0, b = 0, c = 0, d = 0;
            cudaError_t te = cudaEventElapsedTime(&a, base, r.h2d_begin);
            if (te != cudaSuccess) fprintf(stderr, "[gdx trace] elapsed: %s\n", cudaGetErrorString(te));
            cudaEventElapsedTime(&b, base, r.k_begin);
            cudaEventElapsedTime(&c, base, r.k_end);
            cudaEventElapsedTime(&d, base, r.d2h_end);
            fprintf(stderr, "[gdx trace] chunk %2d slot %d nq %8llu bytes %10llu | h2d %.3f..%.3f | kernels %.3f..%.3f | d2h ..%.3f\n",
                    r.k, r.k % kSlots, (unsigned long long)r.nq, (unsigned long long)r.bytes, a, b, b, c, d);
        }
        for (auto &r : trace_rows) {
            cudaEventDestroy(r.h2d_begin);
            cudaEventDestroy(r.d2h_end);
        }
    }
    CUDA_TRY(cudaMemcpy(ws->small.h, ws->small.d, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    double ms = 0;
    for (int s = 0; s < kSlots; ++s)
        for (size_t i = 0; i + 1 < ws->slot[s].ev_used; i += 2) {
            float t = 0;
            cudaEventElapsedTime(&t, ws->slot[s].ev[i], ws->slot[s].ev[i + 1]);
            ms += t;
        }
    t_stats.queries = nq;
    t_stats.lf_steps = ws->small.h[4];
    t_stats.walk_steps = ws->small.h[5];
    t_stats.verified_queries = ws->small.h[9];
    t_stats.kernel_ms_search = ms;
    uint64_t bad = kNoError;
    for (int s = 0; s < kSlots; ++s) bad = std::min(bad, ws->small.h[s]);
    if (bad != kNoError) {
        t_error_query = bad;
        return fail(GDX_ERR_INVALID_SYMBOL,
                    "query %llu: symbol in io representation should be valid (alphabet.rs:195-198)",
                    (unsigned long long)bad);
    }
    return GDX_OK;
}

gdx_status begin_call(const gdx_index *idx, const char *what) {
    if (!idx) return fail(GDX_ERR_BAD_ARG, "%s: idx is NULL", what);
    t_stats = gdx_stats{};
    return GDX_OK;
}

// CSR + expand + walk on device-resident intervals; results in ws->hit_offsets (n+1) and ws->hits.
gdx_status locate_device_intervals(const gdx_index *idx, Workspace *ws, const uint64_t *d_starts,
                                   const uint64_t *d_ends, uint64_t n, uint64_t *total_out) {
    cudaStream_t st = ws->slot[0].stream;
    CUDA_TRY(ws->counts.reserve((n + 1) * 8));
    CUDA_TRY(ws->hit_offsets.reserve((n + 1) * 8));
    CUDA_TRY(cudaMemsetAsync(ws->small.d + 4, 0, 12 * sizeof(uint64_t), st));
    CUDA_TRY(cudaMemsetAsync(ws->counts.as<uint64_t>() + n, 0, 8, st));
    unsigned long long *d_big = reinterpret_cast<unsigned long long *>(ws->small.d + 6);
    CUDA_TRY(cudaEventRecord(ws->ev_a, st));
    if (n) k_interval_counts<<<(unsigned)div_up(n, 256), 256, 0, st>>>(d_starts, d_ends, n, ws->counts.as<uint64_t>(), d_big);
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ws->counts.as<uint64_t>(),
                                           ws->hit_offsets.as<uint64_t>(), n + 1, st));
    CUDA_TRY(ws->scan_tmp.reserve(tmp_bytes));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(ws->scan_tmp.p, tmp_bytes, ws->counts.as<uint64_t>(),
                                           ws->hit_offsets.as<uint64_t>(), n + 1, st));
    CUDA_TRY(cudaMemcpyAsync(ws->small.h + 8, ws->hit_offsets.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ws->small.h + 6, d_big, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const uint64_t total = ws->small.h[8], nbig = ws->small.h[6];
    *total_out = total;
    t_stats.kernel_launches += 2;
    if (total) {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t need = (total > ws->rows.cap / 8 ? total * 8 : 0) + (total > ws->hits.cap / 16 ? total * 16 : 0);
        if (need > free_b + ws->rows.cap + ws->hits.cap)
            return fail(GDX_ERR_OOM, "%llu hits do not fit into device memory; split the batch",
                        (unsigned long long)total);
        CUDA_TRY(ws->rows.reserve(total * 8));
        CUDA_TRY(ws->hits.reserve(total * 16));
        CUDA_TRY(ws->big_list.reserve((nbig + 1) * 8));
        unsigned long long *d_cursor = reinterpret_cast<unsigned long long *>(ws->small.d + 7);
        k_expand_rows<<<(unsigned)div_up(n, 256), 256, 0, st>>>(d_starts, d_ends, ws->hit_offsets.as<uint64_t>(), n,
                                                               ws->rows.as<uint64_t>(), ws->big_list.as<uint64_t>(),
                                                               d_cursor);
        if (nbig) {
            dim3 grid((unsigned)nbig, 32);
            k_expand_big_rows<<<grid, 256, 0, st>>>(d_starts, d_ends, ws->hit_offsets.as<uint64_t>(),
                                                    ws->big_list.as<uint64_t>(), ws->rows.as<uint64_t>());
            t_stats.kernel_launches += 1;
        }
        unsigned long long *d_walk = reinterpret_cast<unsigned long long *>(ws->small.d + 5);
        gdx_status s2 = dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
            k_locate_walk<decltype(L)><<<(unsigned)div_up(total, 256), 256, 0, st>>>(
                idx->dev, ws->rows.as<uint64_t>(), total, ws->hits.as<ulonglong2>(), d_walk);
            return GDX_OK;
        });
        GDX_TRY(s2);
        CUDA_TRY(cudaGetLastError());
        t_stats.kernel_launches += 2;
    }
    CUDA_TRY(cudaEventRecord(ws->ev_b, st));
    return GDX_OK;
}

void release_pinned_hits(const gdx_index *idx, void *p) {
    std::lock_guard<std::mutex> lk(idx->mu);
    for (auto &b : idx->pinned)
        if (b.p == p) b.in_use = false;
}

gdx_status acquire_pinned_hits(const gdx_index *idx, uint64_t bytes, void **out, uint64_t *cap_out = nullptr) {
    std::lock_guard<std::mutex> lk(idx->mu);
    PinnedHits *best = nullptr;
    for (auto &p : idx->pinned)
        if (!p.in_use && p.cap >= bytes && (!best || p.cap < best->cap)) best = &p;
    if (!best) {
        for (auto &p : idx->pinned)  // recycle the largest free but too small buffer
            if (!p.in_use && (!best || p.cap > best->cap)) best = &p;
        if (best) {
            cudaFreeHost(best->p);
            best->p = nullptr;
            best->cap = 0;
        } else {
            idx->pinned.push_back(PinnedHits{});
            best = &idx->pinned.back();
        }
        const uint64_t want = align_up(bytes + bytes / 4 + 4096, 4096);
        cudaError_t e = cudaMallocHost(&best->p, want);
        if (e != cudaSuccess) {
            best->p = nullptr;
            return fail(GDX_ERR_OOM, "pinned host allocation of %llu bytes failed: %s", (unsigned long long)want,
                        cudaGetErrorString(e));
        }
        best->cap = want;
    }
    best->in_use = true;
    *out = best->p;
    if (cap_out) *cap_out = best->cap;
    return GDX_OK;
}

gdx_status finish_locate(const gdx_index *idx, Workspace *ws, uint64_t n, uint64_t total, uint64_t *hit_offsets,
                         gdx_hit **hits, uint64_t *num_hits) {
    cudaStream_t st = ws->slot[0].stream;
    void *h = nullptr;
    GDX_TRY(acquire_pinned_hits(idx, total * sizeof(gdx_hit), &h));
    if (total) CUDA_TRY(cudaMemcpyAsync(h, ws->hits.p, total * sizeof(gdx_hit), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(hit_offsets, ws->hit_offsets.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ws->small.h, ws->small.d, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ws->ev_a, ws->ev_b);
    t_stats.kernel_ms_locate = ms;
    t_stats.hits = total;
    t_stats.walk_steps += ws->small.h[5];
    *hits = (gdx_hit *)h;
    *num_hits = total;
    return GDX_OK;
}

}  // namespace

extern "C" gdx_status gdx_cursors_many(const gdx_index *idx, const gdx_queries *queries, uint64_t *starts,
                                       uint64_t *ends) {
    GDX_TRY(begin_call(idx, "gdx_cursors_many"));
    GDX_TRY(check_queries(queries));
    if (queries->nq && (!starts || !ends)) return fail(GDX_ERR_BAD_ARG, "output is NULL");
    DeviceGuard guard(idx->device);
    WsLease lease(idx);
    if (!lease.w) return fail(GDX_ERR_CUDA, "could not create a CUDA workspace: %s", cudaGetErrorString(cudaGetLastError()));
    return search_host(idx, lease.w, queries, starts, ends, 0, nullptr, nullptr);
}

extern "C" gdx_status gdx_count_many(const gdx_index *idx, const gdx_queries *queries, uint64_t *counts) {
    GDX_TRY(begin_call(idx, "gdx_count_many"));
    GDX_TRY(check_queries(queries));
    if (queries->nq && !counts) return fail(GDX_ERR_BAD_ARG, "output is NULL");
    DeviceGuard guard(idx->device);
    WsLease lease(idx);
    if (!lease.w) return fail(GDX_ERR_CUDA, "could not create a CUDA workspace: %s", cudaGetErrorString(cudaGetLastError()));
    return search_host(idx, lease.w, queries, counts, nullptr, 1, nullptr, nullptr);
}

extern "C" gdx_status gdx_locate_many(const gdx_index *idx, const gdx_queries *queries, uint64_t *hit_offsets,
                                      gdx_hit **hits, uint64_t *num_hits) {
    GDX_TRY(begin_call(idx, "gdx_locate_many"));
    GDX_TRY(check_queries(queries));
    if (!hit_offsets || !hits || !num_hits) return fail(GDX_ERR_BAD_ARG, "output is NULL");
    *hits = nullptr;
    *num_hits = 0;
    DeviceGuard guard(idx->device);
    WsLease lease(idx);
    Workspace *ws = lease.w;
    if (!ws) return fail(GDX_ERR_CUDA, "could not create a CUDA workspace: %s", cudaGetErrorString(cudaGetLastError()));
    const uint64_t n = queries->nq;
    static const bool pipelined = !(getenv("GDX_LOCATE_PIPELINE") && atoi(getenv("GDX_LOCATE_PIPELINE")) == 0);
    if (pipelined) {
        // every chunk of the search pipeline carries on with counts -> scan -> expand -> walk -> D2H
        LocatePipe lp{idx, ws, hit_offsets};
        lp.stage_offsets = n * 8 >= kStageMinBytes && !is_pinned(hit_offsets);
        GDX_TRY(acquire_pinned_hits(idx, std::max<uint64_t>(n, 4096) * sizeof(gdx_hit), &lp.pinned, &lp.pinned_cap));
        gdx_status st = search_host(idx, ws, queries, nullptr, nullptr, 2, nullptr, nullptr, &lp);
        if (st != GDX_OK) {
            for (int s2 = 0; s2 < kSlots; ++s2) cudaStreamSynchronize(ws->slot[s2].stream);
            release_pinned_hits(idx, lp.pinned);
            return st;
        }
        hit_offsets[n] = lp.total;
        t_stats.hits = lp.total;
        *hits = (gdx_hit *)lp.pinned;
        *num_hits = lp.total;
        return GDX_OK;
    }
    // a buffer may only be replaced once nothing in flight uses it: every call ends synchronized
    CUDA_TRY(ws->starts.reserve((n + 1) * 8));
    CUDA_TRY(ws->ends.reserve((n + 1) * 8));
    GDX_TRY(search_host(idx, ws, queries, nullptr, nullptr, 2, ws->starts.as<uint64_t>(), ws->ends.as<uint64_t>()));
    uint64_t total = 0;
    GDX_TRY(locate_device_intervals(idx, ws, ws->starts.as<uint64_t>(), ws->ends.as<uint64_t>(), n, &total));
    return finish_locate(idx, ws, n, total, hit_offsets, hits, num_hits);
}

extern "C" gdx_status gdx_locate_intervals(const gdx_index *idx, const uint64_t *starts, const uint64_t *ends,
                                           uint64_t n, uint64_t *hit_offsets, gdx_hit **hits, uint64_t *num_hits) {
    GDX_TRY(begin_call(idx, "gdx_locate_intervals"));
    if (!hit_offsets || !hits || !num_hits || (n && (!starts || !ends))) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    *hits = nullptr;
    *num_hits = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (starts[i] > ends[i] || ends[i] > idx->h.n)
            return fail(GDX_ERR_BAD_ARG, "interval %llu is not inside [0, text_len]", (unsigned long long)i);
    DeviceGuard guard(idx->device);
    WsLease lease(idx);
    Workspace *ws = lease.w;
    if (!ws) return fail(GDX_ERR_CUDA, "could not create a CUDA workspace: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStream_t st = ws->slot[0].stream;
    CUDA_TRY(ws->starts.reserve((n + 1) * 8));
    CUDA_TRY(ws->ends.reserve((n + 1) * 8));
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(ws->starts.p, starts, n * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ws->ends.p, ends, n * 8, cudaMemcpyHostToDevice, st));
    }
    uint64_t total = 0;
    GDX_TRY(locate_device_intervals(idx, ws, ws->starts.as<uint64_t>(), ws->ends.as<uint64_t>(), n, &total));
    return finish_locate(idx, ws, n, total, hit_offsets, hits, num_hits);
}

extern "C" void gdx_free_hits(const gdx_index *idx, gdx_hit *hits) {
    if (!idx || !hits) return;
    std::lock_guard<std::mutex> lk(idx->mu);
    for (auto &p : idx->pinned)
        if (p.p == hits) p.in_use = false;
}

extern "C" gdx_status gdx_extend_many(const gdx_index *idx, uint64_t *starts, uint64_t *ends,
                                      const uint8_t *io_symbols, uint64_t n) {
    GDX_TRY(begin_call(idx, "gdx_extend_many"));
    if (n == 0) return GDX_OK;
    if (!starts || !ends || !io_symbols) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(idx->device);
    WsLease lease(idx);
    Workspace *ws = lease.w;
    if (!ws) return fail(GDX_ERR_CUDA, "could not create a CUDA workspace: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStream_t st = ws->slot[0].stream;
    CUDA_TRY(ws->starts.reserve(n * 8));
    CUDA_TRY(ws->ends.reserve(n * 8));
    CUDA_TRY(ws->symbols.reserve(n));
    CUDA_TRY(cudaMemsetAsync(ws->small.d, 0xff, 16, st));  // [0] invalid symbol, [1] cursor out of bounds
    CUDA_TRY(cudaStreamSynchronize(st));
    // large pageable cursor arrays go through pinned staging (see search_host)
    Slot &sl0 = ws->slot[0];
    const bool stage = n * 8 >= kStageMinBytes && !(is_pinned(starts) && is_pinned(ends));
    const uint64_t *src_s = starts, *src_e = ends;
    if (stage) {
        CUDA_TRY(sl0.h_out_a.reserve(n * 8));
        CUDA_TRY(sl0.h_out_b.reserve(n * 8));
        HostPool::get().copy(sl0.h_out_a.p, starts, n * 8);
        HostPool::get().copy(sl0.h_out_b.p, ends, n * 8);
        src_s = (const uint64_t *)sl0.h_out_a.p;
        src_e = (const uint64_t *)sl0.h_out_b.p;
    }
    uint64_t *d_s = ws->starts.as<uint64_t>(), *d_e = ws->ends.as<uint64_t>();
    uint8_t *d_c = ws->symbols.as<uint8_t>();
    // chunks of 1 M cursors round-robin over the workspace streams: upload, kernel and download of
    // neighbouring chunks overlap (PCIe is full duplex).  Results land in the staging buffers (pageable
    // caller arrays: handed over only on success) or directly in the caller's pinned arrays (on error their
    // contents are unspecified -- the reference panics in that case).
    const uint64_t kChunk = 1ull << 20;
    uint64_t launches = 0;
    for (uint64_t off = 0, k = 0; off < n; off += kChunk, ++k) {
        const uint64_t cn = std::min<uint64_t>(kChunk, n - off);
        cudaStream_t cs = ws->slot[k % kSlots].stream;
        CUDA_TRY(cudaMemcpyAsync(d_s + off, src_s + off, cn * 8, cudaMemcpyHostToDevice, cs));
        CUDA_TRY(cudaMemcpyAsync(d_e + off, src_e + off, cn * 8, cudaMemcpyHostToDevice, cs));
        CUDA_TRY(cudaMemcpyAsync(d_c + off, io_symbols + off, cn, cudaMemcpyHostToDevice, cs));
        GDX_TRY(dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
            k_extend<decltype(L)><<<(unsigned)div_up(cn, 256), 256, 0, cs>>>(idx->dev, d_s + off, d_e + off, d_c + off, cn,
                                                                           ws->small.d, off);
            return GDX_OK;
        }));
        CUDA_TRY(cudaGetLastError());
        ++launches;
        uint64_t *dst_s = stage ? (uint64_t *)sl0.h_out_a.p : starts, *dst_e = stage ? (uint64_t *)sl0.h_out_b.p : ends;
        CUDA_TRY(cudaMemcpyAsync(dst_s + off, d_s + off, cn * 8, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(dst_e + off, d_e + off, cn * 8, cudaMemcpyDeviceToHost, cs));
    }
    for (int s2 = 0; s2 < kSlots; ++s2) CUDA_TRY(cudaStreamSynchronize(ws->slot[s2].stream));
    CUDA_TRY(cudaMemcpyAsync(ws->small.h, ws->small.d, 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    t_stats.kernel_launches = launches;
    if (ws->small.h[1] != kNoError)  // checked on the device: text_with_rank_support/mod.rs:106-110
        return fail(GDX_ERR_BAD_ARG, "cursor %llu is outside [0, text_len]", (unsigned long long)ws->small.h[1]);
    if (ws->small.h[0] != kNoError) {
        t_error_query = ws->small.h[0];
        return fail(GDX_ERR_INVALID_SYMBOL, "cursor %llu: symbol in io representation should be valid (alphabet.rs:195-198)",
                    (unsigned long long)ws->small.h[0]);
    }
    if (stage) {
        HostPool::get().copy(starts, sl0.h_out_a.p, n * 8);
        HostPool::get().copy(ends, sl0.h_out_b.p, n * 8);
    }
    return GDX_OK;
}

extern "C" gdx_status gdx_cursor_for_query(const gdx_index *idx, const uint8_t *query, uint64_t len, uint64_t *start,
                                           uint64_t *end) {
    if (!idx || !start || !end || (len && !query)) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    // single-query path of the reference (lib.rs:217-235): when the lookup-table interval is already
    // empty, the next symbol to the left is still translated (and may panic) before the loop breaks.
    const uint32_t D = idx->h.lookup_depth;
    gdx_queries q = {query, nullptr, len, 1};
    uint8_t dummy = 0;
    if (len == 0) q.bytes = &dummy;
    gdx_status st = gdx_cursors_many(idx, &q, start, end);
    if (st != GDX_OK) return st;
    if (D > 0 && len > D && *start == *end) {
        gdx_queries suffix = {query + (len - D), nullptr, D, 1};
        uint64_t s2 = 0, e2 = 0;
        GDX_TRY(gdx_cursors_many(idx, &suffix, &s2, &e2));
        if (s2 == e2 && idx->h.io_to_dense[query[len - D - 1]] == 0) {
            t_error_query = 0;
            return fail(GDX_ERR_INVALID_SYMBOL, "symbol in io representation should be valid (alphabet.rs:195-198)");
        }
    }
    return GDX_OK;
}

// ================================================================================================
// device-resident entry points
// ================================================================================================
static gdx_status search_device(const gdx_index *idx, const gdx_queries *dq_in, uint64_t *a, uint64_t *b, int mode,
                                uint64_t *d_error, void *stream) {
    if (!idx || !dq_in) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(idx->device);
    DevQueries dq;
    dq.bytes = dq_in->bytes;
    dq.offsets = dq_in->offsets;
    dq.fixed_len = dq_in->fixed_len;
    dq.nq = dq_in->nq;
    dq.base = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const SortPlan sp = plan_sort(idx, dq.nq, dq.offsets ? 0 : dq.fixed_len);
    const uint32_t *perm = nullptr;
    void *scratch = nullptr;
    if (sp.use) {  // stream-ordered scratch from the device's memory pool
        CUDA_TRY(cudaMallocAsync(&scratch, sp.total_bytes, st));
        GDX_TRY(sort_queries(idx, dq, sp, scratch, st, &perm));
    }
    GDX_TRY(dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
        launch_search<decltype(L)>(idx, dq, a, b, mode, 0, d_error, nullptr, perm, st);
        return GDX_OK;
    }));
    CUDA_TRY(cudaGetLastError());
    if (scratch) CUDA_TRY(cudaFreeAsync(scratch, st));
    return GDX_OK;
}

extern "C" gdx_status gdx_cursors_many_device(const gdx_index *idx, const gdx_queries *d_queries, uint64_t *d_starts,
                                              uint64_t *d_ends, uint64_t *d_error, void *stream) {
    return search_device(idx, d_queries, d_starts, d_ends, 0, d_error, stream);
}
extern "C" gdx_status gdx_count_many_device(const gdx_index *idx, const gdx_queries *d_queries, uint64_t *d_counts,
                                            uint64_t *d_error, void *stream) {
    return search_device(idx, d_queries, d_counts, nullptr, 1, d_error, stream);
}

extern "C" gdx_status gdx_locate_intervals_device(const gdx_index *idx, const uint64_t *d_starts,
                                                  const uint64_t *d_ends, uint64_t n, const uint64_t *d_hit_offsets,
                                                  uint64_t num_hits, gdx_hit *d_hits, void *stream) {
    if (!idx) return fail(GDX_ERR_BAD_ARG, "idx is NULL");
    if (n == 0 || num_hits == 0) return GDX_OK;
    if (!d_starts || !d_ends || !d_hit_offsets || !d_hits) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(idx->device);
    cudaStream_t st = (cudaStream_t)stream;
    // stream-ordered scratch: rows (8 B per hit), worklist of wide intervals, two counters
    uint64_t *rows = nullptr, *big = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&rows, num_hits * 8, st));
    CUDA_TRY(cudaMallocAsync((void **)&big, (n + 2) * 8, st));
    CUDA_TRY(cudaMemsetAsync(big, 0, 16, st));
    k_expand_rows<<<(unsigned)div_up(n, 256), 256, 0, st>>>(d_starts, d_ends, d_hit_offsets, n, rows, big + 2,
                                                           reinterpret_cast<unsigned long long *>(big));
    CUDA_TRY(cudaGetLastError());
    uint64_t nbig = 0;  // the number of wide intervals decides the next grid: one small D2H + sync
    CUDA_TRY(cudaMemcpyAsync(&nbig, big, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (nbig) {
        dim3 grid((unsigned)nbig, 32);
        k_expand_big_rows<<<grid, 256, 0, st>>>(d_starts, d_ends, d_hit_offsets, big + 2, rows);
    }
    GDX_TRY(dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
        k_locate_walk<decltype(L)><<<(unsigned)div_up(num_hits, 256), 256, 0, st>>>(
            idx->dev, rows, num_hits, reinterpret_cast<ulonglong2 *>(d_hits), nullptr);
        return GDX_OK;
    }));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(rows, st));
    CUDA_TRY(cudaFreeAsync(big, st));
    return GDX_OK;
}

// ================================================================================================
// random-gather ceiling
// ================================================================================================
__global__ void k_fill_random(uint64_t *p, uint64_t nwords) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords;
         i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = mix64(i);
}

extern "C" gdx_status gdx_measure_random_gather(int32_t device_req, uint64_t table_bytes, uint32_t record_bytes,
                                                uint64_t loads, int32_t chained, double *gbps_out,
                                                double *gloads_out) {
    if (record_bytes != 32 && record_bytes != 64 && record_bytes != 128)
        return fail(GDX_ERR_BAD_ARG, "record_bytes must be 32, 64 or 128");
    int device;
    GDX_TRY(resolve_device(device_req, &device));
    DeviceGuard guard(device);
    uint64_t nrec = 1;
    while (nrec * 2 * record_bytes <= table_bytes) nrec *= 2;
    uint8_t *table = nullptr;
    uint64_t *sink = nullptr;
    CUDA_TRY(cudaMalloc(&table, nrec * record_bytes));
    cudaError_t e = cudaMalloc(&sink, 8);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0;
    const unsigned blocks = 148 * 8, threads = 256;
    uint32_t rounds = (uint32_t)std::max<uint64_t>(1, loads / (2ull * blocks * threads));
    auto launch = [&](uint32_t r) {
        if (record_bytes == 32) k_gather<32><<<blocks, threads>>>(table, nrec - 1, r, chained, sink);
        else if (record_bytes == 64) k_gather<64><<<blocks, threads>>>(table, nrec - 1, r, chained, sink);
        else k_gather<128><<<blocks, threads>>>(table, nrec - 1, r, chained, sink);
    };
    if (e == cudaSuccess) {
        k_fill_random<<<blocks, threads>>>((uint64_t *)table, nrec * record_bytes / 8);
        launch(rounds / 8 + 1);  // warm-up
        e = cudaEventCreate(&e0);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaEventRecord(e0);
    if (e == cudaSuccess) {
        launch(rounds);
        e = cudaEventRecord(e1);
    }
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(table);
    if (sink) cudaFree(sink);
    if (e != cudaSuccess) return fail(GDX_ERR_CUDA, "gather microbenchmark failed: %s", cudaGetErrorString(e));
    const double nloads = 2.0 * blocks * threads * (double)rounds;
    if (gloads_out) *gloads_out = nloads / (ms * 1e-3) / 1e9;
    if (gbps_out) *gbps_out = nloads * record_bytes / (ms * 1e-3) / 1e9;
    return GDX_OK;
}
