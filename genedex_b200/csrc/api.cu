// api.cu -- C ABI of genedex_b200 (include/genedex_b200.h): index handles, device image
// construction from host parts, chunked H2D / kernel / D2H pipelines, locate CSR plumbing.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <type_traits>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <new>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/genedex_b200.h"
#include "device_build.h"
#include "device_index.h"
#include "host_build.h"
#include "host_pack.h"
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
        if (_e != cudaSuccess) {                                                                 \
            cudaGetLastError(); /* a failed allocation must not fail the next call's error check */ \
            return fail(_e == cudaErrorMemoryAllocation ? GDX_ERR_OOM : GDX_ERR_CUDA,            \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,        \
                        __LINE__);                                                               \
        }                                                                                        \
    } while (0)

#define GDX_TRY(expr)                  \
    do {                               \
        gdx_status _s = (expr);        \
        if (_s != GDX_OK) return _s;   \
    } while (0)

// exceptions must not cross the C boundary (std::bad_alloc from a host vector, std::system_error from a thread)
template <class F>
gdx_status guarded(F &&f) noexcept {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return fail(GDX_ERR_OOM, "out of host memory");
    } catch (const std::exception &e) {
        return fail(GDX_ERR_BAD_ARG, "unexpected failure: %s", e.what());
    } catch (...) {
        return fail(GDX_ERR_BAD_ARG, "unexpected failure");
    }
}

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

uint64_t env_bytes(const char *name, uint64_t dflt) {
    const char *e = getenv(name);
    return e && *e ? (uint64_t)strtoull(e, nullptr, 10) : dflt;
}
// smaller transfers go through the driver's own staging (GDX_STAGE_MIN_BYTES: tests force the staged paths)
const uint64_t kStageMinBytes = env_bytes("GDX_STAGE_MIN_BYTES", 4ull << 20);

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
    // staging for pageable caller buffers / packed queries / narrow results
    HBuf h_in, h_off, h_out_a, h_out_b;
    // queries of a packed chunk that hold a byte without a 2-bit code: offsets + slots + IO bytes, re-run raw
    HBuf h_x;
    DBuf xbuf;
    std::vector<cudaEvent_t> ev_loc;  // pairs (begin, end) around the locate kernels of the current call
    size_t ev_loc_used = 0;
    cudaEvent_t ev_h2d = nullptr, ev_out = nullptr;
    cudaEvent_t ev_link = nullptr;  // the chunk's query upload has left the PCIe link
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
            if (cudaEventCreateWithFlags(&s.ev_link, cudaEventDisableTiming) != cudaSuccess) return false;
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
            if (s.ev_link) cudaEventDestroy(s.ev_link);
            for (HBuf *b : {&s.h_in, &s.h_off, &s.h_out_a, &s.h_out_b, &s.h_x}) b->release();
            s.xbuf.release();
            for (auto e : s.ev) cudaEventDestroy(e);
            for (auto e : s.ev_loc) cudaEventDestroy(e);
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
    cudaEvent_t next_locate_event(Slot &s) {
        if (s.ev_loc_used == s.ev_loc.size()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            s.ev_loc.push_back(e);
        }
        return s.ev_loc[s.ev_loc_used++];
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
    void *seed_lut = nullptr;      // accelerator outside the image (gdx_index_set_seed_table_depth)
    uint64_t seed_lut_bytes = 0;
    void *row_ctx = nullptr;       // accelerator outside the image (gdx_index_set_row_context_table)
    uint64_t row_ctx_bytes = 0;
    bool verify = true;            // finish one-row intervals through the text section (gdx_index_set_text_verification)
    PackTable pack;                // io byte -> 2-bit code (host packer, host_pack.h)
    // host-buffer queries hold this shared for the whole call; (re)building or freeing an accelerator needs it
    // exclusively and reports GDX_ERR_BUSY instead of pulling memory from under a running kernel
    mutable std::shared_mutex cfg_mu;
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
    const uint32_t *h_samples32 = nullptr;  // host, the crate's own Vec<u32> for 32-bit storage
    // accelerator policy of the index (travels with the image)
    uint32_t accel_flags = 0;
    uint64_t accel_budget = 0;
};

uint32_t accel_flags_from_config(uint32_t flags) {
    return ((flags & GDX_FLAG_NO_DENSE_SUFFIX_ARRAY) ? kAccelNoDenseSA : 0u) |
           ((flags & GDX_FLAG_NO_SEED_TABLE) ? kAccelNoSeedTable : 0u) |
           ((flags & GDX_FLAG_NO_ROW_CONTEXT_TABLE) ? kAccelNoRowContext : 0u);
}

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

// wide_override: -1 = decide from the text length (and GDX_FORCE_WIDE), 0 / 1 = as given (header validation)
gdx_status plan_header(const ImageSources &src, ImageHeader &h, int wide_override = -1) {
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
    h.accel_flags = src.accel_flags;
    h.accel_budget = src.accel_budget;
    // 64-bit samples / lookup entries are only needed beyond 2^32 - 1 symbols; GDX_FORCE_WIDE=1 selects them
    // for any text so that this path can be tested without a 4.3 G symbol index
    const char *fw = getenv("GDX_FORCE_WIDE");
    h.wide = (src.n > 0xffffffffull || (fw && atoi(fw) != 0)) ? 1 : 0;
    if (wide_override >= 0) h.wide = (src.n > 0xffffffffull || wide_override) ? 1 : 0;
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

// An image header that comes from a file or a peer is untrusted: every size, offset and the record layout are
// re-derived from its scalar fields and must match, so that no kernel ever indexes with a value the planner
// would not have produced.
gdx_status validate_header(const ImageHeader &h) {
    if (h.magic != kImageMagic || h.version != GDX_ABI_VERSION)
        return fail(GDX_ERR_BAD_ARG, "not a genedex_b200 image header of this version (magic/version mismatch)");
    gdx_alphabet a;
    memcpy(a.io_to_dense, h.io_to_dense, 256);
    a.num_dense_symbols = h.sigma;
    a.num_searchable_dense_symbols = h.ns;
    GDX_TRY(validate_alphabet(a));
    if (h.sampling_rate == 0 || h.lookup_depth > kMaxLookupDepth || h.storage > GDX_I64 || h.ntexts == 0 ||
        h.n_border > h.n || h.ntexts > h.n || (h.text_bits != 0 && h.text_bits != 4 && h.text_bits != 8) || h.has_isa > 1 ||
        h.wide > 1 || h.n > (1ull << 48))
        return fail(GDX_ERR_BAD_ARG, "image header: field out of range");
    ImageSources src;
    src.alphabet = &a;
    src.storage = h.storage;
    src.sampling_rate = h.sampling_rate;
    src.lookup_depth = h.lookup_depth;
    src.n = h.n;
    src.ntexts = h.ntexts;
    src.n_border = h.n_border;
    src.d_text = h.text_bits ? reinterpret_cast<const uint8_t *>(&a) : nullptr;            // presence only
    src.d_isa32 = h.has_isa ? reinterpret_cast<const uint32_t *>(&a) : nullptr;            // presence only
    src.accel_flags = h.accel_flags;
    src.accel_budget = h.accel_budget;
    ImageHeader want;
    GDX_TRY(plan_header(src, want, (int)h.wide));
    if (memcmp(&want, &h, sizeof(ImageHeader)) != 0)
        return fail(GDX_ERR_BAD_ARG, "image header is inconsistent (sizes / offsets do not match its own fields)");
    return GDX_OK;
}

// run-time overrides of kernel parameters (measurements only)
bool verify_enabled();
gdx_status init_policies(gdx_index *idx) {
    build_pack_table(idx->h.io_to_dense, idx->h.ns, idx->pack);
    idx->verify = verify_enabled();
    if (const char *vm = getenv("GDX_VERIFY_MIN"))
        if (atoi(vm) > 0) idx->dev.verify_min_remaining = (uint32_t)atoi(vm);
    // count: also finish intervals of up to this many rows by text comparisons (A/B: profiles/README.md)
    if (const char *vr = getenv("GDX_VERIFY_ROWS"))
        if (atoi(vr) > 0) idx->dev.verify_max_rows = (uint32_t)std::min(atoi(vr), 64);
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

// ---- row context table accelerator (include/genedex_b200.h: gdx_index_set_row_context_table) ------------
void drop_row_context(gdx_index *idx) {
    if (!idx->row_ctx) return;
    idx->dev.row_context = nullptr;
    cudaFree(idx->row_ctx);
    idx->row_ctx = nullptr;
    idx->row_ctx_bytes = 0;
}

bool row_context_possible(const gdx_index *idx) {
    const ImageHeader &h = idx->h;
    return h.n > 0 && !h.wide && h.n < (1ull << 32) && h.ns >= 1 && h.ns <= 4 && h.text_bits != 0;
}

gdx_status build_row_context(gdx_index *idx) {
    if (idx->row_ctx) return GDX_OK;
    if (!row_context_possible(idx))
        return fail(GDX_ERR_UNSUPPORTED, "the row context table needs the text section, at most 4 searchable symbols and a "
                                         "text shorter than 2^32 symbols");
    const uint64_t n = idx->h.n, bytes = n * 16;
    void *d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(GDX_ERR_OOM, "row context table: %llu bytes of device memory not available", (unsigned long long)bytes);
    }
    gdx_status st = dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
        k_build_row_context<decltype(L)><<<(unsigned)div_up(n, 256), 256>>>(idx->dev, reinterpret_cast<uint4 *>(d), n);
        return GDX_OK;
    });
    cudaError_t e = st == GDX_OK ? cudaDeviceSynchronize() : cudaSuccess;
    if (st != GDX_OK || e != cudaSuccess) {
        cudaFree(d);
        return st != GDX_OK ? st : fail(GDX_ERR_CUDA, "row context table: %s", cudaGetErrorString(e));
    }
    idx->row_ctx = d;
    idx->row_ctx_bytes = bytes;
    idx->dev.row_context = d;
    return GDX_OK;
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

// Accelerator policy (include/genedex_b200.h): flags and byte budget live in the image header, so every way of
// making a replica follows the same rule.  Without a budget an accelerator may take a quarter of the memory
// that is free right now.  GDX_DENSE_SA / GDX_SEED_TABLE override the policy for measurements.
uint64_t accel_room(const gdx_index *idx) {
    if (idx->h.accel_budget) {
        const uint64_t used = idx->dense_sa_bytes + idx->seed_lut_bytes + idx->row_ctx_bytes;
        return idx->h.accel_budget > used ? idx->h.accel_budget - used : 0;
    }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 0;
    return free_b / 4;
}

void auto_seed_table(gdx_index *idx) {
    if (!idx || idx->h.n == 0 || idx->h.ns < 2) return;
    const char *e = getenv("GDX_SEED_TABLE");
    uint32_t depth = 0;
    if (e) {
        if (atoi(e) <= 0) return;
        depth = (uint32_t)atoi(e);
    } else {
        if (idx->h.accel_flags & kAccelNoSeedTable) return;
        // the deepest level with at most four entries per text position (then most k-mers of the text have one row
        // and almost no LF step is left): 3.1 Gbp DNA -> depth 16 (34 GB; 0.82 instead of 0.97 ms per 7.5 M queries
        // against depth 15), 500 M residues of protein -> depth 7 (10 GB); the budget below has the last word
        while (seed_entries(idx->h.ns, depth + 1) && seed_entries(idx->h.ns, depth + 1) <= 4 * idx->h.n) ++depth;
        const uint64_t esz = idx->h.wide ? 16 : 8, room = accel_room(idx);
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return;
        // the finished level must fit the budget; the temporary previous level only has to fit the device
        while (depth > 0 && (seed_entries(idx->h.ns, depth) * esz > room ||
                             (seed_entries(idx->h.ns, depth) + seed_entries(idx->h.ns, depth - 1)) * esz > free_b))
            --depth;
    }
    if (depth <= idx->h.lookup_depth) return;  // the configured table is at least as deep
    if (build_seed_table(idx, depth) != GDX_OK) t_error.clear();  // optional: not an error of the call
}

// best effort after every way of creating a replica; the caller holds a DeviceGuard and has freed its temporaries
void auto_dense_sa(gdx_index *idx) {
    if (!idx || idx->h.n == 0) return;
    const char *e = getenv("GDX_DENSE_SA");
    if (e && atoi(e) == 0) return;
    if (!(e && atoi(e) != 0)) {
        if (idx->h.accel_flags & kAccelNoDenseSA) return;
        if (idx->h.n * (idx->h.wide ? 8ull : 4ull) > accel_room(idx)) return;
    }
    if (build_dense_sa(idx) != GDX_OK) t_error.clear();  // optional: not an error of the call
}

// built last (it is the largest: 16 B per text position) and only from what the other two left: the budget when
// one is set, else up to half of the memory that is still free
void auto_row_context(gdx_index *idx) {
    if (!idx || !row_context_possible(idx)) return;
    const char *e = getenv("GDX_ROW_CONTEXT");
    if (e && atoi(e) == 0) return;
    if (!(e && atoi(e) != 0)) {
        if (idx->h.accel_flags & kAccelNoRowContext) return;
        const uint64_t room = idx->h.accel_budget ? accel_room(idx) : 2 * accel_room(idx);
        if (idx->h.n * 16 > room) return;
    }
    if (build_row_context(idx) != GDX_OK) t_error.clear();  // optional: not an error of the call
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
                constexpr int B = LT::kPlanes;
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
        } else if (src.h_samples32) {
            if (!h.wide) {
                IMG_TRY(cudaMemcpy(dst, src.h_samples32, h.n_samples * 4, cudaMemcpyHostToDevice));
            } else {
                std::vector<uint64_t> wide(h.n_samples);
                for (uint64_t i = 0; i < h.n_samples; ++i) wide[i] = src.h_samples32[i];
                IMG_TRY(cudaMemcpy(dst, wide.data(), h.n_samples * 8, cudaMemcpyHostToDevice));
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

gdx_status check_text_offsets(const uint64_t *text_offsets, uint64_t num_texts) {
    for (uint64_t i = 0; i < num_texts; ++i)
        if (text_offsets[i + 1] < text_offsets[i])
            return fail(GDX_ERR_BAD_ARG, "text_offsets must be non-decreasing (text %llu)", (unsigned long long)i);
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
extern "C" uint32_t gdx_host_pool_resize(uint32_t threads) {
    try {
        return HostPool::get().resize(threads);
    } catch (...) {  // (a thread could not be started: the pool keeps the workers it has)
        return HostPool::get().threads();
    }
}
extern "C" void gdx_host_pack_tuning(int32_t prefetch_bytes, int32_t streaming_stores) { set_pack_tuning(prefetch_bytes, streaming_stores); }
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
    return guarded([&]() -> gdx_status {
        if (!out) return fail(GDX_ERR_BAD_ARG, "out is NULL");
        *out = nullptr;
        if (!text_offsets || !alphabet || !config) return fail(GDX_ERR_BAD_ARG, "NULL argument");
        if (num_texts == 0) return fail(GDX_ERR_BAD_ARG, "there should be at least one text (construction/mod.rs:300)");
        GDX_TRY(check_text_offsets(text_offsets, num_texts));
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
        src.accel_flags = accel_flags_from_config(config->flags);
        src.accel_budget = config->accelerator_budget_bytes;
    
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
                auto_dense_sa(*out);
                auto_seed_table(*out);
                auto_row_context(*out);
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
            auto_dense_sa(*out);
            auto_seed_table(*out);
            auto_row_context(*out);
        }
        return st;
    });
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
    if ((parts->sampled_suffix_array != nullptr) == (parts->sampled_suffix_array_u32 != nullptr) && parts->text_len)
        return fail(GDX_ERR_BAD_ARG, "exactly one of sampled_suffix_array / sampled_suffix_array_u32 must be set");
    src.h_samples64 = parts->sampled_suffix_array;
    src.h_samples32 = parts->sampled_suffix_array_u32;
    src.accel_flags = accel_flags_from_config(parts->flags);
    src.accel_budget = parts->accelerator_budget_bytes;
    return GDX_OK;
}

extern "C" gdx_status gdx_index_create_from_bwt(const uint8_t *bwt, const gdx_parts *parts, int32_t device_req,
                                                gdx_index **out) {
    return guarded([&]() -> gdx_status {
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
            auto_row_context(*out);
        }
        return st;
    });
}

extern "C" gdx_status gdx_index_create_from_parts(const gdx_parts *parts, int32_t device_req, gdx_index **out) {
    return guarded([&]() -> gdx_status {
        if (!out) return fail(GDX_ERR_BAD_ARG, "out is NULL");
        *out = nullptr;
        ImageSources src;
        std::vector<uint64_t> rows, pos;
        GDX_TRY(sources_from_parts(parts, src, rows, pos));
        if (!parts->interleaved_blocks) return fail(GDX_ERR_BAD_ARG, "interleaved_blocks is NULL");
        int device;
        GDX_TRY(resolve_device(device_req, &device));
        DeviceGuard guard(device);
        // the reference's blocks -> dense BWT on the device (symbol_at of the variant), then the common path
        const uint32_t block_bits = parts->block_bits ? parts->block_bits : 64;
        if ((block_bits != 64 && block_bits != 512) || parts->rank_variant > GDX_RANK_FLAT)
            return fail(GDX_ERR_BAD_ARG, "rank_variant must be GDX_RANK_CONDENSED or GDX_RANK_FLAT, block_bits 64 or 512");
        const bool flat = parts->rank_variant == GDX_RANK_FLAT;
        const uint32_t words = block_bits / 64, used = block_bits - (flat ? 16 : 0);
        const uint32_t units = flat ? parts->alphabet.num_dense_symbols : choose_layout(parts->alphabet.num_dense_symbols).planes;
        const uint64_t nwords = div_up(src.n + 1, used) * units * words;
        uint64_t *d_blocks = nullptr;
        uint8_t *d_bwt = nullptr;
        CUDA_TRY(cudaMalloc(&d_blocks, nwords * 8));
        cudaError_t e = cudaMalloc(&d_bwt, src.n ? src.n : 1);
        if (e == cudaSuccess) e = cudaMemcpy(d_blocks, parts->interleaved_blocks, nwords * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && src.n) {
            k_ref_blocks_to_bwt<<<(unsigned)div_up(src.n, 256), 256>>>(d_blocks, flat ? 1u : 0u, words, units, src.n, d_bwt);
            e = cudaGetLastError();
        }
        cudaFree(d_blocks);
        if (e != cudaSuccess) {
            if (d_bwt) cudaFree(d_bwt);
            return fail(GDX_ERR_CUDA, "block upload failed: %s", cudaGetErrorString(e));
        }
        src.d_bwt = d_bwt;
        gdx_status st = build_image(src, device, out);
        cudaFree(d_bwt);
        if (st == GDX_OK) {
            auto_dense_sa(*out);
            auto_seed_table(*out);
            auto_row_context(*out);
        }
        return st;
    });
}

extern "C" gdx_status gdx_concat_texts(const uint8_t *texts, const uint64_t *text_offsets, uint64_t num_texts,
                                       const gdx_alphabet *alphabet, uint8_t *dense_out, uint64_t *sentinels_out,
                                       uint64_t *count_out) {
    return guarded([&]() -> gdx_status {
        if (!text_offsets || !alphabet || !dense_out || !sentinels_out || !count_out || num_texts == 0)
            return fail(GDX_ERR_BAD_ARG, "NULL argument or no texts");
        GDX_TRY(check_text_offsets(text_offsets, num_texts));
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
    });
}

extern "C" gdx_status gdx_suffix_array(const uint8_t *dense_text, uint64_t n, uint32_t sigma, uint32_t where,
                                       int32_t device_req, uint64_t *sa_out) {
    return guarded([&]() -> gdx_status {
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
    });
}

extern "C" gdx_status gdx_index_download_bwt(const gdx_index *idx, uint8_t *bwt_out) {
    return guarded([&]() -> gdx_status {
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
    });
}

extern "C" gdx_status gdx_index_download_samples(const gdx_index *idx, uint64_t *samples_out) {
    return guarded([&]() -> gdx_status {
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
    });
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
    if (idx->row_ctx) cudaFree(idx->row_ctx);
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
    out->row_context_entry_bytes = idx->row_ctx ? 16 : 0;
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
    return guarded([&]() -> gdx_status {
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
    });
}

extern "C" gdx_status gdx_index_load_from_file(const char *path, int32_t device_req, gdx_index **out,
                                               void *user_data_out, uint64_t user_capacity, uint64_t *user_bytes_out) {
    return guarded([&]() -> gdx_status {
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
        GDX_TRY(validate_header(h));
        // the sizes in the prefix must add up to the size of the file before anything is allocated from them
        const off_t here = ftello(fc.f);
        if (here < 0 || fseeko(fc.f, 0, SEEK_END) != 0) return fail(GDX_ERR_BAD_ARG, "cannot seek in %s", path);
        const off_t file_size = ftello(fc.f);
        if (fseeko(fc.f, here, SEEK_SET) != 0 || file_size < here || p.user_bytes > (uint64_t)(file_size - here) ||
            p.image_bytes != (uint64_t)(file_size - here) - p.user_bytes)
            return fail(GDX_ERR_BAD_ARG, "%s is truncated or corrupt (section sizes do not add up to the file size)", path);
        if (user_bytes_out) *user_bytes_out = p.user_bytes;
        if (p.user_bytes) {
            const uint64_t keep = user_data_out ? std::min<uint64_t>(p.user_bytes, user_capacity) : 0;
            if (keep && fread(user_data_out, keep, 1, fc.f) != 1) return fail(GDX_ERR_BAD_ARG, "%s is truncated", path);
            if (fseeko(fc.f, here + (off_t)p.user_bytes, SEEK_SET) != 0) return fail(GDX_ERR_BAD_ARG, "cannot seek in %s", path);
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
    });
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
    return guarded([&]() -> gdx_status {
        if (!header || !device_image || !out) return fail(GDX_ERR_BAD_ARG, "NULL argument");
        *out = nullptr;
        ImageHeader h;
        memcpy(&h, header, sizeof h);
        GDX_TRY(validate_header(h));
        int device;
        GDX_TRY(resolve_device(device_req, &device));
        gdx_index *idx = new (std::nothrow) gdx_index();
        if (!idx) return fail(GDX_ERR_OOM, "out of host memory");
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
            auto_row_context(idx);
        }
        *out = idx;
        return GDX_OK;
    });
}

extern "C" gdx_status gdx_index_set_seed_table_depth(gdx_index *idx, int32_t depth) {
    return guarded([&]() -> gdx_status {
        if (!idx) return fail(GDX_ERR_BAD_ARG, "NULL argument");
        std::unique_lock<std::shared_mutex> cfg(idx->cfg_mu, std::try_to_lock);
        if (!cfg.owns_lock()) return fail(GDX_ERR_BUSY, "queries are running on this index: the seed table cannot change now");
        DeviceGuard guard(idx->device);
        if (depth <= 0) {
            drop_seed_table(idx);
            return GDX_OK;
        }
        return build_seed_table(idx, (uint32_t)depth);
    });
}

extern "C" gdx_status gdx_index_set_row_context_table(gdx_index *idx, int32_t on) {
    return guarded([&]() -> gdx_status {
        if (!idx) return fail(GDX_ERR_BAD_ARG, "NULL argument");
        std::unique_lock<std::shared_mutex> cfg(idx->cfg_mu, std::try_to_lock);
        if (!cfg.owns_lock()) return fail(GDX_ERR_BUSY, "queries are running on this index: the row context table cannot change now");
        DeviceGuard guard(idx->device);
        if (on) return build_row_context(idx);
        drop_row_context(idx);
        return GDX_OK;
    });
}

extern "C" gdx_status gdx_index_set_text_verification(gdx_index *idx, int32_t on) {
    return guarded([&]() -> gdx_status {
        if (!idx) return fail(GDX_ERR_BAD_ARG, "NULL argument");
        std::unique_lock<std::shared_mutex> cfg(idx->cfg_mu, std::try_to_lock);
        if (!cfg.owns_lock()) return fail(GDX_ERR_BUSY, "queries are running on this index");
        idx->verify = on != 0;
        return GDX_OK;
    });
}

extern "C" gdx_status gdx_index_set_dense_suffix_array(gdx_index *idx, int32_t on) {
    return guarded([&]() -> gdx_status {
        if (!idx) return fail(GDX_ERR_BAD_ARG, "NULL argument");
        std::unique_lock<std::shared_mutex> cfg(idx->cfg_mu, std::try_to_lock);
        if (!cfg.owns_lock()) return fail(GDX_ERR_BUSY, "queries are running on this index: the dense suffix array cannot change now");
        DeviceGuard guard(idx->device);
        if (on) return build_dense_sa(idx);
        drop_dense_sa(idx);
        return GDX_OK;
    });
}

// ---- NCCL, loaded at run time ---------------------------------------------------------------------------
// The library does not link libnccl: a process that already carries one (PyTorch bundles its own libnccl.so.2)
// must not get a second copy, and single-GPU users need none.  The handful of entry points used here has been
// ABI-stable since NCCL 2.
namespace {
struct Nccl {
    typedef struct ncclComm *comm_t;
    struct unique_id {
        char internal[GDX_NCCL_UNIQUE_ID_BYTES];
    };
    int (*GetUniqueId)(unique_id *) = nullptr;
    int (*CommInitRank)(comm_t *, int, unique_id, int) = nullptr;
    int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int /*ncclDataType_t*/, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
    static constexpr int kUint8 = 1;  // ncclUint8
};

const Nccl &nccl() {
    static const Nccl n = [] {
        Nccl r;
        const char *names[] = {getenv("GDX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        void *h = nullptr;
        for (const char *nm : names)
            if (nm && *nm && (h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!h) {
            r.why = "libnccl.so.2 could not be loaded";
            return r;
        }
        auto sym = [&](const char *name) { return dlsym(h, name); };
        r.GetUniqueId = (decltype(r.GetUniqueId))sym("ncclGetUniqueId");
        r.CommInitRank = (decltype(r.CommInitRank))sym("ncclCommInitRank");
        r.CommInitAll = (decltype(r.CommInitAll))sym("ncclCommInitAll");
        r.CommDestroy = (decltype(r.CommDestroy))sym("ncclCommDestroy");
        r.Broadcast = (decltype(r.Broadcast))sym("ncclBroadcast");
        r.GroupStart = (decltype(r.GroupStart))sym("ncclGroupStart");
        r.GroupEnd = (decltype(r.GroupEnd))sym("ncclGroupEnd");
        r.GetErrorString = (decltype(r.GetErrorString))sym("ncclGetErrorString");
        r.ok = r.GetUniqueId && r.CommInitRank && r.CommInitAll && r.CommDestroy && r.Broadcast && r.GroupStart &&
               r.GroupEnd && r.GetErrorString;
        if (!r.ok) r.why = "libnccl.so.2 lacks an expected entry point";
        return r;
    }();
    return n;
}

#define NCCL_TRY(expr)                                                                                     \
    do {                                                                                                   \
        int _r = (expr);                                                                                   \
        if (_r != 0) return fail(GDX_ERR_CUDA, "%s failed: %s", #expr, nccl().GetErrorString(_r));         \
    } while (0)

thread_local const char *t_transport = "";
constexpr uint64_t kBroadcastPiece = 1ull << 30;
}  // namespace

extern "C" const char *gdx_replicate_transport(void) { return t_transport; }

extern "C" gdx_status gdx_index_replicate(const gdx_index *idx, const int32_t *devices, int32_t n_devices,
                                          gdx_index **out_replicas) {
    return guarded([&]() -> gdx_status {
        if (!idx || !devices || !out_replicas || n_devices < 0) return fail(GDX_ERR_BAD_ARG, "bad argument");
        t_transport = "";
        for (int i = 0; i < n_devices; ++i) out_replicas[i] = nullptr;
        std::vector<int> dev(n_devices);
        for (int i = 0; i < n_devices; ++i) GDX_TRY(resolve_device(devices[i], &dev[i]));
        std::vector<void *> img(n_devices, nullptr);
        auto free_images = [&] {
            for (int i = 0; i < n_devices; ++i)
                if (img[i]) {
                    DeviceGuard g(dev[i]);
                    cudaFree(img[i]);
                    img[i] = nullptr;
                }
        };
        const uint64_t bytes = idx->h.image_bytes;
        for (int i = 0; i < n_devices; ++i) {
            DeviceGuard g(dev[i]);
            cudaError_t e = cudaMalloc(&img[i], bytes ? bytes : 1);
            if (e != cudaSuccess) {
                free_images();
                return fail(GDX_ERR_OOM, "replica on device %d: %s", dev[i], cudaGetErrorString(e));
            }
        }
        // communicator over the source device + every distinct other target; targets on the source device (and
        // repeated targets) are filled by device-to-device copies afterwards
        std::vector<int> comm_dev{idx->device};
        std::vector<int> member(n_devices, -1);  // rank of target i in the communicator, -1 = copy
        for (int i = 0; i < n_devices; ++i) {
            if (std::find(comm_dev.begin(), comm_dev.end(), dev[i]) != comm_dev.end()) continue;
            member[i] = (int)comm_dev.size();
            comm_dev.push_back(dev[i]);
        }
        const bool want_nccl = comm_dev.size() > 1 && !(getenv("GDX_REPLICATE") && strcmp(getenv("GDX_REPLICATE"), "peer") == 0);
        bool done = comm_dev.size() <= 1;
        if (want_nccl && nccl().ok) {
            const Nccl &N = nccl();
            const int world = (int)comm_dev.size();
            std::vector<Nccl::comm_t> comms(world, nullptr);
            std::vector<cudaStream_t> streams(world, nullptr);
            gdx_status st = [&]() -> gdx_status {
                NCCL_TRY(N.CommInitAll(comms.data(), world, comm_dev.data()));
                for (int r = 0; r < world; ++r) {
                    DeviceGuard g(comm_dev[r]);
                    CUDA_TRY(cudaStreamCreateWithFlags(&streams[r], cudaStreamNonBlocking));
                }
                for (uint64_t off = 0; off < bytes; off += kBroadcastPiece) {
                    const uint64_t nb = std::min(kBroadcastPiece, bytes - off);
                    NCCL_TRY(N.GroupStart());
                    for (int r = 0; r < world; ++r) {
                        void *recv = (uint8_t *)idx->image + off;  // root: in place
                        for (int i = 0; i < n_devices && r > 0; ++i)
                            if (member[i] == r) recv = (uint8_t *)img[i] + off;
                        DeviceGuard g(comm_dev[r]);
                        NCCL_TRY(N.Broadcast((const uint8_t *)idx->image + off, recv, nb, Nccl::kUint8, 0, comms[r], streams[r]));
                    }
                    NCCL_TRY(N.GroupEnd());
                }
                for (int r = 0; r < world; ++r) {
                    DeviceGuard g(comm_dev[r]);
                    CUDA_TRY(cudaStreamSynchronize(streams[r]));
                }
                return GDX_OK;
            }();
            for (int r = 0; r < world; ++r) {
                DeviceGuard g(comm_dev[r]);
                if (streams[r]) cudaStreamDestroy(streams[r]);
                if (comms[r]) N.CommDestroy(comms[r]);
            }
            if (st != GDX_OK) {
                free_images();
                return st;
            }
            t_transport = "nccl";
            done = true;
        }
        for (int i = 0; i < n_devices; ++i) {
            if (done && member[i] >= 0) continue;
            // same device as the source or as an earlier target (or no NCCL): a plain copy
            DeviceGuard g(dev[i]);
            const void *from = idx->image;
            int from_dev = idx->device;
            if (done && member[i] < 0 && dev[i] != idx->device)
                for (int j = 0; j < i; ++j)
                    if (dev[j] == dev[i] && member[j] >= 0) {
                        from = img[j];
                        from_dev = dev[j];
                    }
            cudaError_t e = cudaMemcpyPeer(img[i], dev[i], from, from_dev, bytes);
            if (e != cudaSuccess) {
                free_images();
                return fail(GDX_ERR_CUDA, "copy to device %d failed: %s", dev[i], cudaGetErrorString(e));
            }
            if (!done && *t_transport == 0) t_transport = "peer";
        }
        if (*t_transport == 0) t_transport = "peer";
        for (int i = 0; i < n_devices; ++i) {
            gdx_status st = gdx_index_adopt_image(&idx->h, img[i], dev[i], 1, &out_replicas[i]);
            if (st != GDX_OK) {
                for (int j = 0; j < i; ++j) {
                    gdx_index_destroy(out_replicas[j]);
                    out_replicas[j] = nullptr;
                    img[j] = nullptr;
                }
                free_images();
                return st;
            }
            img[i] = nullptr;  // owned by the replica now
        }
        return GDX_OK;
    });
}

extern "C" gdx_status gdx_nccl_unique_id(void *id_out) {
    if (!id_out) return fail(GDX_ERR_BAD_ARG, "id_out is NULL");
    if (!nccl().ok) return fail(GDX_ERR_UNSUPPORTED, "NCCL is not available: %s", nccl().why.c_str());
    Nccl::unique_id id;
    NCCL_TRY(nccl().GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return GDX_OK;
}

extern "C" gdx_status gdx_index_broadcast(const gdx_index *src, const void *unique_id, int32_t rank, int32_t world,
                                          int32_t root, int32_t device_req, gdx_index **out) {
    return guarded([&]() -> gdx_status {
        if (!out || !unique_id || world < 1 || rank < 0 || rank >= world || root < 0 || root >= world)
            return fail(GDX_ERR_BAD_ARG, "bad argument");
        *out = nullptr;
        if ((rank == root) != (src != nullptr)) return fail(GDX_ERR_BAD_ARG, "src must be given on the root rank and only there");
        if (!nccl().ok) return fail(GDX_ERR_UNSUPPORTED, "NCCL is not available: %s", nccl().why.c_str());
        const Nccl &N = nccl();
        int device;
        GDX_TRY(resolve_device(rank == root ? src->device : device_req, &device));
        DeviceGuard guard(device);
        Nccl::unique_id id;
        memcpy(&id, unique_id, sizeof id);
        Nccl::comm_t comm = nullptr;
        NCCL_TRY(N.CommInitRank(&comm, world, id, rank));
        cudaStream_t stream = nullptr;
        void *d_hdr = nullptr, *image = nullptr;
        ImageHeader h;
        gdx_status st = [&]() -> gdx_status {
            CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaMalloc(&d_hdr, sizeof(ImageHeader)));
            if (rank == root) CUDA_TRY(cudaMemcpyAsync(d_hdr, &src->h, sizeof(ImageHeader), cudaMemcpyHostToDevice, stream));
            NCCL_TRY(N.Broadcast(d_hdr, d_hdr, sizeof(ImageHeader), Nccl::kUint8, root, comm, stream));
            CUDA_TRY(cudaMemcpyAsync(&h, d_hdr, sizeof(ImageHeader), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            GDX_TRY(validate_header(h));
            if (rank == root) image = src->image;
            else CUDA_TRY(cudaMalloc(&image, h.image_bytes ? h.image_bytes : 1));
            for (uint64_t off = 0; off < h.image_bytes; off += kBroadcastPiece) {
                const uint64_t nb = std::min(kBroadcastPiece, h.image_bytes - off);
                NCCL_TRY(N.Broadcast((uint8_t *)image + off, (uint8_t *)image + off, nb, Nccl::kUint8, root, comm, stream));
            }
            CUDA_TRY(cudaStreamSynchronize(stream));
            return GDX_OK;
        }();
        if (d_hdr) cudaFree(d_hdr);
        if (stream) cudaStreamDestroy(stream);
        N.CommDestroy(comm);
        if (st != GDX_OK) {
            if (rank != root && image) cudaFree(image);
            return st;
        }
        t_transport = "nccl";
        if (rank == root) {
            *out = const_cast<gdx_index *>(src);
            return GDX_OK;
        }
        st = gdx_index_adopt_image(&h, image, device, 1, out);
        if (st != GDX_OK) cudaFree(image);
        return st;
    });
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

// mode 0: cursors (starts, ends); 1: counts; 2: locate intervals (interval or resolved hit);
// narrow: results are written as uint32 (modes 0 and 1, texts shorter than 2^32)
template <class L>
void launch_search(const gdx_index *idx, const DevQueries &dq, uint64_t *a, uint64_t *b, int mode, bool narrow,
                   uint64_t qbase, uint64_t *err, unsigned long long *steps, const uint32_t *perm,
                   const uint32_t *slot_map, cudaStream_t stream) {
    if (dq.nq == 0) return;
    const unsigned grid = (unsigned)div_up(dq.nq, 256);
    const bool verify = idx->dev.text && idx->verify;
    const int mf = mode | (narrow ? kModeNarrow : 0);
#define GDX_LAUNCH(V, C, P) k_search<L, V, C, P><<<grid, 256, 0, stream>>>(idx->dev, dq, a, b, mf, qbase, err, steps, perm, slot_map)
    if (dq.packed) {
        if (verify && mode != 0) GDX_LAUNCH(true, false, true);
        else if (verify && idx->dev.isa) GDX_LAUNCH(true, true, true);
        else GDX_LAUNCH(false, false, true);
    } else {
        if (verify && mode != 0) GDX_LAUNCH(true, false, false);
        else if (verify && idx->dev.isa) GDX_LAUNCH(true, true, false);
        else GDX_LAUNCH(false, false, false);
    }
#undef GDX_LAUNCH
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

gdx_status check_queries(const gdx_index *idx, const gdx_queries *q) {
    if (!q) return fail(GDX_ERR_BAD_ARG, "queries is NULL");
    if (q->offsets && q->nq && q->offsets[q->nq] < q->offsets[0])
        return fail(GDX_ERR_BAD_ARG, "queries->offsets must be non-decreasing");
    if (q->nq && !q->offsets && q->fixed_len && !q->bytes) return fail(GDX_ERR_BAD_ARG, "queries->bytes is NULL");
    if (q->encoding > GDX_QUERIES_PACKED_2BIT) return fail(GDX_ERR_BAD_ARG, "unknown queries->encoding %u", q->encoding);
    if (q->encoding == GDX_QUERIES_PACKED_2BIT && idx && idx->h.ns > 4)
        return fail(GDX_ERR_UNSUPPORTED, "2-bit packed queries need an alphabet with at most 4 searchable symbols (this one has %u)",
                    idx->h.ns);
    return GDX_OK;
}

// index of the first symbol of query i in the batch's symbol stream (IO bytes or 2-bit codes)
uint64_t query_bytes_end(const gdx_queries *q, uint64_t i) {
    return q->offsets ? q->offsets[i] : i * q->fixed_len + q->first_symbol;
}

bool env_flag(const char *name, bool dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) != 0 : dflt;
}
// host-side 2-bit packing of large IO-byte batches (GDX_PACK_QUERIES=0 switches it off for A/B runs)
bool pack_enabled() {
    static const bool on = env_flag("GDX_PACK_QUERIES", true);
    return on;
}
// results of large batches cross PCIe as uint32 when the text is shorter than 2^32 (GDX_NARROW_RESULTS=0: off)
bool narrow_enabled() {
    static const bool on = env_flag("GDX_NARROW_RESULTS", true);
    return on;
}
const uint64_t kPackMinBytes = env_bytes("GDX_PACK_MIN_BYTES", 1ull << 20);

gdx_status acquire_pinned_hits(const gdx_index *idx, uint64_t bytes, void **out, uint64_t *cap_out);
void release_pinned_hits(const gdx_index *idx, void *p);

// K3 + K4 over `n_hits` rows.  With a sampled suffix array (walks of geometric length) the compacting kernel
// keeps every lane busy; with the dense array (one load per row) the plain one-thread-per-hit kernel is the
// shorter program.  GDX_LOCATE_COMPACT=0/1 forces either for A/B runs.
template <class L>
void launch_locate_walk(const gdx_index *idx, const uint64_t *rows, uint64_t n_hits, ulonglong2 *hits,
                        unsigned long long *d_walk, cudaStream_t stream, bool hit32 = false) {
    static const int forced = getenv("GDX_LOCATE_COMPACT") ? atoi(getenv("GDX_LOCATE_COMPACT")) : -1;
    const bool compacting = forced >= 0 ? forced != 0 : idx->dev.sampling_rate > 1;
    if (compacting)
        k_locate_walk_compact<L><<<(unsigned)div_up(n_hits, kWalkSlice), 256, 0, stream>>>(idx->dev, rows, n_hits, hits, d_walk,
                                                                                           hit32 ? 1 : 0);
    else
        k_locate_walk<L><<<(unsigned)div_up(n_hits, 256), 256, 0, stream>>>(idx->dev, rows, n_hits, hits, d_walk, hit32 ? 1 : 0);
}

// State of a pipelined gdx_locate_many: every chunk of the search pipeline continues, on its own
// stream, with counts -> scan -> expand -> walk -> D2H of its hits, so that the locate kernels and the
// hit copies of chunk k overlap the query upload of the later chunks.
struct LocatePipe {
    const gdx_index *idx;
    Workspace *ws;
    uint64_t *hit_offsets;  // caller's array, nq + 1 entries (NULL in the compact form)
    uint32_t *hit_counts = nullptr;  // compact form: hits per query, nq entries, instead of the CSR offsets
    uint64_t hit_bytes = sizeof(gdx_hit);  // 16, or 8 for gdx_hit32
    void *pinned = nullptr; // library-owned pinned hit buffer (grows)
    uint64_t pinned_cap = 0;
    uint64_t total = 0;     // hits of all finished chunks
    struct Pending {
        int slot;
        uint64_t q0, cq;
        bool valid = false;
    } pending_offsets;
    // chunks whose hit totals are on their way to the host, oldest first.  A chunk is finished two chunks after
    // it was issued (kSlots - 1): by then its total has long arrived, so the issuing thread never waits for a
    // search kernel -- and its slot is only reused by the chunk after that.
    std::deque<Pending> pending;
    bool stage_offsets = false;  // the caller's hit_offsets array is pageable

    // hand the previous chunk's CSR offsets from the slot's pinned staging to the caller
    gdx_status flush_offsets() {
        if (!pending_offsets.valid) return GDX_OK;
        pending_offsets.valid = false;
        Slot &ps = ws->slot[pending_offsets.slot];
        CUDA_TRY(cudaEventSynchronize(ps.ev_out));
        if (hit_counts) HostPool::get().copy(hit_counts + pending_offsets.q0, ps.h_out_a.p, pending_offsets.cq * 4);
        else HostPool::get().copy(hit_offsets + pending_offsets.q0, ps.h_out_a.p, pending_offsets.cq * 8);
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
        pending.push_back(Pending{slot, q0, cq, true});
        t_stats.kernel_launches += 2;
        return GDX_OK;
    }

    // the oldest chunk once more than `keep` are waiting (keep = 0: the oldest one, if any)
    Pending take_pending(size_t keep) {
        Pending p;
        if (pending.size() > keep) {
            p = pending.front();
            pending.pop_front();
        }
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
            if ((base + n_hits) * hit_bytes > pinned_cap) {  // grow the pinned result buffer (rare)
                for (int s2 = 0; s2 < kSlots; ++s2) CUDA_TRY(cudaStreamSynchronize(ws->slot[s2].stream));
                void *bigger = nullptr;
                uint64_t cap = 0;
                GDX_TRY(acquire_pinned_hits(idx, 2 * (base + n_hits) * hit_bytes, &bigger, &cap));
                if (pinned) {
                    memcpy(bigger, pinned, base * hit_bytes);
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
            cudaEvent_t ev_begin = ws->next_locate_event(sl), ev_end = ws->next_locate_event(sl);
            CUDA_TRY(cudaEventRecord(ev_begin, sl.stream));
            k_expand_rows<<<(unsigned)div_up(cq, 256), 256, 0, sl.stream>>>(
                sl.out_a.as<uint64_t>(), sl.out_b.as<uint64_t>(), sl.local_off.as<uint64_t>(), cq,
                sl.rows.as<uint64_t>(), sl.big.as<uint64_t>(), reinterpret_cast<unsigned long long *>(sl.d_words + 1));
            if (nbig) {
                dim3 grid((unsigned)nbig, 32);
                k_expand_big_rows<<<grid, 256, 0, sl.stream>>>(sl.out_a.as<uint64_t>(), sl.out_b.as<uint64_t>(),
                                                               sl.local_off.as<uint64_t>(), sl.big.as<uint64_t>(),
                                                               sl.rows.as<uint64_t>());
            }
            unsigned long long *d_walk = reinterpret_cast<unsigned long long *>(ws->small.d + 10);
            GDX_TRY(dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
                launch_locate_walk<decltype(L)>(idx, sl.rows.as<uint64_t>(), n_hits, sl.hits.as<ulonglong2>(), d_walk, sl.stream,
                                                hit_bytes == 8);
                return GDX_OK;
            }));
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaEventRecord(ev_end, sl.stream));
            CUDA_TRY(cudaMemcpyAsync((uint8_t *)pinned + base * hit_bytes, sl.hits.p, n_hits * hit_bytes,
                                     cudaMemcpyDeviceToHost, sl.stream));
            t_stats.d2h_bytes += n_hits * hit_bytes;
            t_stats.kernel_launches += 2 + (nbig ? 1 : 0);
        }
        // per query: global CSR offsets (u64), or in the compact form just the number of hits (u32; the counts are
        // narrowed into the offsets buffer, which the expansion above no longer needs)
        const uint64_t per_q = hit_counts ? 4 : 8;
        if (hit_counts) k_narrow_u64<<<(unsigned)div_up(cq, 256), 256, 0, sl.stream>>>(sl.counts.as<uint64_t>(), cq, sl.local_off.as<uint32_t>());
        else k_add_base<<<(unsigned)div_up(cq, 256), 256, 0, sl.stream>>>(sl.local_off.as<uint64_t>(), cq, base);
        CUDA_TRY(cudaGetLastError());
        void *dst = hit_counts ? (void *)(hit_counts + q0) : (void *)(hit_offsets + q0);
        if (stage_offsets) {
            GDX_TRY(flush_offsets());  // frees the staging of the chunk before
            CUDA_TRY(sl.h_out_a.reserve(cq * per_q));
            CUDA_TRY(cudaMemcpyAsync(sl.h_out_a.p, sl.local_off.p, cq * per_q, cudaMemcpyDeviceToHost, sl.stream));
            CUDA_TRY(cudaEventRecord(sl.ev_out, sl.stream));
            pending_offsets = Pending{p.slot, q0, cq, true};
        } else {
            CUDA_TRY(cudaMemcpyAsync(dst, sl.local_off.p, cq * per_q, cudaMemcpyDeviceToHost, sl.stream));
        }
        t_stats.d2h_bytes += cq * per_q;
        t_stats.kernel_launches += 1;
        total += n_hits;
        return GDX_OK;
    }
};

// Chunked pipeline over kSlots streams: stage / pack -> H2D(queries) -> k_search -> D2H(results) per chunk.
// If dev_a/dev_b are given the results stay on the device (locate path) and nothing is copied back.
// out_elem: bytes per element of the caller's result arrays (8, or 4 for the *_u32 entry points).
gdx_status search_host(const gdx_index *idx, Workspace *ws, const gdx_queries *qs, void *out_a_v, void *out_b_v,
                       uint32_t out_elem, int mode, uint64_t *dev_a, uint64_t *dev_b, LocatePipe *lp = nullptr) {
    const uint64_t nq = qs->nq;
    uint8_t *out_a = (uint8_t *)out_a_v, *out_b = (uint8_t *)out_b_v;
    for (int s = 0; s < kSlots; ++s) ws->slot[s].ev_used = ws->slot[s].ev_loc_used = 0;
    // error words / counters are reset on slot 0's stream; the other slot streams (non-blocking: no implicit
    // ordering with anything) wait for that before their first kernel
    CUDA_TRY(cudaMemsetAsync(ws->small.d, 0xff, 4 * sizeof(uint64_t), ws->slot[0].stream));
    CUDA_TRY(cudaMemsetAsync(ws->small.d + 4, 0, 12 * sizeof(uint64_t), ws->slot[0].stream));
    CUDA_TRY(cudaEventRecord(ws->ev_a, ws->slot[0].stream));
    for (int s = 1; s < kSlots; ++s) CUDA_TRY(cudaStreamWaitEvent(ws->slot[s].stream, ws->ev_a, 0));
    unsigned long long *d_steps = reinterpret_cast<unsigned long long *>(ws->small.d + 4);
    static const bool trace = getenv("GDX_TRACE") && atoi(getenv("GDX_TRACE")) != 0;
    struct TraceRow {
        int k;
        uint64_t nq, bytes;
        cudaEvent_t h2d_begin, k_begin, k_end, d2h_end;
        double host_stage_ms;
    };
    std::vector<TraceRow> trace_rows;
    const auto t_host0 = std::chrono::steady_clock::now();
    const uint64_t sym_end = query_bytes_end(qs, nq);
    const uint64_t total_syms = sym_end - query_bytes_end(qs, 0);
    const bool prepacked = qs->encoding == GDX_QUERIES_PACKED_2BIT;
    const uint64_t total_in = prepacked ? (total_syms + 3) / 4 : total_syms;
    // Ordinary (pageable) caller memory would make every cudaMemcpyAsync a blocking, driver-staged copy at
    // 8-14 GB/s.  Large batches are staged through the slots' own pinned buffers instead, filled by the host
    // thread pool; while they are staged, IO bytes of a <= 4-symbol alphabet are packed to 2 bits (a quarter
    // of the PCIe bytes) -- that pays for pinned caller buffers too.
    bool pack_on = nq && !prepacked && idx->pack.usable && pack_enabled() && total_in >= kPackMinBytes;
    const bool src_pinned = nq && is_pinned(qs->bytes);
    const bool stage_in = nq && total_in >= kStageMinBytes && !src_pinned;
    // Pinned IO bytes can also cross the link as they are, by DMA, without any CPU work.  The host packer and
    // the link then share the batch: before the pool starts on a packed chunk (which keeps the CPU busy for
    // t_pack), a raw chunk is put on the link that is just large enough to keep it busy for that long on top of
    // what is still queued there.  The queue is observed (events behind every upload), so a wrong guess of
    // either rate corrects itself: too much raw data -> the backlog grows -> the next raw chunks shrink.
    // GDX_PACK_HYBRID=0: every chunk is packed.  GDX_LINK_GBS: initial guess of the link rate (default 50).
    static const bool hybrid_enabled = env_flag("GDX_PACK_HYBRID", true);
    static const double link_guess = (double)env_bytes("GDX_LINK_GBS", 50) * 1e6;  // bytes per ms
    const bool hybrid = pack_on && src_pinned && hybrid_enabled;
    struct Upload {
        int slot;
        uint64_t bytes;
    };
    std::vector<Upload> on_link;  // uploads issued, oldest first; retired when their event has fired
    auto link_backlog = [&]() -> uint64_t {
        size_t done = 0;
        while (done < on_link.size() && cudaEventQuery(ws->slot[on_link[done].slot].ev_link) == cudaSuccess) ++done;
        on_link.erase(on_link.begin(), on_link.begin() + done);
        uint64_t b = 0;
        for (const Upload &u : on_link) b += u.bytes;
        return b;
    };
    double pack_rate = 4.0e6 * HostPool::get().threads();  // symbols per ms, refined from every packed chunk
    bool raw_next = hybrid;                                  // the link is idle at the start: open with a raw chunk
    uint64_t raw_want = kChunkFirst;

    const bool stage_off = nq && qs->offsets && nq * 8 >= kStageMinBytes && !is_pinned(qs->offsets);
    const bool to_host = nq && !dev_a && !lp;
    const bool large_out = nq * 8 >= kStageMinBytes;
    // results as uint32 on the device: always for the u32 entry points, for large batches of the u64 ones
    // (widened by the pool while they are handed to the caller)
    // (pinned uint64 result arrays are filled by DMA without any CPU work: they stay 64-bit on the wire)
    const bool narrow = to_host && !idx->h.wide &&
                        (out_elem == 4 || (large_out && narrow_enabled() && !(is_pinned(out_a) && (mode != 0 || is_pinned(out_b)))));
    const bool widen = narrow && out_elem == 8;
    const bool stage_out = to_host && (widen || (large_out && !is_pinned(out_a)));
    const uint32_t dev_elem = narrow ? 4 : 8;
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
        if (widen) {
            HostPool::get().widen_u32((uint64_t *)out_a + p.q0, (const uint32_t *)ps.h_out_a.p, p.cq);
            if (mode == 0) HostPool::get().widen_u32((uint64_t *)out_b + p.q0, (const uint32_t *)ps.h_out_b.p, p.cq);
        } else {
            HostPool::get().copy(out_a + p.q0 * out_elem, ps.h_out_a.p, p.cq * out_elem);
            if (mode == 0) HostPool::get().copy(out_b + p.q0 * out_elem, ps.h_out_b.p, p.cq * out_elem);
        }
        return GDX_OK;
    };
    for (int s2 = 0; s2 < kSlots; ++s2) ws->slot[s2].h2d_pending = false;
    // Large batches: size every slot once for the largest chunk this call can cut (the adaptive schedule cuts
    // different chunks in every call; growing a buffer in mid-flight means cudaFree / cudaFreeHost + a new
    // allocation, milliseconds each, with the streams drained).  The streams are idle here.
    if (nq && total_in >= kStageMinBytes) {
        const uint64_t max_sym = std::min<uint64_t>(total_syms, 4 * kChunkMax) + 64;
        const uint64_t max_q = qs->fixed_len ? std::min<uint64_t>(nq, max_sym / qs->fixed_len + 1) : 0;
        for (int s2 = 0; s2 < kSlots; ++s2) {
            Slot &sl = ws->slot[s2];
            CUDA_TRY(sl.bytes.reserve(max_sym));
            if (pack_on) CUDA_TRY(sl.h_in.reserve(max_sym / 4 + 8));
            else if (stage_in) CUDA_TRY(sl.h_in.reserve(std::min<uint64_t>(total_in, kChunkMax) + 64));
            if (max_q && !dev_a) {
                CUDA_TRY(sl.out_a.reserve(max_q * 8));
                if (mode == 0 || lp) CUDA_TRY(sl.out_b.reserve(max_q * 8));
                if (stage_out) {
                    CUDA_TRY(sl.h_out_a.reserve(max_q * dev_elem));
                    if (mode == 0) CUDA_TRY(sl.h_out_b.reserve(max_q * dev_elem));
                }
            }
        }
    }
    std::vector<uint64_t> exc;       // symbol positions (chunk relative) the packer could not encode
    std::vector<uint32_t> exc_q;     // the queries (chunk relative) that hold them
    uint64_t packed_queries = 0, exception_queries = 0, h2d_bytes = 0, d2h_bytes = 0;
    uint64_t budget = kChunkFirst;
    uint64_t q0 = 0;
    int k = 0;
    while (q0 < nq) {
        // chunk [q0, q1): about `budget` input bytes (packed bytes count four-fold: the budget is PCIe time),
        // at least one query
        bool pack_chunk = pack_on;
        if (hybrid && pack_on && raw_next) pack_chunk = false;
        const uint64_t unit = (pack_chunk || prepacked) ? 4 : 1;
        const uint64_t remaining = sym_end - query_bytes_end(qs, q0);
        uint64_t want = (hybrid && pack_on && !pack_chunk) ? raw_want : budget * unit;
        if (remaining <= want) want = remaining > 2 * kChunkTail * unit ? remaining - kChunkTail * unit : remaining;
        uint64_t q1;
        if (qs->offsets) {
            const uint64_t *ob = qs->offsets + q0 + 1, *oe = qs->offsets + nq + 1;
            const uint64_t *it = std::upper_bound(ob, oe, qs->offsets[q0] + want);
            q1 = q0 + (uint64_t)(it - ob);
            if (q1 == q0) q1 = q0 + 1;
        } else {
            q1 = q0 + (qs->fixed_len ? std::max<uint64_t>(1, want / qs->fixed_len) : kChunkMaxQueries);
        }
        q1 = std::min<uint64_t>(q1, std::min<uint64_t>(nq, q0 + kChunkMaxQueries));
        if (pack_chunk || !(hybrid && pack_on)) budget = std::min<uint64_t>(budget * 2, kChunkMax);
        const uint64_t cq = q1 - q0;
        const uint64_t sym0 = query_bytes_end(qs, q0), sym1 = query_bytes_end(qs, q1), nsym = sym1 - sym0;
        if (sym1 < sym0 || sym1 > sym_end) {  // (inside a chunk the kernel checks every query against the chunk's extent)
            t_error_query = q0;
            return fail(GDX_ERR_BAD_ARG, "queries->offsets must not decrease and must stay inside the batch (near query %llu)",
                        (unsigned long long)q0);
        }
        Slot &sl = ws->slot[k % kSlots];
        const SortPlan sp = plan_sort(idx, cq, qs->offsets ? 0 : qs->fixed_len);
        // growing a slot buffer frees the old one: only safe once the slot's stream has drained
        const uint64_t in_cap = nsym + 64;  // enough for either representation of the chunk
        if (sl.bytes.cap < in_cap || (qs->offsets && sl.offsets.cap < (cq + 1) * 8) ||
            (!dev_a && (sl.out_a.cap < cq * 8 || ((mode == 0 || lp) && sl.out_b.cap < cq * 8))) ||
            (sp.use && sl.sort.cap < sp.total_bytes))
            CUDA_TRY(cudaStreamSynchronize(sl.stream));
        CUDA_TRY(sl.bytes.reserve(in_cap));
        cudaEvent_t ev_h2d = nullptr;
        if (trace) {
            cudaEventCreate(&ev_h2d);
            cudaEventRecord(ev_h2d, sl.stream);
        }
        if (sl.h2d_pending) {  // the slot's staging buffers are free again once their last upload is done
            CUDA_TRY(cudaEventSynchronize(sl.ev_h2d));
            sl.h2d_pending = false;
        }
        const auto t_stage0 = std::chrono::steady_clock::now();
        DevQueries dq;
        dq.bytes = sl.bytes.as<uint8_t>();
        dq.packed = nullptr;
        dq.offsets = nullptr;
        dq.offsets32 = nullptr;
        dq.fixed_len = qs->fixed_len;
        dq.nq = cq;
        dq.base = 0;
        dq.shift = 0;
        dq.limit = nsym;  // (+ shift once a packed chunk decides it)
        // large variable-length batches: offsets cross PCIe chunk relative as uint32
        const bool narrow_off = qs->offsets && nq * 8 >= kStageMinBytes && nsym + 4 < 0xffffffffull && narrow_enabled();
        bool used_staging = false;
        uint64_t nx = 0;  // queries of this chunk that go through the IO-byte kernel after the packed one
        bool chunk_packed = false;
        if (nsym && pack_chunk) {
            CUDA_TRY(sl.h_in.reserve((nsym + 3) / 4 + 8));
            pack2_parallel(idx->pack, qs->bytes + sym0, nsym, (uint8_t *)sl.h_in.p, exc);
            exc_q.clear();
            for (uint64_t pos : exc) {
                uint64_t q;
                if (qs->offsets) {
                    const uint64_t *ob = qs->offsets + q0, *oe = qs->offsets + q1 + 1;
                    q = (uint64_t)(std::upper_bound(ob, oe, sym0 + pos) - ob) - 1;
                    q = std::min(q, cq - 1);  // (offsets that are not sorted are reported by the kernel; stay in the chunk)
                } else {
                    q = pos / qs->fixed_len;
                }
                if (exc_q.empty() || exc_q.back() != (uint32_t)q) exc_q.push_back((uint32_t)q);
            }
            if (exc_q.size() * 16 > cq) {
                // this is not a batch of plain searchable symbols: IO bytes from here on
                pack_on = false;
            } else {
                chunk_packed = true;
                nx = exc_q.size();
                CUDA_TRY(cudaMemcpyAsync(sl.bytes.p, sl.h_in.p, (nsym + 3) / 4, cudaMemcpyHostToDevice, sl.stream));
                h2d_bytes += (nsym + 3) / 4;
                dq.bytes = nullptr;
                dq.packed = sl.bytes.as<uint32_t>();
                used_staging = true;
            }
        }
        if (nsym && !chunk_packed) {
            if (prepacked) {  // the caller's stream, from the byte that holds the chunk's first symbol
                const uint64_t b0 = sym0 / 4, b1 = (sym1 + 3) / 4;
                const uint8_t *src = qs->bytes + b0;
                if (stage_in) {
                    CUDA_TRY(sl.h_in.reserve(b1 - b0));
                    HostPool::get().copy(sl.h_in.p, src, b1 - b0);
                    src = (const uint8_t *)sl.h_in.p;
                    used_staging = true;
                }
                CUDA_TRY(cudaMemcpyAsync(sl.bytes.p, src, b1 - b0, cudaMemcpyHostToDevice, sl.stream));
                h2d_bytes += b1 - b0;
                dq.bytes = nullptr;
                dq.packed = sl.bytes.as<uint32_t>();
                dq.shift = sym0 & 3;
                dq.limit = nsym + dq.shift;
            } else {
                const uint8_t *src = qs->bytes + sym0;
                if (stage_in) {
                    CUDA_TRY(sl.h_in.reserve(nsym));
                    HostPool::get().copy(sl.h_in.p, src, nsym);
                    src = (const uint8_t *)sl.h_in.p;
                    used_staging = true;
                }
                CUDA_TRY(cudaMemcpyAsync(sl.bytes.p, src, nsym, cudaMemcpyHostToDevice, sl.stream));
                h2d_bytes += nsym;
            }
        } else if (!nsym && prepacked) {
            dq.bytes = nullptr;
            dq.packed = sl.bytes.as<uint32_t>();
        }
        if (qs->offsets && narrow_off) {
            // chunk-relative 32-bit offsets, narrowed by the pool into pinned staging: half the PCIe bytes
            CUDA_TRY(sl.offsets.reserve((cq + 1) * 8));
            CUDA_TRY(sl.h_off.reserve((cq + 1) * 4));
            const uint64_t *osrc = qs->offsets + q0;
            uint32_t *o32 = (uint32_t *)sl.h_off.p;
            const uint64_t rel = sym0 - dq.shift, cnt = cq + 1;
            constexpr uint64_t kPiece = 1ull << 18;
            HostPool::get().parallel_for(div_up(cnt, kPiece), [&](uint64_t piece) {
                const uint64_t lo = piece * kPiece, hi = std::min(cnt, lo + kPiece);
                for (uint64_t i = lo; i < hi; ++i) o32[i] = (uint32_t)(osrc[i] - rel);
            });
            used_staging = true;
            CUDA_TRY(cudaMemcpyAsync(sl.offsets.p, o32, cnt * 4, cudaMemcpyHostToDevice, sl.stream));
            h2d_bytes += cnt * 4;
            dq.offsets32 = sl.offsets.as<uint32_t>();
            dq.shift = 0;
        } else if (qs->offsets) {
            CUDA_TRY(sl.offsets.reserve((cq + 1) * 8));
            const uint64_t *osrc = qs->offsets + q0;
            if (stage_off) {
                CUDA_TRY(sl.h_off.reserve((cq + 1) * 8));
                HostPool::get().copy(sl.h_off.p, osrc, (cq + 1) * 8);
                osrc = (const uint64_t *)sl.h_off.p;
                used_staging = true;
            }
            CUDA_TRY(cudaMemcpyAsync(sl.offsets.p, osrc, (cq + 1) * 8, cudaMemcpyHostToDevice, sl.stream));
            h2d_bytes += (cq + 1) * 8;
            dq.offsets = sl.offsets.as<uint64_t>();
            dq.base = sym0 - dq.shift;
            dq.shift = 0;
        }
        // queries with an exception byte: their offsets, result slots and IO bytes in one upload
        DevQueries xq = {};
        const uint32_t *x_slots = nullptr;
        if (nx) {
            uint64_t xbytes = 0;
            for (uint32_t q : exc_q) xbytes += query_bytes_end(qs, q0 + q + 1) - query_bytes_end(qs, q0 + q);
            const uint64_t off_slots = (nx + 1) * 8, off_bytes = align_up(off_slots + nx * 4, 8), tot = off_bytes + xbytes + 8;
            if (sl.xbuf.cap < tot) CUDA_TRY(cudaStreamSynchronize(sl.stream));
            CUDA_TRY(sl.xbuf.reserve(tot));
            CUDA_TRY(sl.h_x.reserve(tot));
            uint64_t *xo = (uint64_t *)sl.h_x.p;
            uint32_t *xs = (uint32_t *)((uint8_t *)sl.h_x.p + off_slots);
            uint8_t *xb = (uint8_t *)sl.h_x.p + off_bytes;
            uint64_t acc = 0;
            for (uint64_t i = 0; i < nx; ++i) {
                xo[i] = acc;
                xs[i] = exc_q[i];
                acc += query_bytes_end(qs, q0 + exc_q[i] + 1) - query_bytes_end(qs, q0 + exc_q[i]);
            }
            xo[nx] = acc;
            constexpr uint64_t kPiece = 4096;  // queries per piece
            HostPool::get().parallel_for(div_up(nx, kPiece), [&](uint64_t piece) {
                const uint64_t lo = piece * kPiece, hi = std::min(nx, lo + kPiece);
                for (uint64_t i = lo; i < hi; ++i) {
                    const uint64_t b0 = query_bytes_end(qs, q0 + exc_q[i]);
                    memcpy(xb + xo[i], qs->bytes + b0, xo[i + 1] - xo[i]);
                }
            });
            CUDA_TRY(cudaMemcpyAsync(sl.xbuf.p, sl.h_x.p, off_bytes + xbytes, cudaMemcpyHostToDevice, sl.stream));
            h2d_bytes += off_bytes + xbytes;
            xq.bytes = sl.xbuf.as<uint8_t>() + off_bytes;
            xq.offsets = sl.xbuf.as<uint64_t>();
            xq.nq = nx;
            xq.limit = xbytes;
            x_slots = reinterpret_cast<const uint32_t *>(sl.xbuf.as<uint8_t>() + off_slots);
            used_staging = true;
        }
        const double stage_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_stage0).count();
        if (hybrid) {
            // this slot's event is about to be re-recorded: its older upload has long left the link
            for (size_t i = 0; i < on_link.size();)
                if (on_link[i].slot == k % kSlots) on_link.erase(on_link.begin() + i);
                else ++i;
            CUDA_TRY(cudaEventRecord(sl.ev_link, sl.stream));
            on_link.push_back(Upload{k % kSlots, chunk_packed ? (nsym + 3) / 4 : nsym});
            if (chunk_packed && stage_ms > 0.02) pack_rate = 0.5 * pack_rate + 0.5 * (double)nsym / stage_ms;
            if (chunk_packed) {
                // how long will the pool need for the next packed chunk, and will the link last that long?
                const double t_pack = (double)(budget * 4) / pack_rate;
                const double have = (double)link_backlog();
                const double need = t_pack * link_guess;
                raw_next = pack_on && need - have >= (double)(2ull << 20);
                raw_want = raw_next ? std::min<uint64_t>((uint64_t)(need - have), 4 * kChunkMax) : 0;
            } else {
                raw_next = false;  // a raw chunk is always followed by a packed one
            }
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
        if (used_staging) {
            CUDA_TRY(cudaEventRecord(sl.ev_h2d, sl.stream));
            sl.h2d_pending = true;
        }
        cudaEvent_t e0 = ws->next_event(sl), e1 = ws->next_event(sl);
        CUDA_TRY(cudaEventRecord(e0, sl.stream));
        if (trace) trace_rows.push_back(TraceRow{k, cq, chunk_packed || prepacked ? (nsym + 3) / 4 : nsym, ev_h2d, e0, e1, nullptr, stage_ms});
        const uint32_t *perm = nullptr;
        if (sp.use) {
            CUDA_TRY(sl.sort.reserve(sp.total_bytes));
            GDX_TRY(sort_queries(idx, dq, sp, sl.sort.p, sl.stream, &perm));
            t_stats.kernel_launches += 2;
        }
        gdx_status st = dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
            launch_search<decltype(L)>(idx, dq, a, b, mode, narrow, q0, ws->small.d + (k % kSlots), d_steps, perm, nullptr,
                                       sl.stream);
            if (nx)  // overwrites the slots of the queries the packed kernel could not see correctly
                launch_search<decltype(L)>(idx, xq, a, b, mode, narrow, q0, ws->small.d + (k % kSlots), d_steps, nullptr,
                                           x_slots, sl.stream);
            return GDX_OK;
        });
        GDX_TRY(st);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(e1, sl.stream));
        if (chunk_packed || prepacked) packed_queries += cq;
        exception_queries += nx;
        if (lp) {  // locate continues per chunk; the previous chunk is finished while this one runs
            GDX_TRY(lp->stage_counts(sl, k % kSlots, q0, cq));
            GDX_TRY(lp->finish(lp->take_pending(kSlots - 1)));
        } else if (stage_out) {  // results go to pinned staging; the previous chunk's are handed over meanwhile
            PendingOut prev = pend_out;
            CUDA_TRY(sl.h_out_a.reserve(cq * dev_elem));
            CUDA_TRY(cudaMemcpyAsync(sl.h_out_a.p, a, cq * dev_elem, cudaMemcpyDeviceToHost, sl.stream));
            if (mode == 0) {
                CUDA_TRY(sl.h_out_b.reserve(cq * dev_elem));
                CUDA_TRY(cudaMemcpyAsync(sl.h_out_b.p, b, cq * dev_elem, cudaMemcpyDeviceToHost, sl.stream));
            }
            CUDA_TRY(cudaEventRecord(sl.ev_out, sl.stream));
            pend_out = PendingOut{k % kSlots, q0, cq, true};
            d2h_bytes += cq * dev_elem * (mode == 0 ? 2 : 1);
            GDX_TRY(finish_out(prev));
        } else if (!dev_a) {
            d2h_bytes += cq * out_elem * (mode == 0 ? 2 : 1);
            CUDA_TRY(cudaMemcpyAsync(out_a + q0 * out_elem, a, cq * out_elem, cudaMemcpyDeviceToHost, sl.stream));
            if (mode == 0) CUDA_TRY(cudaMemcpyAsync(out_b + q0 * out_elem, b, cq * out_elem, cudaMemcpyDeviceToHost, sl.stream));
        }
        if (trace) {
            cudaEventCreate(&trace_rows.back().d2h_end);
            cudaEventRecord(trace_rows.back().d2h_end, sl.stream);
        }
        t_stats.kernel_launches += 1 + (nx ? 1 : 0);
        q0 = q1;
        ++k;
    }
    while (lp && !lp->pending.empty()) GDX_TRY(lp->finish(lp->take_pending(0)));
    if (lp) GDX_TRY(lp->flush_offsets());
    GDX_TRY(finish_out(pend_out));
    const double t_issue = trace ? std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count() : 0;
    for (int s = 0; s < kSlots; ++s) CUDA_TRY(cudaStreamSynchronize(ws->slot[s].stream));
    if (trace && !trace_rows.empty()) {
        const double t_sync = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
        fprintf(stderr, "[gdx trace] host: all chunks issued at %.3f ms, streams drained at %.3f ms (%u pool threads)\n", t_issue,
                t_sync, HostPool::get().threads());
        cudaEvent_t base = trace_rows[0].h2d_begin;
        for (auto &r : trace_rows) {
            float a = 0, b = 0, c = 0, d = 0;
            cudaError_t te = cudaEventElapsedTime(&a, base, r.h2d_begin);
            if (te != cudaSuccess) fprintf(stderr, "[gdx trace] elapsed: %s\n", cudaGetErrorString(te));
            cudaEventElapsedTime(&b, base, r.k_begin);
            cudaEventElapsedTime(&c, base, r.k_end);
            cudaEventElapsedTime(&d, base, r.d2h_end);
            fprintf(stderr, "[gdx trace] chunk %2d slot %d nq %8llu h2d bytes %10llu host stage %.3f ms | h2d %.3f..%.3f | kernels %.3f..%.3f | d2h ..%.3f\n",
                    r.k, r.k % kSlots, (unsigned long long)r.nq, (unsigned long long)r.bytes, r.host_stage_ms, a, b, b, c, d);
        }
        for (auto &r : trace_rows) {
            cudaEventDestroy(r.h2d_begin);
            cudaEventDestroy(r.d2h_end);
        }
    }
    CUDA_TRY(cudaMemcpy(ws->small.h, ws->small.d, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    double ms = 0, ms_loc = 0;
    for (int s = 0; s < kSlots; ++s) {
        for (size_t i = 0; i + 1 < ws->slot[s].ev_used; i += 2) {
            float t = 0;
            cudaEventElapsedTime(&t, ws->slot[s].ev[i], ws->slot[s].ev[i + 1]);
            ms += t;
        }
        for (size_t i = 0; i + 1 < ws->slot[s].ev_loc_used; i += 2) {
            float t = 0;
            cudaEventElapsedTime(&t, ws->slot[s].ev_loc[i], ws->slot[s].ev_loc[i + 1]);
            ms_loc += t;
        }
    }
    t_stats.queries = nq;
    t_stats.lf_steps = ws->small.h[4];
    t_stats.walk_steps = ws->small.h[5] + ws->small.h[10];
    t_stats.locate_walk_steps = ws->small.h[10];
    t_stats.verified_queries = ws->small.h[9];
    t_stats.kernel_ms_search = ms;
    t_stats.kernel_ms_locate = ms_loc;
    t_stats.packed_queries = packed_queries;
    t_stats.exception_queries = exception_queries;
    t_stats.h2d_bytes = h2d_bytes;
    t_stats.d2h_bytes += d2h_bytes;
    uint64_t bad = kNoError;
    for (int s = 0; s < kSlots; ++s) bad = std::min(bad, ws->small.h[s]);
    if (bad != kNoError && (bad & kBadOffsetFlag)) {
        t_error_query = bad & ~kBadOffsetFlag;
        return fail(GDX_ERR_BAD_ARG, "query %llu: queries->offsets must not decrease and must stay inside the batch",
                    (unsigned long long)t_error_query);
    }
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
        unsigned long long *d_walk = reinterpret_cast<unsigned long long *>(ws->small.d + 10);
        gdx_status s2 = dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
            launch_locate_walk<decltype(L)>(idx, ws->rows.as<uint64_t>(), total, ws->hits.as<ulonglong2>(), d_walk, st);
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
            cudaGetLastError();
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
                         gdx_hit **hits, uint64_t *num_hits, bool write_last = true) {
    cudaStream_t st = ws->slot[0].stream;
    void *h = nullptr;
    GDX_TRY(acquire_pinned_hits(idx, total * sizeof(gdx_hit), &h));
    const gdx_status copied = [&]() -> gdx_status {
        if (total) CUDA_TRY(cudaMemcpyAsync(h, ws->hits.p, total * sizeof(gdx_hit), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(hit_offsets, ws->hit_offsets.p, (n + (write_last ? 1 : 0)) * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(ws->small.h, ws->small.d, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return GDX_OK;
    }();
    if (copied != GDX_OK) {
        cudaStreamSynchronize(st);
        release_pinned_hits(idx, h);
        return copied;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, ws->ev_a, ws->ev_b);
    t_stats.kernel_ms_locate = ms;
    t_stats.hits = total;
    t_stats.walk_steps += ws->small.h[10];
    t_stats.locate_walk_steps = ws->small.h[10];
    *hits = (gdx_hit *)h;
    *num_hits = total;
    return GDX_OK;
}

}  // namespace

namespace {

// a failed call must not return while one of its streams still reads or writes the caller's buffers; the error that
// made it fail stays the reported one
void drain(Workspace *ws) {
    for (int s = 0; s < kSlots; ++s) cudaStreamSynchronize(ws->slot[s].stream);
    cudaGetLastError();
}

// mode 0: cursors, 1: counts; elem: bytes per element of the caller's arrays
gdx_status search_many_impl(const gdx_index *idx, const gdx_queries *queries, void *a, void *b, uint32_t elem, int mode,
                            const char *what) {
    GDX_TRY(begin_call(idx, what));
    GDX_TRY(check_queries(idx, queries));
    if (queries->nq && (!a || (mode == 0 && !b))) return fail(GDX_ERR_BAD_ARG, "output is NULL");
    if (elem == 4 && idx->h.wide) return fail(GDX_ERR_UNSUPPORTED, "32-bit results need a text shorter than 2^32 symbols");
    std::shared_lock<std::shared_mutex> cfg(idx->cfg_mu);
    DeviceGuard guard(idx->device);
    WsLease lease(idx);
    if (!lease.w) return fail(GDX_ERR_CUDA, "could not create a CUDA workspace: %s", cudaGetErrorString(cudaGetLastError()));
    const gdx_status st = search_host(idx, lease.w, queries, a, b, elem, mode, nullptr, nullptr);
    if (st != GDX_OK) drain(lease.w);  // copies into the caller's arrays may still be in flight
    return st;
}

// write_last = false: hit_offsets[nq] is left alone (it is the first entry of the next shard's range)
// hit_counts != NULL: the compact form (gdx_hit32 hits, u32 hits per query, no offsets)
gdx_status locate_many_impl(const gdx_index *idx, const gdx_queries *queries, uint64_t *hit_offsets, gdx_hit **hits,
                            uint64_t *num_hits, bool write_last = true, uint32_t *hit_counts = nullptr) {
    GDX_TRY(begin_call(idx, "gdx_locate_many"));
    GDX_TRY(check_queries(idx, queries));
    if ((!hit_offsets && !hit_counts) || !hits || !num_hits) return fail(GDX_ERR_BAD_ARG, "output is NULL");
    *hits = nullptr;
    *num_hits = 0;
    if (hit_counts && (idx->h.wide || idx->h.ntexts > 0xffffffffull))
        return fail(GDX_ERR_UNSUPPORTED, "32-bit hits need a text shorter than 2^32 symbols");
    std::shared_lock<std::shared_mutex> cfg(idx->cfg_mu);
    DeviceGuard guard(idx->device);
    WsLease lease(idx);
    Workspace *ws = lease.w;
    if (!ws) return fail(GDX_ERR_CUDA, "could not create a CUDA workspace: %s", cudaGetErrorString(cudaGetLastError()));
    const uint64_t n = queries->nq;
    static const bool pipelined_env = !(getenv("GDX_LOCATE_PIPELINE") && atoi(getenv("GDX_LOCATE_PIPELINE")) == 0);
    const bool pipelined = pipelined_env || hit_counts;  // (the compact form exists in the pipelined path only)
    if (pipelined) {
        // every chunk of the search pipeline carries on with counts -> scan -> expand -> walk -> D2H
        LocatePipe lp{idx, ws, hit_offsets};
        lp.hit_counts = hit_counts;
        lp.hit_bytes = hit_counts ? 8 : sizeof(gdx_hit);
        lp.stage_offsets = n * 8 >= kStageMinBytes && !is_pinned(hit_counts ? (const void *)hit_counts : (const void *)hit_offsets);
        GDX_TRY(acquire_pinned_hits(idx, std::max<uint64_t>(n, 4096) * lp.hit_bytes, &lp.pinned, &lp.pinned_cap));
        gdx_status st = search_host(idx, ws, queries, nullptr, nullptr, 8, 2, nullptr, nullptr, &lp);
        if (st != GDX_OK) {
            drain(ws);
            release_pinned_hits(idx, lp.pinned);
            return st;
        }
        if (write_last && hit_offsets) hit_offsets[n] = lp.total;
        t_stats.hits = lp.total;
        *hits = (gdx_hit *)lp.pinned;
        *num_hits = lp.total;
        return GDX_OK;
    }
    // a buffer may only be replaced once nothing in flight uses it: every call ends synchronized
    CUDA_TRY(ws->starts.reserve((n + 1) * 8));
    CUDA_TRY(ws->ends.reserve((n + 1) * 8));
    gdx_status st = search_host(idx, ws, queries, nullptr, nullptr, 8, 2, ws->starts.as<uint64_t>(), ws->ends.as<uint64_t>());
    uint64_t total = 0;
    if (st == GDX_OK) st = locate_device_intervals(idx, ws, ws->starts.as<uint64_t>(), ws->ends.as<uint64_t>(), n, &total);
    if (st == GDX_OK) st = finish_locate(idx, ws, n, total, hit_offsets, hits, num_hits, write_last);
    if (st != GDX_OK) drain(ws);
    return st;
}

}  // namespace

extern "C" gdx_status gdx_cursors_many(const gdx_index *idx, const gdx_queries *queries, uint64_t *starts,
                                       uint64_t *ends) {
    return guarded([&] { return search_many_impl(idx, queries, starts, ends, 8, 0, "gdx_cursors_many"); });
}
extern "C" gdx_status gdx_cursors_many_u32(const gdx_index *idx, const gdx_queries *queries, uint32_t *starts,
                                           uint32_t *ends) {
    return guarded([&] { return search_many_impl(idx, queries, starts, ends, 4, 0, "gdx_cursors_many_u32"); });
}
extern "C" gdx_status gdx_count_many(const gdx_index *idx, const gdx_queries *queries, uint64_t *counts) {
    return guarded([&] { return search_many_impl(idx, queries, counts, nullptr, 8, 1, "gdx_count_many"); });
}
extern "C" gdx_status gdx_count_many_u32(const gdx_index *idx, const gdx_queries *queries, uint32_t *counts) {
    return guarded([&] { return search_many_impl(idx, queries, counts, nullptr, 4, 1, "gdx_count_many_u32"); });
}
extern "C" gdx_status gdx_locate_many(const gdx_index *idx, const gdx_queries *queries, uint64_t *hit_offsets,
                                      gdx_hit **hits, uint64_t *num_hits) {
    return guarded([&] { return locate_many_impl(idx, queries, hit_offsets, hits, num_hits); });
}

extern "C" gdx_status gdx_locate_many_compact(const gdx_index *idx, const gdx_queries *queries, uint32_t *hit_counts,
                                              gdx_hit32 **hits, uint64_t *num_hits) {
    return guarded([&]() -> gdx_status {
        if (!hit_counts) return fail(GDX_ERR_BAD_ARG, "output is NULL");
        return locate_many_impl(idx, queries, nullptr, reinterpret_cast<gdx_hit **>(hits), num_hits, true, hit_counts);
    });
}

extern "C" uint64_t gdx_packed_bytes(uint64_t total_symbols) { return align_up((total_symbols + 3) / 4, 4); }

extern "C" gdx_status gdx_pack_queries(const gdx_index *idx, const gdx_queries *q, uint8_t *packed_out,
                                       uint64_t *first_unencodable) {
    return guarded([&]() -> gdx_status {
        if (!idx || !packed_out) return fail(GDX_ERR_BAD_ARG, "NULL argument");
        GDX_TRY(check_queries(idx, q));
        if (q->encoding != GDX_QUERIES_IO_BYTES) return fail(GDX_ERR_BAD_ARG, "the batch is already packed");
        if (!idx->pack.usable)
            return fail(GDX_ERR_UNSUPPORTED, "2-bit packing needs an alphabet with at most 4 searchable symbols (this one has %u)",
                        idx->h.ns);
        const uint64_t s0 = query_bytes_end(q, 0), s1 = query_bytes_end(q, q->nq);
        // symbols in front of the first query keep their place in the stream so that offsets stay valid
        const uint64_t nbytes = gdx_packed_bytes(s1);
        std::vector<uint64_t> exc;
        if (s0) memset(packed_out, 0, (s0 + 3) / 4);
        // the packer writes whole bytes: start on the byte that holds symbol s0 (codes before it belong to no query)
        const uint64_t a0 = s0 & ~3ull;
        pack2_parallel(idx->pack, q->bytes + a0, s1 - a0, packed_out + a0 / 4, exc);
        for (uint64_t i = (s1 + 3) / 4; i < nbytes; ++i) packed_out[i] = 0;
        uint64_t first = ~0ull;
        for (uint64_t pos : exc) {
            const uint64_t sym = a0 + pos;
            if (sym < s0) continue;
            first = q->offsets ? (uint64_t)(std::upper_bound(q->offsets, q->offsets + q->nq + 1, sym) - q->offsets) - 1
                               : (sym - q->first_symbol) / q->fixed_len;
            break;
        }
        if (first_unencodable) *first_unencodable = first;
        return GDX_OK;
    });
}

extern "C" gdx_status gdx_pack_symbols(const gdx_alphabet *alphabet, const uint8_t *io_bytes, uint64_t n,
                                       uint8_t *packed_out, uint64_t *exception_positions, uint64_t capacity,
                                       uint64_t *num_exceptions) {
    return guarded([&]() -> gdx_status {
        if (!alphabet || (n && (!io_bytes || !packed_out))) return fail(GDX_ERR_BAD_ARG, "NULL argument");
        GDX_TRY(validate_alphabet(*alphabet));
        PackTable t;
        build_pack_table(alphabet->io_to_dense, alphabet->num_searchable_dense_symbols, t);
        if (!t.usable) return fail(GDX_ERR_UNSUPPORTED, "2-bit packing needs an alphabet with at most 4 searchable symbols");
        std::vector<uint64_t> exc;
        pack2_parallel(t, io_bytes, n, packed_out, exc);
        if (num_exceptions) *num_exceptions = exc.size();
        for (uint64_t i = 0; i < exc.size() && i < capacity && exception_positions; ++i) exception_positions[i] = exc[i];
        return GDX_OK;
    });
}

// ================================================================================================
// one batch over several replicas
// ================================================================================================
extern "C" void gdx_shard_range(uint64_t n_items, uint32_t shard, uint32_t n_shards, uint64_t *begin, uint64_t *end) {
    // contiguous, balanced, order-preserving: the first n_items % n_shards shards get one item more
    if (n_shards == 0) n_shards = 1;
    if (shard > n_shards) shard = n_shards;
    const uint64_t base = n_items / n_shards, rem = n_items % n_shards;
    const uint64_t b = (uint64_t)shard * base + std::min<uint64_t>(shard, rem);
    if (begin) *begin = b;
    if (end) *end = shard >= n_shards ? b : b + base + (shard < rem ? 1 : 0);
}

namespace {

// the queries [b, e) of a batch as a batch of their own
gdx_queries sub_batch(const gdx_queries &q, uint64_t b, uint64_t e) {
    gdx_queries s = q;
    s.nq = e - b;
    if (q.offsets) {
        s.offsets = q.offsets + b;  // absolute stream positions: the symbol pointer stays
    } else {
        const uint64_t sym = (uint64_t)q.first_symbol + b * q.fixed_len;
        if (q.encoding == GDX_QUERIES_PACKED_2BIT) {
            s.bytes = q.bytes + sym / 4;
            s.first_symbol = (uint32_t)(sym & 3);
        } else {
            s.bytes = q.bytes + sym;
            s.first_symbol = 0;
        }
    }
    return s;
}

struct ShardResult {
    gdx_status st = GDX_OK;
    std::string error;
    uint64_t error_query = 0;
    gdx_stats stats = {};
};

// runs work(k, begin, end) for every local shard on its own host thread (shard 0 on the caller's) and merges
// status, error text and statistics into the caller's thread-local state
template <class F>
gdx_status run_sharded(gdx_index *const *replicas, uint32_t n_local, uint32_t first_shard, uint32_t n_shards,
                       const gdx_queries *queries, F &&work) {
    if (!replicas || n_local == 0 || n_shards == 0 || first_shard + n_local > n_shards)
        return fail(GDX_ERR_BAD_ARG, "bad shard arguments (n_local %u, first_shard %u, n_shards %u)", n_local, first_shard, n_shards);
    for (uint32_t k = 0; k < n_local; ++k) {
        if (!replicas[k]) return fail(GDX_ERR_BAD_ARG, "replica %u is NULL", k);
        const ImageHeader &h0 = replicas[0]->h, &hk = replicas[k]->h;
        if (hk.n != h0.n || hk.sigma != h0.sigma || hk.ns != h0.ns || hk.sampling_rate != h0.sampling_rate ||
            hk.lookup_depth != h0.lookup_depth || hk.ntexts != h0.ntexts)
            return fail(GDX_ERR_BAD_ARG, "replica %u is not a replica of replica 0", k);
    }
    GDX_TRY(check_queries(replicas[0], queries));
    std::vector<ShardResult> res(n_local);
    auto run = [&](uint32_t k) {
        const bool helps = HostPool::t_caller_helps;
        if (n_local > 1) HostPool::t_caller_helps = false;  // the pool alone does the staging work of all shards
        uint64_t b = 0, e = 0;
        gdx_shard_range(queries->nq, first_shard + k, n_shards, &b, &e);
        ShardResult &r = res[k];
        r.st = guarded([&] { return work(k, b, e); });
        if (r.st != GDX_OK) {
            r.error = t_error;
            r.error_query = b + t_error_query;
        }
        r.stats = t_stats;
        HostPool::t_caller_helps = helps;
    };
    std::vector<std::thread> threads;
    threads.reserve(n_local);
    uint32_t started = 1;
    try {
        for (; started < n_local; ++started) threads.emplace_back(run, started);
    } catch (const std::exception &) {  // could not start a thread: the shards without one run here, one after the other
    }
    run(0);
    for (uint32_t k = started; k < n_local; ++k) run(k);
    for (auto &t : threads) t.join();
    gdx_stats sum = {};
    gdx_status st = GDX_OK;
    for (uint32_t k = 0; k < n_local; ++k) {
        const gdx_stats &x = res[k].stats;
        sum.queries += x.queries;
        sum.lf_steps += x.lf_steps;
        sum.hits += x.hits;
        sum.walk_steps += x.walk_steps;
        sum.locate_walk_steps += x.locate_walk_steps;
        sum.kernel_launches += x.kernel_launches;
        sum.verified_queries += x.verified_queries;
        sum.packed_queries += x.packed_queries;
        sum.exception_queries += x.exception_queries;
        sum.h2d_bytes += x.h2d_bytes;
        sum.d2h_bytes += x.d2h_bytes;
        sum.kernel_ms_search = std::max(sum.kernel_ms_search, x.kernel_ms_search);  // the shards run side by side
        sum.kernel_ms_locate = std::max(sum.kernel_ms_locate, x.kernel_ms_locate);
        if (st == GDX_OK && res[k].st != GDX_OK) {  // the error of the lowest failing shard = the first offending query
            st = res[k].st;
            t_error = res[k].error;
            t_error_query = res[k].error_query;
        }
    }
    sum.shards = n_local;
    t_stats = sum;
    return st;
}

}  // namespace

extern "C" gdx_status gdx_count_many_sharded(gdx_index *const *replicas, uint32_t n_local, uint32_t first_shard,
                                             uint32_t n_shards, const gdx_queries *queries, uint64_t *counts) {
    return guarded([&] {
        return run_sharded(replicas, n_local, first_shard, n_shards, queries, [&](uint32_t k, uint64_t b, uint64_t e) {
            const gdx_queries sub = sub_batch(*queries, b, e);
            return search_many_impl(replicas[k], &sub, counts ? counts + b : nullptr, nullptr, 8, 1, "gdx_count_many_sharded");
        });
    });
}

extern "C" gdx_status gdx_cursors_many_sharded(gdx_index *const *replicas, uint32_t n_local, uint32_t first_shard,
                                               uint32_t n_shards, const gdx_queries *queries, uint64_t *starts,
                                               uint64_t *ends) {
    return guarded([&] {
        return run_sharded(replicas, n_local, first_shard, n_shards, queries, [&](uint32_t k, uint64_t b, uint64_t e) {
            const gdx_queries sub = sub_batch(*queries, b, e);
            return search_many_impl(replicas[k], &sub, starts ? starts + b : nullptr, ends ? ends + b : nullptr, 8, 0,
                                    "gdx_cursors_many_sharded");
        });
    });
}

extern "C" gdx_status gdx_locate_many_sharded_compact(gdx_index *const *replicas, uint32_t n_local, uint32_t first_shard,
                                                      uint32_t n_shards, const gdx_queries *queries, uint32_t *hit_counts,
                                                      gdx_hit32 **shard_hits, uint64_t *shard_num_hits) {
    return guarded([&]() -> gdx_status {
        if (!hit_counts || !shard_hits || !shard_num_hits) return fail(GDX_ERR_BAD_ARG, "output is NULL");
        for (uint32_t k = 0; k < n_local; ++k) {
            shard_hits[k] = nullptr;
            shard_num_hits[k] = 0;
        }
        gdx_status st = run_sharded(replicas, n_local, first_shard, n_shards, queries, [&](uint32_t k, uint64_t b, uint64_t e) {
            const gdx_queries sub = sub_batch(*queries, b, e);
            return locate_many_impl(replicas[k], &sub, nullptr, reinterpret_cast<gdx_hit **>(&shard_hits[k]), &shard_num_hits[k], true,
                                    hit_counts + b);
        });
        uint64_t total = 0;
        for (uint32_t k = 0; k < n_local; ++k) {
            if (st != GDX_OK && shard_hits[k]) {
                gdx_free_hits(replicas[k], reinterpret_cast<gdx_hit *>(shard_hits[k]));
                shard_hits[k] = nullptr;
                shard_num_hits[k] = 0;
            }
            total += shard_num_hits[k];
        }
        t_stats.hits = total;
        return st;
    });
}

extern "C" gdx_status gdx_locate_many_sharded(gdx_index *const *replicas, uint32_t n_local, uint32_t first_shard,
                                              uint32_t n_shards, const gdx_queries *queries, uint64_t *hit_offsets,
                                              gdx_hit **shard_hits, uint64_t *shard_first_hit) {
    return guarded([&]() -> gdx_status {
        if (!hit_offsets || !shard_hits || !shard_first_hit) return fail(GDX_ERR_BAD_ARG, "output is NULL");
        for (uint32_t k = 0; k < n_local; ++k) shard_hits[k] = nullptr;
        std::vector<uint64_t> n_hits(n_local, 0), begins(n_local, 0), ends(n_local, 0);
        // every shard writes its own CSR offsets (starting at 0) into its range of hit_offsets -- not the entry one
        // past the range, which is the next shard's first; the ranges are rebased below
        gdx_status st = run_sharded(replicas, n_local, first_shard, n_shards, queries, [&](uint32_t k, uint64_t b, uint64_t e) {
            const gdx_queries sub = sub_batch(*queries, b, e);
            begins[k] = b;
            ends[k] = e;
            return locate_many_impl(replicas[k], &sub, hit_offsets + b, &shard_hits[k], &n_hits[k], false);
        });
        if (st != GDX_OK) {
            for (uint32_t k = 0; k < n_local; ++k)
                if (shard_hits[k]) {
                    gdx_free_hits(replicas[k], shard_hits[k]);
                    shard_hits[k] = nullptr;
                }
            return st;
        }
        uint64_t acc = 0;
        for (uint32_t k = 0; k < n_local; ++k) {
            shard_first_hit[k] = acc;
            if (acc) {
                uint64_t *o = hit_offsets + begins[k];
                const uint64_t cnt = ends[k] - begins[k], base = acc;
                constexpr uint64_t kPiece = 1ull << 18;
                HostPool::get().parallel_for(div_up(cnt, kPiece), [&](uint64_t piece) {
                    const uint64_t lo = piece * kPiece, hi = std::min(cnt, lo + kPiece);
                    for (uint64_t i = lo; i < hi; ++i) o[i] += base;
                });
            }
            acc += n_hits[k];
        }
        shard_first_hit[n_local] = acc;
        hit_offsets[ends[n_local - 1]] = acc;
        t_stats.hits = acc;
        return GDX_OK;
    });
}

extern "C" gdx_status gdx_locate_intervals(const gdx_index *idx, const uint64_t *starts, const uint64_t *ends,
                                           uint64_t n, uint64_t *hit_offsets, gdx_hit **hits, uint64_t *num_hits) {
    return guarded([&]() -> gdx_status {
        GDX_TRY(begin_call(idx, "gdx_locate_intervals"));
        if (!hit_offsets || !hits || !num_hits || (n && (!starts || !ends))) return fail(GDX_ERR_BAD_ARG, "NULL argument");
        *hits = nullptr;
        *num_hits = 0;
        for (uint64_t i = 0; i < n; ++i)
            if (starts[i] > ends[i] || ends[i] > idx->h.n)
                return fail(GDX_ERR_BAD_ARG, "interval %llu is not inside [0, text_len]", (unsigned long long)i);
        std::shared_lock<std::shared_mutex> cfg(idx->cfg_mu);
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
        gdx_status rc = locate_device_intervals(idx, ws, ws->starts.as<uint64_t>(), ws->ends.as<uint64_t>(), n, &total);
        if (rc == GDX_OK) rc = finish_locate(idx, ws, n, total, hit_offsets, hits, num_hits);
        if (rc != GDX_OK) drain(ws);
        return rc;
    });
}

extern "C" void gdx_free_hits(const gdx_index *idx, gdx_hit *hits) {
    if (!idx || !hits) return;
    std::lock_guard<std::mutex> lk(idx->mu);
    for (auto &p : idx->pinned)
        if (p.p == hits) p.in_use = false;
}

static gdx_status extend_many_on(const gdx_index *idx, Workspace *ws, uint64_t *starts, uint64_t *ends,
                                 const uint8_t *io_symbols, uint64_t n);

extern "C" gdx_status gdx_extend_many(const gdx_index *idx, uint64_t *starts, uint64_t *ends,
                                      const uint8_t *io_symbols, uint64_t n) {
    return guarded([&]() -> gdx_status {
        GDX_TRY(begin_call(idx, "gdx_extend_many"));
        if (n == 0) return GDX_OK;
        if (!starts || !ends || !io_symbols) return fail(GDX_ERR_BAD_ARG, "NULL argument");
        std::shared_lock<std::shared_mutex> cfg(idx->cfg_mu);
        DeviceGuard guard(idx->device);
        WsLease lease(idx);
        Workspace *ws = lease.w;
        if (!ws) return fail(GDX_ERR_CUDA, "could not create a CUDA workspace: %s", cudaGetErrorString(cudaGetLastError()));
        const gdx_status rc = extend_many_on(idx, ws, starts, ends, io_symbols, n);
        if (rc != GDX_OK) drain(ws);
        return rc;
    });
}

static gdx_status extend_many_on(const gdx_index *idx, Workspace *ws, uint64_t *starts, uint64_t *ends,
                                 const uint8_t *io_symbols, uint64_t n) {
    {
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
}

extern "C" gdx_status gdx_cursor_for_query(const gdx_index *idx, const uint8_t *query, uint64_t len, uint64_t *start,
                                           uint64_t *end) {
    return guarded([&]() -> gdx_status {
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
    });
}

// ================================================================================================
// device-resident entry points
// ================================================================================================
static gdx_status search_device(const gdx_index *idx, const gdx_queries *dq_in, uint64_t *a, uint64_t *b, int mode,
                                uint64_t *d_error, void *stream) {
    if (!idx || !dq_in) return fail(GDX_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard(idx->device);
    // (offsets is a device pointer here: nothing of the batch may be dereferenced on the host)
    if (dq_in->encoding > GDX_QUERIES_PACKED_2BIT) return fail(GDX_ERR_BAD_ARG, "unknown queries->encoding %u", dq_in->encoding);
    if (dq_in->encoding == GDX_QUERIES_PACKED_2BIT && idx->h.ns > 4)
        return fail(GDX_ERR_UNSUPPORTED, "2-bit packed queries need an alphabet with at most 4 searchable symbols");
    DevQueries dq;
    const bool packed = dq_in->encoding == GDX_QUERIES_PACKED_2BIT;
    if (packed && (reinterpret_cast<uintptr_t>(dq_in->bytes) & 3u))
        return fail(GDX_ERR_BAD_ARG, "a packed device batch must be 4-byte aligned (and a multiple of 4 bytes long)");
    dq.bytes = packed ? nullptr : dq_in->bytes;
    dq.packed = packed ? reinterpret_cast<const uint32_t *>(dq_in->bytes) : nullptr;
    dq.offsets = dq_in->offsets;
    dq.offsets32 = nullptr;
    dq.fixed_len = dq_in->fixed_len;
    dq.nq = dq_in->nq;
    dq.base = 0;
    dq.shift = dq_in->first_symbol;
    dq.limit = ~0ull;  // device-resident batch: its extent is the caller's contract (header)
    cudaStream_t st = (cudaStream_t)stream;
    const SortPlan sp = plan_sort(idx, dq.nq, dq.offsets ? 0 : dq.fixed_len);
    const uint32_t *perm = nullptr;
    void *scratch = nullptr;
    if (sp.use) {  // stream-ordered scratch from the device's memory pool
        CUDA_TRY(cudaMallocAsync(&scratch, sp.total_bytes, st));
        GDX_TRY(sort_queries(idx, dq, sp, scratch, st, &perm));
    }
    GDX_TRY(dispatch_layout(idx->h.layout, [&](auto L) -> gdx_status {
        launch_search<decltype(L)>(idx, dq, a, b, mode, false, 0, d_error, nullptr, perm, nullptr, st);
        return GDX_OK;
    }));
    CUDA_TRY(cudaGetLastError());
    if (scratch) CUDA_TRY(cudaFreeAsync(scratch, st));
    return GDX_OK;
}

extern "C" gdx_status gdx_cursors_many_device(const gdx_index *idx, const gdx_queries *d_queries, uint64_t *d_starts,
                                              uint64_t *d_ends, uint64_t *d_error, void *stream) {
    return guarded([&] { return search_device(idx, d_queries, d_starts, d_ends, 0, d_error, stream); });
}
extern "C" gdx_status gdx_count_many_device(const gdx_index *idx, const gdx_queries *d_queries, uint64_t *d_counts,
                                            uint64_t *d_error, void *stream) {
    return guarded([&] { return search_device(idx, d_queries, d_counts, nullptr, 1, d_error, stream); });
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
        launch_locate_walk<decltype(L)>(idx, rows, num_hits, reinterpret_cast<ulonglong2 *>(d_hits), nullptr, st);
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
