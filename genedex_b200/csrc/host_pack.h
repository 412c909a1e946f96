// host_pack.h -- host-side staging helpers of the C ABI: a small thread pool and the 2-bit query packer.
//
// End to end the search path is bound by the PCIe transfer of the query bytes (375 MB per 7.5 M length-50
// queries, 7.2 ms at 52 GB/s, against 1 ms of kernels).  For alphabets with at most four searchable symbols
// (every DNA alphabet of the reference, alphabet.rs:251-300) the queries are therefore packed to 2 bits per
// symbol on the host while they are staged into pinned memory -- the stage that pageable caller memory
// needs anyway -- and the search kernel reads the packed words directly.  Bytes the packer cannot encode
// (invalid bytes, valid but unsearchable symbols such as `N`) are reported as exceptions; the queries that
// hold them are re-run through the IO-byte kernel so that the reference's lazy invalid-symbol behaviour
// (batch_computed_cursors.rs:84-87,106-113) stays exact.
#ifndef GDX_HOST_PACK_H
#define GDX_HOST_PACK_H

#include <stdint.h>

#include <functional>
#include <vector>

namespace gdx {

// A pool of host threads for the staging work between caller memory and pinned buffers.  Any number of
// jobs may be submitted concurrently (one per GPU worker thread of a sharded call); every job owns its
// counters, so a worker that is still leaving a finished job can never touch the next one.
class HostPool {
public:
    static HostPool &get();
    unsigned threads() const;  // workers + the calling thread
    unsigned resize(unsigned total);  // see host_pack.cpp; returns threads()
    // fn(piece) for every piece in [0, pieces); the caller works too; returns when all pieces are done
    void parallel_for(uint64_t pieces, const std::function<void(uint64_t)> &fn);
    void copy(void *dst, const void *src, uint64_t bytes);
    void widen_u32(uint64_t *dst, const uint32_t *src, uint64_t n);
    // false on the per-GPU threads of a sharded call: several callers at once would otherwise put more runnable
    // threads on the machine than it has CPUs (each of them only waits for its job then)
    static thread_local bool t_caller_helps;

private:
    HostPool();
    ~HostPool();
    struct Impl;
    Impl *impl_;
};

// io byte -> 2-bit code (dense - 1) for the searchable symbols, 0xff for everything else
struct PackTable {
    uint8_t code[256];
    // nibble-split form for the SIMD packer: a byte x is encodable iff lo_class[x & 15] & hi_class[x >> 4],
    // and its code is code_lo[x & 15] (checked exhaustively when the table is built)
    uint8_t lo_class[16], hi_class[16], code_lo[16];
    bool simd_ok;
    bool usable;  // the alphabet has at most four searchable symbols
};
// run-time form of GDX_PACK_PREFETCH / GDX_PACK_STREAM (A/B measurements inside one process)
void set_pack_tuning(int prefetch_bytes, int stream);
void build_pack_table(const uint8_t io_to_dense[256], uint32_t num_searchable, PackTable &out);

// Packs src[0, n) into dst (4 symbols per byte, symbol i at bits [2(i%4), 2(i%4)+2) of byte i/4, the last
// byte zero-filled).  Positions of bytes without a code are appended to `exceptions` (they are packed as 0).
void pack2_serial(const PackTable &t, const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t pos0,
                  std::vector<uint64_t> &exceptions);
// The same on the pool; `exceptions` comes back sorted.  dst must hold (n + 3) / 4 bytes.
void pack2_parallel(const PackTable &t, const uint8_t *src, uint64_t n, uint8_t *dst,
                    std::vector<uint64_t> &exceptions);

}  // namespace gdx
#endif
