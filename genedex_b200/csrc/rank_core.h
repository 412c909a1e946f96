// rank_core.h -- device rank-record layouts and their arithmetic.
//
// genedex keeps the occurrence table in three arrays (bit planes, u16 block offsets, superblock
// offsets; src/text_with_rank_support/condensed.rs:24-47) so one rank touches three cache lines.
// On the GPU one rank must be ONE aligned record, so planes and in-superblock offsets of a block
// are co-located, and count[c] (src/lib.rs:273-275) is folded into the superblock table:
//
//   LF(c, i) = sbc[(i >> 16) * noff + (c-1)]         (u64, = count[c] + #c in [0, superblock start))
//            + off_c(record(i))                      (u16, #c in [superblock start, block start))
//            + popcount(match_c(planes) & below(i))  (condensed.rs:306-340)
//
// Layout K32 (sigma <= 6, e.g. every DNA alphabet): 64 positions / 32 B = one DRAM sector
//     bytes  0..23  three u64 bit planes (plane p holds bit p of each symbol, block.rs:142-192)
//     bytes 24..31  four u16 offsets for dense symbols 1..4
//   With sigma == 6 (ascii_dna_with_n) symbol 5 has no slot; its rank is derived exactly as
//   i - #sentinels before i - sum of the four stored ranks (rare path: `N` is not searchable and
//   only the locate walk steps over it).
// Layout KG<B> (any sigma, B = ceil(log2 sigma) planes): 128 positions / `stride` bytes
//     bytes [16p, 16p+16)           plane p: low u64 = positions 0..63, high u64 = 64..127
//     bytes 16B + 2(c-1), 2 bytes   u16 offset of dense symbol c, c = 1..sigma-1
//   stride = 16B + 2(sigma-1) rounded up to 32 (protein sigma=21: 120 -> 128 B = one cache line).
// The rank of the sentinel (dense 0) is never needed by the search path: queries cannot contain
// it (alphabet.rs:195-198) and the locate walk returns before ranking it
// (sampled_suffix_array.rs:121-126).
//
// Everything here is __host__ __device__ so that tests/host_emul can run the exact arithmetic on
// the CPU against the oracle (test infrastructure; the product only ever runs it on the GPU).
#ifndef GDX_RANK_CORE_H
#define GDX_RANK_CORE_H

#include <stdint.h>

#if defined(__CUDACC__)
#define GDX_HD __host__ __device__ __forceinline__
#else
#define GDX_HD inline
#endif

namespace gdx {

constexpr uint32_t kSuperblockLog2 = 16;  // condensed.rs:34: 65536 positions per superblock
constexpr uint32_t kLayoutK32 = 0;
constexpr uint32_t kLayoutKG = 1;

struct RankLayout {
    uint32_t kind;            // kLayoutK32 / kLayoutKG
    uint32_t planes;          // ceil(log2 sigma), condensed.rs:417-419
    uint32_t noff;            // dense symbols 1..noff have an in-record offset
    uint32_t stride;          // record bytes
    uint32_t log2_pos;        // log2(positions per record): 6 or 7
    uint32_t derived_symbol;  // 0 = none, else the symbol whose rank is derived (K32, sigma == 6)
};

GDX_HD uint32_t ilog2_ceil(uint32_t v) {  // condensed.rs:417-419
    uint32_t b = 0;
    while ((1u << b) < v) ++b;
    return b;
}

GDX_HD RankLayout choose_layout(uint32_t sigma) {
    RankLayout L;
    L.planes = ilog2_ceil(sigma);
    if (L.planes == 0) L.planes = 1;
    if (sigma <= 6) {
        L.kind = kLayoutK32;
        L.noff = sigma - 1 < 4 ? sigma - 1 : 4;
        L.stride = 32;
        L.log2_pos = 6;
        L.derived_symbol = sigma == 6 ? 5 : 0;
    } else {
        L.kind = kLayoutKG;
        L.noff = sigma - 1;
        L.stride = (16 * L.planes + 2 * (sigma - 1) + 31) / 32 * 32;
        L.log2_pos = 7;
        L.derived_symbol = 0;
    }
    return L;
}

GDX_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// bits [0, bit) set; bit in [0, 64]
GDX_HD uint64_t below(uint32_t bit) { return bit >= 64 ? ~0ull : ((1ull << bit) - 1ull); }

// condensed.rs:306-336: AND over planes of (bit p of c ? plane : ~plane)
template <int B>
GDX_HD uint64_t match_planes(const uint64_t *planes, uint32_t c) {
    uint64_t acc = ~0ull;
#pragma unroll
    for (int p = 0; p < B; ++p) {
        uint64_t w = planes[p];
        acc &= ((c >> p) & 1u) ? w : ~w;
    }
    return acc;
}

// ---- K32 ------------------------------------------------------------------------------------
// rec[0..2] planes, rec[3] = four packed u16 offsets
GDX_HD uint32_t k32_offset(const uint64_t rec[4], uint32_t c /*1..4*/) {
    return (uint32_t)(rec[3] >> (16u * (c - 1u))) & 0xffffu;
}
GDX_HD uint32_t k32_local_rank(const uint64_t rec[4], uint32_t c /*1..4*/, uint32_t bit /*0..63*/) {
    return k32_offset(rec, c) + (uint32_t)popc64(match_planes<3>(rec, c) & below(bit));
}
// sum of the local ranks of symbols 1..4 (for the derived symbol)
GDX_HD uint32_t k32_local_rank_sum(const uint64_t rec[4], uint32_t bit) {
    uint32_t s = 0;
#pragma unroll
    for (uint32_t c = 1; c <= 4; ++c) s += k32_local_rank(rec, c, bit);
    return s;
}
GDX_HD uint32_t k32_symbol_at(const uint64_t rec[4], uint32_t bit) {  // condensed.rs:343-362
    return (uint32_t)(((rec[0] >> bit) & 1u) | (((rec[1] >> bit) & 1u) << 1) |
                      (((rec[2] >> bit) & 1u) << 2));
}

// ---- KG<B> ----------------------------------------------------------------------------------
// lo[p], hi[p]: the two words of plane p; bit in 0..127
template <int B>
GDX_HD uint32_t kg_block_count(const uint64_t *lo, const uint64_t *hi, uint32_t c, uint32_t bit) {
    uint64_t mlo = match_planes<B>(lo, c), mhi = match_planes<B>(hi, c);
    uint64_t masklo = below(bit);                       // bit >= 64 -> all ones
    uint64_t maskhi = bit > 64 ? below(bit - 64) : 0ull;
    return (uint32_t)(popc64(mlo & masklo) + popc64(mhi & maskhi));
}
template <int B>
GDX_HD uint32_t kg_symbol_at(const uint64_t *lo, const uint64_t *hi, uint32_t bit) {
    const bool low = bit < 64;
    uint32_t sh = bit & 63u, s = 0;
#pragma unroll
    for (int p = 0; p < B; ++p) s |= (uint32_t)(((low ? lo[p] : hi[p]) >> sh) & 1u) << p;
    return s;
}

// ---- packing (one record from its symbols) -----------------------------------------------------
// symbols: up to 64 dense symbols of one K32 block (nsym <= 64), planes out
GDX_HD void pack_planes64(const uint8_t *symbols, uint32_t nsym, uint32_t nplanes, uint64_t *planes) {
    for (uint32_t p = 0; p < nplanes; ++p) planes[p] = 0;
    for (uint32_t j = 0; j < nsym; ++j) {
        uint32_t s = symbols[j];
        for (uint32_t p = 0; p < nplanes; ++p) planes[p] |= (uint64_t)((s >> p) & 1u) << j;
    }
}

// text id of a concatenated-text position = lower_bound over the sorted sentinel positions,
// clamped to the last text (text_id_search_tree.rs:35-64)
GDX_HD uint64_t lower_bound_u64(const uint64_t *a, uint64_t n, uint64_t key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace gdx
#endif
