// host_build.h -- host side of index construction (construction is not on the search path).
//
// Restates what genedex does on the host before the search structures exist
// (src/construction/mod.rs:25-57): concatenate + densely encode the texts, count table, suffix
// array, BWT + text-border lookup, SA sampling.  The reference calls libsais (C, un-vendored);
// here the suffix array comes from an SA-IS written for this repo (suffix arrays are unique, so
// any correct SACA reproduces the reference's array).
#ifndef GDX_HOST_BUILD_H
#define GDX_HOST_BUILD_H

#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/genedex_b200.h"

namespace gdx {

struct ConcatText {
    std::vector<uint8_t> text;         // dense symbols, one 0 sentinel after each text
    std::vector<uint64_t> sentinels;   // position of each sentinel
    std::vector<uint64_t> count;       // sigma + 1 exclusive prefix sums (lib.rs:95)
};

struct HostParts {
    uint64_t n = 0;
    std::vector<uint8_t> bwt;
    std::vector<uint64_t> samples;       // SA[0], SA[s], ...
    std::vector<uint64_t> isa_samples;   // ISA[0], ISA[s], ... (row of every s-th text position)
    std::vector<uint64_t> border_rows;   // ascending
    std::vector<uint64_t> border_pos;
};

uint64_t storage_max(uint32_t storage);

// construction/mod.rs:255-336; returns GDX_ERR_INVALID_SYMBOL (with *bad_text) like the panic
gdx_status concat_texts(const uint8_t *texts, const uint64_t *text_offsets, uint64_t num_texts,
                        const gdx_alphabet &alphabet, ConcatText &out, uint64_t *bad_text);

// suffix array of the dense text (libsais convention: end of text < every symbol, the 0 sentinels
// are ordinary symbols), 64-bit entries
void suffix_array_sais(const uint8_t *text, uint64_t n, uint32_t sigma, std::vector<int64_t> &sa);

// bwt.rs:93-116 + sampled_suffix_array.rs:27-54
void parts_from_suffix_array(const uint8_t *text, uint64_t n, const int64_t *sa,
                             uint32_t sampling_rate, HostParts &out);

// suffix array + parts in one go; uses 32-bit suffix array entries when the text is shorter than 2^31
// (half the memory of the 64-bit path, like the reference's i32 index storage)
void host_parts_from_text(const uint8_t *text, uint64_t n, uint32_t sigma, uint32_t sampling_rate, HostParts &out);

}  // namespace gdx
#endif
