/*
 * gdx_oracle.h -- CPU restatement ("oracle") of genedex's FM-index search path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker / CPU baseline.  The product (genedex_b200/) never links or calls it.
 *
 * Parity status: PINNED by the reference's own known-answer tests (tests/test_oracle_golden.py
 * replays tests/fmindex.rs:20-154, tests/text_with_rank_support.rs:77-119,
 * src/text_id_search_tree.rs:160-172, src/construction/mod.rs:374-398,
 * src/sampled_suffix_array.rs:168-179, examples/basic_usage.rs:16, examples/cursor.rs:18,24 and the
 * proptest-regression edge inputs) and by hypothesis replays of the reference's property tests
 * against naive search / naive rank.  The reference itself (Rust) cannot be compiled in this
 * environment (no cargo/rustc, un-vendored libsais); see DESIGN.md.
 *
 * Every function cites the reference file:line (relative to /root/reference) that it restates.
 */
#ifndef GDX_ORACLE_H
#define GDX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Index storage integer `I` of the reference (src/construction/mod.rs:59-73, 146-252). */
enum { GDXO_I32 = 0, GDXO_U32 = 1, GDXO_I64 = 2 };

/* Return codes; the non-zero ones model the reference's panics. */
enum {
    GDXO_OK = 0,
    GDXO_PANIC_INVALID_SYMBOL = 1, /* alphabet.rs:195-198 "symbol in io representation should be valid" */
    GDXO_PANIC_LOOKUP_OOB = 2,     /* lookup_table.rs:215-216 slice index out of bounds            */
    GDXO_PANIC_TEXT_TOO_LONG = 3,  /* construction/mod.rs:34                                          */
    GDXO_PANIC_BAD_CONFIG = 4,     /* config.rs:28 sampling rate 0; alphabet.rs:166-189               */
    GDXO_ERR_ALLOC = 5
};

typedef struct gdxo_index gdxo_index;

typedef struct {
    uint64_t text_id;
    uint64_t position;
} gdxo_hit; /* lib.rs:331-335 */

/* ---- construction (construction/mod.rs:25-57, lib.rs:118-142) --------------------------------- */

/* texts: concatenated IO bytes of all texts; text_offsets has ntexts+1 entries. */
int gdxo_build(const uint8_t *texts, const uint64_t *text_offsets, uint64_t ntexts,
               const uint8_t io_to_dense[256], uint32_t sigma, uint32_t num_searchable,
               uint32_t sampling_rate, uint32_t lookup_depth, int storage, gdxo_index **out);

/* Build from already computed parts (used at sizes where the oracle's simple SACA is too slow):
 * bwt = dense BWT (n bytes); sampled_sa = SA[0], SA[s], SA[2s], ... (ceil(n/s) entries, 64-bit);
 * border_rows/border_pos = the text_border_lookup pairs (bwt.rs:108-116), any order; may be NULL
 * together with sampled_sa when only count is needed. */
int gdxo_from_parts(const uint8_t *bwt, uint64_t n, const uint8_t io_to_dense[256], uint32_t sigma,
                    uint32_t num_searchable, const uint64_t *count /* sigma+1 */,
                    const uint64_t *sentinel_indices, uint64_t ntexts, const uint64_t *sampled_sa,
                    uint64_t n_samples, uint32_t sampling_rate, const uint64_t *border_rows,
                    const uint64_t *border_pos, uint64_t n_border, uint32_t lookup_depth,
                    int storage, int nthreads, gdxo_index **out);

void gdxo_free(gdxo_index *idx);

/* ---- introspection (for pinning the intermediate structures) ---------------------------------- */
uint64_t gdxo_text_len(const gdxo_index *idx);  /* lib.rs:292-294 (includes sentinels) */
uint64_t gdxo_num_texts(const gdxo_index *idx); /* lib.rs:287-289 */
const uint8_t *gdxo_dense_text(const gdxo_index *idx); /* NULL when built from parts */
const int64_t *gdxo_suffix_array(const gdxo_index *idx); /* full SA, NULL when built from parts */
const uint8_t *gdxo_bwt(const gdxo_index *idx);
const uint64_t *gdxo_count_array(const gdxo_index *idx); /* sigma+1 entries, lib.rs:95 */
const uint64_t *gdxo_sentinel_indices(const gdxo_index *idx);
const uint64_t *gdxo_frequency_table(const gdxo_index *idx); /* 256 entries */
uint64_t gdxo_num_border(const gdxo_index *idx);
const uint64_t *gdxo_border_rows(const gdxo_index *idx); /* sorted ascending */
const uint64_t *gdxo_border_pos(const gdxo_index *idx);
uint64_t gdxo_num_samples(const gdxo_index *idx);
uint64_t gdxo_sample(const gdxo_index *idx, uint64_t k);
/* condensed.rs:24-30: the three arrays in the reference's own layout */
const uint64_t *gdxo_blocks(const gdxo_index *idx, uint64_t *len);
const uint16_t *gdxo_block_offsets(const gdxo_index *idx, uint64_t *len);
uint64_t gdxo_superblock_offset(const gdxo_index *idx, uint64_t k);
uint64_t gdxo_num_superblock_offsets(const gdxo_index *idx);
uint64_t gdxo_lookup_table_len(const gdxo_index *idx, uint32_t depth);
void gdxo_lookup_entry(const gdxo_index *idx, uint32_t depth, uint64_t i, uint64_t *start,
                       uint64_t *end);

/* ---- rank structure alone (text_with_rank_support/mod.rs:88-133) ------------------------------ */
typedef struct gdxo_rank gdxo_rank;
gdxo_rank *gdxo_rank_construct(const uint8_t *dense_text, uint64_t n, uint32_t sigma, int storage);
void gdxo_rank_free(gdxo_rank *r);
uint64_t gdxo_rank_query(const gdxo_rank *r, uint8_t symbol, uint64_t idx); /* condensed.rs:291-341 */
uint8_t gdxo_rank_symbol_at(const gdxo_rank *r, uint64_t idx);              /* condensed.rs:343-362 */
/* batched twin, condensed.rs:137-287; borders may be start>end, nq <= 64 */
void gdxo_rank_batch(const gdxo_rank *r, const uint8_t *symbols, uint64_t *starts, uint64_t *ends,
                     uint32_t nq);

/* ---- the other rank variants behind the same trait (lib.rs:104-113): Condensed / Flat x Block64 / Block512,
 * in the reference's own array layouts (condensed.rs:24-30, flat.rs:24-30, block.rs:66-192) --------------- */
enum { GDXO_RANK_CONDENSED = 0, GDXO_RANK_FLAT = 1 };
typedef struct gdxo_vrank gdxo_vrank;
gdxo_vrank *gdxo_vrank_construct(const uint8_t *dense_text, uint64_t n, uint32_t sigma, int storage,
                                 int variant, uint32_t block_bits /* 64 or 512 */);
void gdxo_vrank_free(gdxo_vrank *r);
uint64_t gdxo_vrank_query(const gdxo_vrank *r, uint8_t symbol, uint64_t idx); /* condensed.rs:291-341, flat.rs:221-246 */
uint8_t gdxo_vrank_symbol_at(const gdxo_vrank *r, uint64_t idx);              /* condensed.rs:343-362, flat.rs:248-267 */
const uint64_t *gdxo_vrank_blocks(const gdxo_vrank *r, uint64_t *len_words);  /* interleaved_blocks as u64 words */
const uint16_t *gdxo_vrank_block_offsets(const gdxo_vrank *r, uint64_t *len); /* condensed only */
uint64_t gdxo_vrank_num_superblock_offsets(const gdxo_vrank *r);
uint64_t gdxo_vrank_superblock_offset(const gdxo_vrank *r, uint64_t k);
uint64_t gdxo_vrank_superblock_size(const gdxo_vrank *r);

/* ---- text id tree alone (text_id_search_tree.rs) ---------------------------------------------- */
uint64_t gdxo_tree_lookup(const uint64_t *sentinel_indices, uint64_t ntexts, uint64_t pos);

/* ---- search ----------------------------------------------------------------------------------- */

/* lib.rs:217-235 single query cursor.  rc may be GDXO_PANIC_*. */
int gdxo_cursor_for_query(const gdxo_index *idx, const uint8_t *q, uint64_t m, uint64_t *start,
                          uint64_t *end);
/* cursor.rs:34-51 */
int gdxo_extend_query_front(const gdxo_index *idx, uint8_t io_symbol, uint64_t *start,
                            uint64_t *end);

/* batch_computed_cursors.rs:36-199 with BATCH_SIZE = 64 (lib.rs:115); queries i =
 * qbytes[qoffsets[i] .. qoffsets[i+1]).  On a panic the index of the batch's first query is
 * written to *panic_query (the reference unwinds out of the iterator at that batch). */
int gdxo_cursors_many(const gdxo_index *idx, const uint8_t *qbytes, const uint64_t *qoffsets,
                      uint64_t nq, uint64_t *starts, uint64_t *ends, uint64_t *panic_query);

/* lib.rs:155-161, contiguous per-thread chunks (README.md:18: the crate's search is single
 * threaded; users split the query list).  nthreads <= 0 means all online cores. */
int gdxo_count_many(const gdxo_index *idx, const uint8_t *qbytes, const uint64_t *qoffsets,
                    uint64_t nq, int nthreads, uint64_t *counts, uint64_t *panic_query);

/* lib.rs:187-197: hits of rows [start,end) in SA-row order; out must hold end-start hits */
int gdxo_locate_interval(const gdxo_index *idx, uint64_t start, uint64_t end, gdxo_hit *out);

/* lib.rs:179-185: CSR output; hit_offsets has nq+1 entries; *hits is malloc'ed (gdxo_free_hits) */
int gdxo_locate_many(const gdxo_index *idx, const uint8_t *qbytes, const uint64_t *qoffsets,
                     uint64_t nq, int nthreads, uint64_t *hit_offsets, gdxo_hit **hits,
                     uint64_t *panic_query);
void gdxo_free_hits(gdxo_hit *hits);

int gdxo_online_cores(void);

/* Independent check of an index made with gdxo_from_parts (its BWT, suffix-array samples and border map come
 * from somewhere else, e.g. from the product's device construction): are they the ones of `dense_text`
 * (construction/mod.rs:255-308 output, n symbols incl. sentinels)?  The text is cut at every sentinel and every
 * ~n/8192 symbols; the row of the suffix at each cut is found with the index' own backward search (sentinel
 * rows: from the definition of the suffix order), then LF-walked down to the cut below.  Every row's BWT
 * symbol must be the text symbol in front of the walked position (bwt.rs:93-105), every sampled row must hold
 * that position (sampled_suffix_array.rs:27-54), every border row too (bwt.rs:108-116), every walk must land
 * on the searched row of the cut below, and all n rows must be visited exactly once -- which pins the BWT and
 * the samples to the text without trusting whoever built them.  O(n) LF steps, parallel over the cuts. */
int gdxo_verify_against_text(const gdxo_index *idx, const uint8_t *dense_text, uint64_t n, int nthreads,
                             uint64_t *violations, uint64_t *rows_visited);

#ifdef __cplusplus
}
#endif
#endif
