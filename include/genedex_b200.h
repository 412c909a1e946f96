/*
 * genedex_b200.h -- C ABI of the B200-native batched FM-index search engine.
 *
 * This is the drop-in boundary for genedex's batched search path.  The reference crate has no FFI
 * of its own; the seam is the crate-private call
 *     FmIndex::cursors_for_many_queries -> BatchComputedCursors   (src/lib.rs:241-246)
 * plus FmIndex::locate_interval (src/lib.rs:187-197).  Every entry point below names the reference
 * item it replaces (file:line relative to the genedex source tree).  INTEGRATION.md shows the Rust
 * `extern "C"` block + safe wrapper a maintainer would add to keep the crate's public API.
 *
 * Conventions
 *   - plain pointers and sizes only; all functions return a gdx_status (0 = OK) and never unwind;
 *   - caller owns every input buffer for the duration of the call only;
 *   - outputs whose size is known up front are caller-allocated; variable-size hit arrays are
 *     library-allocated pinned host memory released with gdx_free_hits();
 *   - re-entrant: any number of host threads may call into one gdx_index concurrently
 *     (the reference's search methods take &self and FmIndex is Send + Sync);
 *   - the library requires a CUDA device (sm_100a).  There is no CPU fallback: without a device
 *     every compute entry point returns GDX_ERR_CUDA.
 */
#ifndef GENEDEX_B200_H
#define GENEDEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDX_ABI_VERSION 1

typedef enum gdx_status {
    GDX_OK = 0,
    /* the reference panics "symbol in io representation should be valid" (src/alphabet.rs:195-198);
     * gdx_last_error_query() gives the index of the first offending query of the call */
    GDX_ERR_INVALID_SYMBOL = 1,
    GDX_ERR_BAD_ARG = 2,
    GDX_ERR_CUDA = 3,
    GDX_ERR_OOM = 4,
    GDX_ERR_TEXT_TOO_LONG = 5, /* src/construction/mod.rs:34 assert */
    GDX_ERR_UNSUPPORTED = 6
} gdx_status;

/* IndexStorage `I` of the reference (src/construction/mod.rs:59-73,146-252). It bounds the total
 * text length exactly like the reference; query results do not depend on it. */
typedef enum gdx_storage { GDX_I32 = 0, GDX_U32 = 1, GDX_I64 = 2 } gdx_storage;

/* where the suffix array is built (construction is not part of the search path) */
typedef enum gdx_construction {
    GDX_CONSTRUCT_HOST = 0,  /* SA-IS on the host (the reference: libsais on the host)          */
    GDX_CONSTRUCT_DEVICE = 1,/* prefix-doubling radix sort on the GPU (n < 2^32 - 1)            */
    GDX_CONSTRUCT_AUTO = 2   /* device when the text and the sort scratch (~36 B/symbol) fit    */
} gdx_construction;

/* src/alphabet.rs:24-28: 256-entry IO->dense table (0 = not in the alphabet; dense 0 is the
 * sentinel), number of dense symbols incl. the sentinel, and how many of them may be searched. */
typedef struct gdx_alphabet {
    uint8_t io_to_dense[256];
    uint32_t num_dense_symbols;
    uint32_t num_searchable_dense_symbols;
} gdx_alphabet;

/* src/config.rs:9-15,72-82 (defaults: sampling rate 4, lookup depth 0, Balanced) */
typedef struct gdx_config {
    uint32_t storage;                /* gdx_storage */
    uint32_t suffix_array_sampling_rate;
    uint32_t lookup_table_depth;
    uint32_t performance_priority;   /* accepted for API parity; never changes a query result */
    uint32_t construction;           /* gdx_construction */
    int32_t device;                  /* CUDA device ordinal, -1 = current device */
    uint32_t flags;                  /* GDX_FLAG_* */
} gdx_config;

/* re-check the device-built suffix array in O(n) (permutation + order) before it is used */
#define GDX_FLAG_VERIFY_SUFFIX_ARRAY 1u
/* do not keep the packed text in the device image (saves n/2 bytes for sigma <= 16, n otherwise);
 * count/locate then run every LF step instead of finishing one-row intervals by a text comparison.
 * Results are identical either way. */
#define GDX_FLAG_NO_TEXT 2u
/* do not keep the sampled inverse suffix array (n / rate entries, only kept together with the text):
 * cursors_for_many_queries then runs every LF step. Results are identical either way. */
#define GDX_FLAG_NO_INVERSE_SAMPLES 4u
/* never build the dense suffix array accelerator (see gdx_index_set_dense_suffix_array) for this index */
#define GDX_FLAG_NO_DENSE_SUFFIX_ARRAY 8u
/* never build the seed table accelerator (see gdx_index_set_seed_table_depth) for this index */
#define GDX_FLAG_NO_SEED_TABLE 16u

/* src/lib.rs:331-335 Hit { text_id, position } */
typedef struct gdx_hit {
    uint64_t text_id;
    uint64_t position;
} gdx_hit;

/* A batch of queries in IO representation: query i = bytes[offsets[i] .. offsets[i+1]).
 * offsets == NULL means every query has length fixed_len and query i starts at i * fixed_len.
 * Replaces `impl IntoIterator<Item = Q: AsRef<[u8]>>` of src/lib.rs:155-158,179-182,241-244. */
typedef struct gdx_queries {
    const uint8_t *bytes;
    const uint64_t *offsets; /* nq + 1 entries, or NULL */
    uint64_t fixed_len;
    uint64_t nq;
} gdx_queries;

/* The host-side structures of a constructed reference index, in the reference's own layout
 * (src/lib.rs:93-100): used to move an index built (or loaded from a savefile) by the Rust crate
 * onto the GPU.  Integers of type `I` are passed widened to 64 bit. */
typedef struct gdx_parts {
    gdx_alphabet alphabet;
    uint32_t storage;
    uint64_t text_len;                    /* incl. one sentinel per text, src/lib.rs:292-294       */
    const uint64_t *count;                /* num_dense_symbols + 1 entries, src/lib.rs:95          */
    /* CondensedTextWithRankSupport<I, Block64>, src/text_with_rank_support/condensed.rs:24-30     */
    const uint64_t *interleaved_blocks;   /* ceil((n+1)/64) * ceil(log2 sigma) words               */
    const uint16_t *interleaved_block_offsets; /* ceil((n+1)/64) * sigma (may be NULL: recomputed) */
    /* SampledSuffixArray, src/sampled_suffix_array.rs:18-23                                       */
    const uint64_t *sampled_suffix_array; /* ceil(n / rate) entries                                */
    uint32_t sampling_rate;
    const uint64_t *text_border_rows;     /* keys of text_border_lookup                            */
    const uint64_t *text_border_positions;/* values of text_border_lookup                          */
    uint64_t num_text_borders;
    /* TexdIdSearchTree, src/text_id_search_tree.rs:6-9                                            */
    const uint64_t *sentinel_indices;
    uint64_t num_texts;
    uint32_t lookup_table_depth;          /* tables are re-derived on the device                   */
} gdx_parts;

typedef struct gdx_index gdx_index;

typedef struct gdx_index_info {
    uint64_t text_len;
    uint64_t num_texts;
    uint32_t num_dense_symbols;
    uint32_t num_searchable_dense_symbols;
    uint32_t storage;
    uint32_t sampling_rate;
    uint32_t lookup_table_depth;
    uint32_t rank_layout;          /* 0: 32 B / 64 positions, 1: generic / 128 positions */
    uint32_t rank_record_bytes;
    uint32_t rank_positions_per_record;
    int32_t device;
    uint64_t image_bytes;          /* size of the device image */
    uint64_t rank_bytes, sample_bytes, lookup_bytes;
    uint64_t num_samples, num_text_borders;
    uint64_t text_bytes;           /* packed text section, 0 if absent */
    uint64_t inverse_sample_bytes; /* sampled inverse suffix array, 0 if absent */
    uint64_t dense_suffix_array_bytes; /* device-only accelerator outside the image, 0 if absent */
    uint64_t seed_table_bytes;         /* device-only accelerator outside the image, 0 if absent */
    uint32_t seed_table_depth;         /* 0 if absent */
    uint32_t reserved;
} gdx_index_info;

/* counters of the last search / locate call on this thread (feeds the roofline arithmetic) */
typedef struct gdx_stats {
    uint64_t queries;
    uint64_t lf_steps;       /* backward-search steps executed (2 rank queries each) */
    uint64_t hits;
    uint64_t walk_steps;     /* LF steps executed by the locate walk                 */
    double kernel_ms_search; /* CUDA-event time of the search kernel(s)              */
    double kernel_ms_locate; /* CUDA-event time of the locate kernels                */
    uint64_t kernel_launches;
    uint64_t verified_queries; /* queries finished by one text comparison instead of LF steps */
} gdx_stats;

uint32_t gdx_abi_version(void);
const char *gdx_last_error_message(void); /* thread local */
uint64_t gdx_last_error_query(void);      /* thread local */
int32_t gdx_device_count(void);

/* ---- construction: FmIndexConfig::construct_index (src/config.rs:63-69 -> src/lib.rs:118-142) -- */
/* texts: concatenated IO bytes; text_offsets: num_texts + 1 entries. */
gdx_status gdx_index_build(const uint8_t *texts, const uint64_t *text_offsets, uint64_t num_texts,
                           const gdx_alphabet *alphabet, const gdx_config *config, gdx_index **out);

/* Upload an index that the reference crate constructed on the host (src/lib.rs:93-100). */
gdx_status gdx_index_create_from_parts(const gdx_parts *parts, int32_t device, gdx_index **out);

/* Same, from the BWT in dense representation instead of the bit planes (n bytes). */
gdx_status gdx_index_create_from_bwt(const uint8_t *bwt, const gdx_parts *parts, int32_t device,
                                     gdx_index **out);

/* Construction utility: suffix array of a dense text (0 = sentinel, ordinary symbol; end of text
 * sorts first -- the libsais convention the reference relies on, construction/mod.rs:88-103),
 * computed on the host (SA-IS) or on the device (prefix doubling). */
gdx_status gdx_suffix_array(const uint8_t *dense_text, uint64_t n, uint32_t num_dense_symbols,
                            uint32_t where /* gdx_construction */, int32_t device, uint64_t *sa_out);

/* Read back parts of a device index (export / interop utilities, not on the search path):
 * the BWT in dense representation (text_len bytes, host buffer) recovered from the rank records
 * with symbol_at (src/text_with_rank_support/condensed.rs:343-362), and count[] (lib.rs:95). */
gdx_status gdx_index_download_bwt(const gdx_index *idx, uint8_t *bwt_out);
gdx_status gdx_index_get_count(const gdx_index *idx, uint64_t *count_out /* num_dense_symbols + 1 */);
/* the sampled suffix array widened to 64 bit (info.num_samples entries,
 * src/sampled_suffix_array.rs:18-23) and text_border_lookup sorted by row (info.num_text_borders) */
gdx_status gdx_index_download_samples(const gdx_index *idx, uint64_t *samples_out);
gdx_status gdx_index_download_text_borders(const gdx_index *idx, uint64_t *rows_out, uint64_t *positions_out);

/* Construction utility: concatenate + densely encode the texts with one 0 sentinel after each text
 * (src/construction/mod.rs:255-308), sentinel positions and count[] (src/construction/mod.rs:318-336).
 * dense_out: total text length + num_texts bytes; sentinels_out: num_texts; count_out: sigma + 1. */
gdx_status gdx_concat_texts(const uint8_t *texts, const uint64_t *text_offsets, uint64_t num_texts,
                            const gdx_alphabet *alphabet, uint8_t *dense_out, uint64_t *sentinels_out,
                            uint64_t *count_out);

void gdx_index_destroy(gdx_index *idx);
gdx_status gdx_index_get_info(const gdx_index *idx, gdx_index_info *out);

/* Dense suffix array accelerator (memory for speed, B200: 180 GB of HBM).  The index keeps the sampled
 * suffix array the configuration asks for (sampled_suffix_array.rs:10-16) -- that is what is saved,
 * exported and replicated.  On top of it a replica can hold SA[row] for EVERY row (4 B per symbol, 8 B
 * beyond 2^32 - 1 symbols), derived on the device from the samples in a fraction of a second:
 * resolving a row (locate, and the text verification of count/locate) is then one load instead of an
 * LF-walk of up to rate - 1 steps plus a load.  Results are identical with and without it.
 * It is built automatically after construction / load / adopt / replicate when it needs at most a
 * quarter of the free device memory (GDX_DENSE_SA=0 never, =1 always; GDX_FLAG_NO_DENSE_SUFFIX_ARRAY
 * per index).  on != 0 builds it now (GDX_ERR_OOM if it does not fit), on == 0 frees it.  Must not run
 * concurrently with queries on the same handle. */
gdx_status gdx_index_set_dense_suffix_array(gdx_index *idx, int32_t on);

/* Seed table accelerator (memory for speed).  The index keeps the lookup tables of the configured depth
 * (lookup_table.rs:19-23) -- that is what is saved, exported and replicated, and what decides every
 * error behaviour.  On top of it a replica can hold ONE level of a deeper lookup table (depth d: ns^d
 * entries of 8 B, 16 B beyond 2^32 - 1 symbols), filled on the device level by level with the same
 * rule as the reference (lookup_table.rs:225-258).  A query whose last d symbols are all searchable starts
 * from that entry instead of the configured table + LF steps; the entry is bit for bit the interval those
 * steps produce (also when it is empty), every other query takes the configured path.  Results, error
 * behaviour and reported lookup_table_depth are unchanged.
 * Built automatically after construction / load / adopt / replicate with the largest d such that
 * ns^d <= text length, if that needs at most a quarter of the free device memory (GDX_SEED_TABLE=0 never,
 * =d that depth; GDX_FLAG_NO_SEED_TABLE per index).  depth > configured depth (re)builds it at that depth now
 * (GDX_ERR_OOM / GDX_ERR_UNSUPPORTED if it does not fit), any smaller depth just frees it.  Must not run concurrently
 * with queries on the same handle. */
gdx_status gdx_index_set_seed_table_depth(gdx_index *idx, int32_t depth);

/* ---- index files: FmIndex::save_to_file / load_from_file (src/lib.rs:296-327) -------------------------
 * The crate serialises its host structs with the `savefile` crate (schema version 0); that byte format
 * is owned by an un-vendored dependency and no reference test pins it, so this is an own versioned
 * container (magic "GDXFILE1" + image header + device image + caller blob), NOT savefile-compatible.
 * `user_data` is an opaque blob of the host language binding (e.g. the alphabet's dense->io table). */
gdx_status gdx_index_save_to_file(const gdx_index *idx, const char *path, const void *user_data,
                                  uint64_t user_bytes);
gdx_status gdx_index_load_from_file(const char *path, int32_t device, gdx_index **out, void *user_data_out,
                                    uint64_t user_capacity, uint64_t *user_bytes_out);

/* ---- replication (no counterpart in the reference) ----------------------------------
 * The device image is one contiguous allocation described by an opaque POD header.  A replica on
 * another GPU (or in another process) is made by copying header + image bytes, e.g. with one
 * ncclBroadcast from rank 0, and adopting them.  `image` must stay valid while the index lives
 * (own_image = 0) or is freed with cudaFree by gdx_index_destroy (own_image != 0). */
uint64_t gdx_index_header_bytes(void);
gdx_status gdx_index_export(const gdx_index *idx, void *header_out, const void **device_image,
                            uint64_t *image_bytes);
gdx_status gdx_index_adopt_image(const void *header, void *device_image, int32_t device,
                                 int32_t own_image, gdx_index **out);
/* single-process convenience: replicas on other devices via peer copies from the source device */
gdx_status gdx_index_replicate(const gdx_index *idx, const int32_t *devices, int32_t n_devices,
                               gdx_index **out_replicas);

/* ---- batched search with host buffers ----------------------------------------------------------*/
/* FmIndex::cursors_for_many_queries (src/lib.rs:241-246): intervals [start,end) in input order. */
gdx_status gdx_cursors_many(const gdx_index *idx, const gdx_queries *queries, uint64_t *starts,
                            uint64_t *ends);
/* FmIndex::count_many (src/lib.rs:155-161) */
gdx_status gdx_count_many(const gdx_index *idx, const gdx_queries *queries, uint64_t *counts);
/* FmIndex::locate_many (src/lib.rs:179-185): CSR result; hit_offsets has nq + 1 entries; the hits
 * of query i are hits[hit_offsets[i] .. hit_offsets[i+1]) in the reference's SA-row order. */
gdx_status gdx_locate_many(const gdx_index *idx, const gdx_queries *queries, uint64_t *hit_offsets,
                           gdx_hit **hits, uint64_t *num_hits);
/* Cursor::locate / FmIndex::locate_interval (src/cursor.rs:71-73, src/lib.rs:187-197) for many
 * cursors at once. */
gdx_status gdx_locate_intervals(const gdx_index *idx, const uint64_t *starts, const uint64_t *ends,
                                uint64_t n, uint64_t *hit_offsets, gdx_hit **hits,
                                uint64_t *num_hits);
void gdx_free_hits(const gdx_index *idx, gdx_hit *hits);
/* Cursor::extend_query_front (src/cursor.rs:34-51) for many cursors: in-place on starts/ends.  On an error
 * status the contents of pinned starts/ends are unspecified (the reference panics); pageable arrays are
 * left untouched. */
gdx_status gdx_extend_many(const gdx_index *idx, uint64_t *starts, uint64_t *ends,
                           const uint8_t *io_symbols, uint64_t n);

/* single-query forms (src/lib.rs:147-149,169-173,217-235); these follow the single-query code path
 * of the reference (the symbol left of an empty lookup interval is still translated). */
gdx_status gdx_cursor_for_query(const gdx_index *idx, const uint8_t *query, uint64_t len,
                                uint64_t *start, uint64_t *end);

/* ---- batched search with device-resident buffers (for pipelines that keep data in HBM) ---------
 * All pointers are device pointers on the index's device; `stream` is a cudaStream_t (NULL = the
 * legacy default stream).  Asynchronous: errors of the kernels are reported through *d_error
 * (device uint64: 0 = none, else 1 + index of the first offending query), which may be NULL. */
gdx_status gdx_cursors_many_device(const gdx_index *idx, const gdx_queries *d_queries,
                                   uint64_t *d_starts, uint64_t *d_ends, uint64_t *d_error,
                                   void *stream);
gdx_status gdx_count_many_device(const gdx_index *idx, const gdx_queries *d_queries,
                                 uint64_t *d_counts, uint64_t *d_error, void *stream);
/* locate rows [start,end) of n intervals whose CSR offsets are already known on the device */
gdx_status gdx_locate_intervals_device(const gdx_index *idx, const uint64_t *d_starts,
                                       const uint64_t *d_ends, uint64_t n,
                                       const uint64_t *d_hit_offsets, uint64_t num_hits,
                                       gdx_hit *d_hits, void *stream);

/* pinned host memory helpers (full-speed asynchronous H2D/D2H) */
gdx_status gdx_host_alloc(uint64_t bytes, void **out);
void gdx_host_free(void *p);

gdx_status gdx_get_stats(gdx_stats *out); /* thread local, last call */

/* measured random-gather ceiling of the device: independent aligned `record_bytes`
 * loads at random offsets over a `table_bytes` table; returns GB/s of requested bytes.
 * chained != 0: the next two addresses of a thread depend on the data just loaded (the access
 * pattern of an LF step) instead of on a counter. */
gdx_status gdx_measure_random_gather(int32_t device, uint64_t table_bytes, uint32_t record_bytes,
                                     uint64_t loads, int32_t chained, double *gbps_out,
                                     double *gloads_out);

#ifdef __cplusplus
}
#endif
#endif /* GENEDEX_B200_H */
