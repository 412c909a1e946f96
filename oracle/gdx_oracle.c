/*
 * gdx_oracle.c -- CPU restatement ("oracle") of genedex's FM-index search path, plain C11.
 *
 * TEST INFRASTRUCTURE ONLY (see gdx_oracle.h).  Citations are file:line under /root/reference.
 * The code is written for obvious correctness first; the batched search and the batched rank keep
 * the reference's pass structure because they double as the CPU baseline in bench.py.
 */
#define _GNU_SOURCE
#include "gdx_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define SUPERBLOCK_SIZE 65536u /* condensed.rs:34,70: u16::MAX + 1 */
#define BLOCK_BITS 64u         /* block.rs:150: Block64::NUM_BITS */
#define BATCH_SIZE 64u         /* lib.rs:115 */

/* ------------------------------------------------------------------------------------------------
 * small helpers
 * ---------------------------------------------------------------------------------------------- */

/* array of the reference's storage integer I: 32-bit for i32/u32, 64-bit for i64 */
typedef struct {
    void *p;
    int wide;
    uint64_t len;
} iarray;

static int iarray_alloc(iarray *a, uint64_t len, int storage) {
    a->wide = (storage == GDXO_I64);
    a->len = len;
    a->p = calloc(len ? len : 1, a->wide ? 8 : 4);
    return a->p ? 0 : -1;
}
static inline uint64_t iarray_get(const iarray *a, uint64_t i) {
    return a->wide ? ((const uint64_t *)a->p)[i] : (uint64_t)((const uint32_t *)a->p)[i];
}
static inline void iarray_set(iarray *a, uint64_t i, uint64_t v) {
    if (a->wide)
        ((uint64_t *)a->p)[i] = v;
    else
        ((uint32_t *)a->p)[i] = (uint32_t)v;
}

static uint64_t storage_max(int storage) {
    switch (storage) {
    case GDXO_I32: return 0x7fffffffull;
    case GDXO_U32: return 0xffffffffull;
    default: return 0x7fffffffffffffffull;
    }
}

/* condensed.rs:417-419 */
static uint32_t ilog2_ceil_for_nonzero(uint64_t value) {
    uint32_t lz = (uint32_t)__builtin_clzll(value);
    uint32_t pow2 = (value & (value - 1)) == 0;
    return 64u - lz - pow2;
}

static uint64_t div_ceil(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

int gdxo_online_cores(void) {
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

typedef void (*range_fn)(uint64_t begin, uint64_t end, int tid, void *ctx);
typedef struct {
    range_fn fn;
    void *ctx;
    uint64_t begin, end;
    int tid;
} range_job;
static void *range_trampoline(void *arg) {
    range_job *j = (range_job *)arg;
    j->fn(j->begin, j->end, j->tid, j->ctx);
    return NULL;
}
/* contiguous chunks of [0,n) rounded to `granule`, one per thread */
static void parallel_ranges(int nthreads, uint64_t n, uint64_t granule, range_fn fn, void *ctx) {
    if (nthreads <= 0) nthreads = gdxo_online_cores();
    if (granule == 0) granule = 1;
    uint64_t ngran = div_ceil(n, granule);
    if ((uint64_t)nthreads > ngran) nthreads = ngran ? (int)ngran : 1;
    if (nthreads <= 1) {
        fn(0, n, 0, ctx);
        return;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    range_job *jobs = (range_job *)malloc(sizeof(range_job) * (size_t)nthreads);
    uint64_t per = div_ceil(ngran, (uint64_t)nthreads) * granule;
    for (int t = 0; t < nthreads; ++t) {
        uint64_t b = per * (uint64_t)t, e = b + per;
        if (b > n) b = n;
        if (e > n) e = n;
        jobs[t] = (range_job){fn, ctx, b, e, t};
        pthread_create(&th[t], NULL, range_trampoline, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/* ------------------------------------------------------------------------------------------------
 * rank structure: CondensedTextWithRankSupport<I, Block64>  (condensed.rs:24-30)
 * ---------------------------------------------------------------------------------------------- */

struct gdxo_rank {
    uint64_t text_len;
    uint32_t alphabet_size;
    uint32_t num_bits;           /* ilog2_ceil(alphabet_size) planes per block */
    uint64_t *blocks;            /* interleaved_blocks: [block][plane]            */
    uint64_t n_blocks_words;
    uint16_t *block_offsets;     /* interleaved_block_offsets: [block][symbol]    */
    uint64_t n_block_offsets;
    iarray superblock_offsets;   /* interleaved_superblock_offsets: [sb][symbol]  */
};

typedef struct {
    gdxo_rank *r;
    const uint8_t *text;
    uint64_t n;
} rank_fill_ctx;

/* condensed.rs:365-415 fill_superblock, restated by definition: block_offsets[k][c] = #c in
 * [superblock start, 64k), planes = bit p of each symbol; the per-superblock totals are left in
 * superblock_offsets for the accumulation pass (condensed.rs:104-116). */
static void rank_fill_range(uint64_t sb_begin, uint64_t sb_end, int tid, void *vctx) {
    (void)tid;
    rank_fill_ctx *c = (rank_fill_ctx *)vctx;
    gdxo_rank *r = c->r;
    const uint32_t sigma = r->alphabet_size, nb = r->num_bits;
    const uint64_t len = c->n + 1; /* condensed.rs:69: n + 1 positions are addressable */
    const uint64_t nblocks = div_ceil(len, BLOCK_BITS);
    uint64_t *sum = (uint64_t *)calloc(sigma, sizeof(uint64_t));
    for (uint64_t sb = sb_begin; sb < sb_end; ++sb) {
        memset(sum, 0, sizeof(uint64_t) * sigma);
        uint64_t blk0 = sb * (SUPERBLOCK_SIZE / BLOCK_BITS);
        uint64_t blk1 = blk0 + SUPERBLOCK_SIZE / BLOCK_BITS;
        if (blk1 > nblocks) blk1 = nblocks;
        for (uint64_t k = blk0; k < blk1; ++k) {
            for (uint32_t s = 0; s < sigma; ++s)
                r->block_offsets[k * sigma + s] = (uint16_t)sum[s]; /* never wraps, see header */
            uint64_t p0 = k * BLOCK_BITS, p1 = p0 + BLOCK_BITS;
            if (p1 > c->n) p1 = c->n;
            for (uint64_t p = p0; p < p1; ++p) {
                uint8_t sym = c->text[p];
                sum[sym]++;
                for (uint32_t b = 0; b < nb; ++b)
                    r->blocks[k * nb + b] |= (uint64_t)((sym >> b) & 1u) << (p - p0);
            }
        }
        for (uint32_t s = 0; s < sigma; ++s)
            iarray_set(&r->superblock_offsets, sb * sigma + s, sum[s]);
    }
    free(sum);
}

static gdxo_rank *rank_construct_mt(const uint8_t *text, uint64_t n, uint32_t sigma, int storage,
                                    int nthreads) {
    if (sigma < 2) return NULL; /* condensed.rs:64 assert!(alphabet_size >= 2) */
    gdxo_rank *r = (gdxo_rank *)calloc(1, sizeof(*r));
    if (!r) return NULL;
    r->text_len = n;
    r->alphabet_size = sigma;
    r->num_bits = ilog2_ceil_for_nonzero(sigma);
    const uint64_t len = n + 1;
    const uint64_t nblocks = div_ceil(len, BLOCK_BITS);
    const uint64_t nsb = div_ceil(len, SUPERBLOCK_SIZE);
    r->n_blocks_words = nblocks * r->num_bits;       /* condensed.rs:72 */
    r->n_block_offsets = nblocks * sigma;            /* condensed.rs:73 */
    r->blocks = (uint64_t *)calloc(r->n_blocks_words, 8);
    r->block_offsets = (uint16_t *)calloc(r->n_block_offsets, 2);
    if (!r->blocks || !r->block_offsets || iarray_alloc(&r->superblock_offsets, nsb * sigma, storage)) {
        gdxo_rank_free(r);
        return NULL;
    }
    rank_fill_ctx ctx = {r, text, n};
    parallel_ranges(nthreads, nsb, 1, rank_fill_range, &ctx);
    /* condensed.rs:104-116: exclusive prefix sum over superblocks, single thread */
    uint64_t *acc = (uint64_t *)calloc(sigma, 8);
    for (uint64_t sb = 0; sb < nsb; ++sb)
        for (uint32_t s = 0; s < sigma; ++s) {
            uint64_t t = iarray_get(&r->superblock_offsets, sb * sigma + s);
            iarray_set(&r->superblock_offsets, sb * sigma + s, acc[s]);
            acc[s] += t;
        }
    free(acc);
    return r;
}

gdxo_rank *gdxo_rank_construct(const uint8_t *dense_text, uint64_t n, uint32_t sigma, int storage) {
    return rank_construct_mt(dense_text, n, sigma, storage, 1);
}

void gdxo_rank_free(gdxo_rank *r) {
    if (!r) return;
    free(r->blocks);
    free(r->block_offsets);
    free(r->superblock_offsets.p);
    free(r);
}

/* block.rs:172-175 Block64::count_ones_before */
static inline uint64_t count_ones_before(uint64_t data, uint64_t idx) {
    uint64_t masked = data & ~(~0ull << idx); /* idx < 64 always (idx % 64) */
    return (uint64_t)__builtin_popcountll(masked);
}

/* condensed.rs:291-341 rank_unchecked */
uint64_t gdxo_rank_query(const gdxo_rank *r, uint8_t symbol, uint64_t idx) {
    const uint32_t sigma = r->alphabet_size, nb = r->num_bits;
    uint64_t sbo = iarray_get(&r->superblock_offsets, (idx / SUPERBLOCK_SIZE) * sigma + symbol);
    uint64_t bo = r->block_offsets[(idx / BLOCK_BITS) * sigma + symbol];
    const uint64_t *planes = r->blocks + (idx / BLOCK_BITS) * nb;
    uint8_t s = symbol;
    uint64_t acc = planes[0];
    if ((s & 1u) == 0) acc = ~acc;
    for (uint32_t b = 1; b < nb; ++b) {
        s >>= 1;
        uint64_t blk = planes[b];
        if ((s & 1u) == 0) blk = ~blk;
        acc &= blk;
    }
    return sbo + bo + count_ones_before(acc, idx % BLOCK_BITS);
}

/* condensed.rs:343-362 symbol_at */
uint8_t gdxo_rank_symbol_at(const gdxo_rank *r, uint64_t idx) {
    const uint32_t nb = r->num_bits;
    const uint64_t *planes = r->blocks + (idx / BLOCK_BITS) * nb;
    uint8_t symbol = 0;
    for (uint32_t b = 0; b < nb; ++b)
        symbol |= (uint8_t)(((planes[b] >> (idx % BLOCK_BITS)) & 1u) << b);
    return symbol;
}

/* condensed.rs:137-287 replace_many_interval_borders_with_ranks_unchecked: the same seven passes
 * over the batch so that the loads of one pass are independent of each other. */
void gdxo_rank_batch(const gdxo_rank *r, const uint8_t *symbols, uint64_t *starts, uint64_t *ends,
                     uint32_t nq) {
    const uint32_t sigma = r->alphabet_size, nb = r->num_bits;
    uint64_t sb_s[BATCH_SIZE], sb_e[BATCH_SIZE], bo_s[BATCH_SIZE], bo_e[BATCH_SIZE];
    const uint64_t *pl_s[BATCH_SIZE], *pl_e[BATCH_SIZE];
    uint64_t acc_s[BATCH_SIZE], acc_e[BATCH_SIZE];
    /* :156-160 superblock offset indices */
    for (uint32_t i = 0; i < nq; ++i) {
        sb_s[i] = (starts[i] / SUPERBLOCK_SIZE) * sigma + symbols[i];
        sb_e[i] = (ends[i] / SUPERBLOCK_SIZE) * sigma + symbols[i];
    }
    /* :164-181 superblock loads */
    if (r->superblock_offsets.wide) {
        const uint64_t *sbo = (const uint64_t *)r->superblock_offsets.p;
        for (uint32_t i = 0; i < nq; ++i) {
            sb_s[i] = sbo[sb_s[i]];
            sb_e[i] = sbo[sb_e[i]];
        }
    } else {
        const uint32_t *sbo = (const uint32_t *)r->superblock_offsets.p;
        for (uint32_t i = 0; i < nq; ++i) {
            sb_s[i] = sbo[sb_s[i]];
            sb_e[i] = sbo[sb_e[i]];
        }
    }
    /* :184-187 block offset indices */
    for (uint32_t i = 0; i < nq; ++i) {
        bo_s[i] = (starts[i] / BLOCK_BITS) * sigma + symbols[i];
        bo_e[i] = (ends[i] / BLOCK_BITS) * sigma + symbols[i];
    }
    /* :191-203 block offset loads */
    for (uint32_t i = 0; i < nq; ++i) {
        bo_s[i] = r->block_offsets[bo_s[i]];
        bo_e[i] = r->block_offsets[bo_e[i]];
    }
    /* :210-218 plane slices */
    for (uint32_t i = 0; i < nq; ++i) {
        pl_s[i] = r->blocks + (starts[i] / BLOCK_BITS) * nb;
        pl_e[i] = r->blocks + (ends[i] / BLOCK_BITS) * nb;
    }
    /* :222-242 first plane */
    for (uint32_t i = 0; i < nq; ++i) {
        acc_s[i] = pl_s[i][0];
        acc_e[i] = pl_e[i][0];
    }
    /* :244-273 negate / and over remaining planes */
    for (uint32_t i = 0; i < nq; ++i) {
        uint8_t s = symbols[i];
        if ((s & 1u) == 0) {
            acc_s[i] = ~acc_s[i];
            acc_e[i] = ~acc_e[i];
        }
        for (uint32_t b = 1; b < nb; ++b) {
            s >>= 1;
            uint64_t x = pl_s[i][b], y = pl_e[i][b];
            if ((s & 1u) == 0) {
                x = ~x;
                y = ~y;
            }
            acc_s[i] &= x;
            acc_e[i] &= y;
        }
    }
    /* :275-286 mask, popcount, sum */
    for (uint32_t i = 0; i < nq; ++i) {
        starts[i] = sb_s[i] + bo_s[i] + count_ones_before(acc_s[i], starts[i] % BLOCK_BITS);
        ends[i] = sb_e[i] + bo_e[i] + count_ones_before(acc_e[i], ends[i] % BLOCK_BITS);
    }
}

/* ------------------------------------------------------------------------------------------------
 * text id search tree (text_id_search_tree.rs)
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    int64_t *nodes; /* Node{isize}: >= 0 threshold, < 0 = !text_id  (:126-154) */
    uint64_t n_nodes;
} text_tree;

static uint64_t next_pow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

/* text_id_search_tree.rs:67-109 add_nodes */
static void tree_add_nodes(int64_t *nodes, uint64_t cur, const uint64_t *indices, uint64_t num,
                           uint64_t offset, uint64_t *max_used) {
    if (cur > *max_used) *max_used = cur;
    if (num == 1) {
        nodes[cur] = ~(int64_t)offset; /* new_leaf :144-148 */
        return;
    }
    uint64_t cur_offset = ((num & (num - 1)) == 0) ? num / 2 : next_pow2(num) / 2;
    nodes[cur] = (int64_t)indices[cur_offset - 1]; /* threshold = *left.last() */
    tree_add_nodes(nodes, cur * 2 + 1, indices, cur_offset, offset, max_used);
    tree_add_nodes(nodes, (cur + 1) * 2, indices + cur_offset, num - cur_offset,
                   offset + cur_offset, max_used);
}

/* text_id_search_tree.rs:13-33 */
static int tree_build(text_tree *t, const uint64_t *sentinels, uint64_t ntexts) {
    uint64_t cap = next_pow2(ntexts) * 2 - 1;
    t->nodes = (int64_t *)calloc(cap, sizeof(int64_t));
    if (!t->nodes) return -1;
    uint64_t max_used = 0;
    tree_add_nodes(t->nodes, 0, sentinels, ntexts, 0, &max_used);
    t->n_nodes = max_used + 1;
    return 0;
}

/* text_id_search_tree.rs:50-64 */
static uint64_t tree_lookup_text_id(const text_tree *t, uint64_t pos) {
    uint64_t cur = 0;
    while (t->nodes[cur] >= 0)
        cur = (pos <= (uint64_t)t->nodes[cur]) ? cur * 2 + 1 : (cur + 1) * 2;
    return (uint64_t)(~t->nodes[cur]);
}

uint64_t gdxo_tree_lookup(const uint64_t *sentinel_indices, uint64_t ntexts, uint64_t pos) {
    text_tree t;
    if (tree_build(&t, sentinel_indices, ntexts)) return (uint64_t)-1;
    uint64_t id = tree_lookup_text_id(&t, pos);
    free(t.nodes);
    return id;
}

/* ------------------------------------------------------------------------------------------------
 * the index (lib.rs:93-100)
 * ---------------------------------------------------------------------------------------------- */

struct gdxo_index {
    /* alphabet (alphabet.rs:24-28) */
    uint8_t io_to_dense[256];
    uint32_t sigma;          /* num_dense_symbols, incl. sentinel */
    uint32_t num_searchable; /* num_searchable_dense_symbols      */
    int storage;
    uint64_t n;              /* total text len incl. sentinels    */
    uint64_t *count;         /* sigma + 1 */
    uint64_t freq[257]; /* 256 table entries + a zero so that count[] can hold sigma+1 sums */
    gdxo_rank *rank;
    /* sampled suffix array (sampled_suffix_array.rs:18-23) */
    iarray samples;
    uint32_t sampling_rate;
    uint64_t *border_rows, *border_pos; /* text_border_lookup, sorted by row */
    uint64_t n_border;
    /* text ids */
    text_tree tree;
    uint64_t *sentinels;
    uint64_t ntexts;
    /* lookup tables (lookup_table.rs:19-23): tables[d] has ns^d (start,end) pairs */
    uint32_t n_tables; /* = max_depth + 1 once filled */
    uint32_t tables_cap;
    iarray *tables;
    uint64_t *factors;
    /* kept only when built from texts */
    uint8_t *text;
    int64_t *sa;
    uint8_t *bwt;
};

/* ---- suffix array: prefix doubling (Manber-Myers / Larsson-Sadakane flavour) -------------------
 * Ordering convention of libsais as used at construction/mod.rs:88-103: plain lexicographic order
 * of the suffixes of the dense text, the 0 sentinels are ordinary symbols, a suffix that is a
 * proper prefix of another one is smaller (end of text < every symbol).  The suffix array is
 * unique, so any correct SACA gives the reference's array. */

typedef struct {
    const int64_t *rank;
    int64_t h, n;
} sa_cmp_ctx;

static int sa_cmp(const void *a, const void *b, void *vctx) {
    const sa_cmp_ctx *c = (const sa_cmp_ctx *)vctx;
    int64_t pa = *(const int64_t *)a + c->h, pb = *(const int64_t *)b + c->h;
    int64_t ka = pa < c->n ? c->rank[pa] : -1, kb = pb < c->n ? c->rank[pb] : -1;
    return (ka > kb) - (ka < kb);
}

static int build_suffix_array(const uint8_t *T, int64_t n, uint32_t sigma, int64_t *sa) {
    if (n == 0) return 0;
    const uint64_t base = (uint64_t)sigma + 1; /* code 0 = beyond the end */
    int64_t k = 1;
    uint64_t nbuckets = base;
    uint64_t bucket_cap = (uint64_t)n * 4 < (1ull << 24) ? (uint64_t)n * 4 : (1ull << 24);
    while (nbuckets * base <= bucket_cap && k < 16) {
        nbuckets *= base;
        ++k;
    }
    uint64_t top = nbuckets / base; /* base^(k-1) */
    int64_t *rank = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t *nr = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    uint64_t *bucket = (uint64_t *)calloc(nbuckets + 1, sizeof(uint64_t));
    uint32_t *key = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
    if (!rank || !nr || !bucket || !key) {
        free(rank); free(nr); free(bucket); free(key);
        return -1;
    }
    uint64_t kk = 0; /* key of the (empty) suffix at n: all digits 0 */
    for (int64_t i = n - 1; i >= 0; --i) {
        kk = ((uint64_t)T[i] + 1) * top + kk / base;
        key[i] = (uint32_t)kk;
        bucket[kk + 1]++;
    }
    for (uint64_t b = 0; b < nbuckets; ++b) bucket[b + 1] += bucket[b];
    for (int64_t i = 0; i < n; ++i) rank[i] = (int64_t)bucket[key[i]]; /* group head index */
    for (int64_t i = 0; i < n; ++i) sa[bucket[key[i]]++] = i;
    free(bucket);
    free(key);

    /* non-singleton groups as (begin,end) pairs, double buffered */
    int64_t cap = 1024, ng = 0;
    int64_t *groups = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)cap);
    for (int64_t i = 0; i < n;) {
        int64_t j = i + 1;
        while (j < n && rank[sa[j]] == i) ++j;
        if (j - i > 1) {
            if (ng == cap) {
                cap *= 2;
                groups = (int64_t *)realloc(groups, sizeof(int64_t) * 2 * (size_t)cap);
            }
            groups[2 * ng] = i;
            groups[2 * ng + 1] = j;
            ++ng;
        }
        i = j;
    }
    for (int64_t h = k; ng > 0; h *= 2) {
        sa_cmp_ctx ctx = {rank, h, n};
        /* phase 1: sort every open group by the rank of the suffix h further on; compute the
         * refined group heads into nr (old ranks are still needed by other groups) */
        for (int64_t g = 0; g < ng; ++g) {
            int64_t b = groups[2 * g], e = groups[2 * g + 1];
            qsort_r(sa + b, (size_t)(e - b), sizeof(int64_t), sa_cmp, &ctx);
            int64_t head = b;
            for (int64_t x = b; x < e; ++x) {
                if (x > b && sa_cmp(&sa[x - 1], &sa[x], &ctx) != 0) head = x;
                nr[sa[x]] = head;
            }
        }
        /* phase 2: publish the new ranks and collect the groups that are still open */
        int64_t cap2 = ng > 16 ? ng : 16, ng2 = 0;
        int64_t *groups2 = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)cap2);
        for (int64_t g = 0; g < ng; ++g) {
            int64_t b = groups[2 * g], e = groups[2 * g + 1];
            for (int64_t x = b; x < e; ++x) rank[sa[x]] = nr[sa[x]];
        }
        for (int64_t g = 0; g < ng; ++g) {
            int64_t b = groups[2 * g], e = groups[2 * g + 1];
            for (int64_t i = b; i < e;) {
                int64_t j = i + 1;
                while (j < e && rank[sa[j]] == i) ++j;
                if (j - i > 1) {
                    if (ng2 == cap2) {
                        cap2 *= 2;
                        groups2 = (int64_t *)realloc(groups2, sizeof(int64_t) * 2 * (size_t)cap2);
                    }
                    groups2[2 * ng2] = i;
                    groups2[2 * ng2 + 1] = j;
                    ++ng2;
                }
                i = j;
            }
        }
        free(groups);
        groups = groups2;
        ng = ng2;
    }
    free(groups);
    free(rank);
    free(nr);
    return 0;
}

/* ---- lookup tables (lookup_table.rs) ---------------------------------------------------------- */

static inline uint32_t max_depth(const gdxo_index *idx) { return idx->n_tables - 1; } /* :142-144 */

/* lookup_table.rs:68-113,147-161: first suffix symbol is the least significant digit */
static int compute_lookup_idx(const gdxo_index *idx, const uint8_t *suffix, uint64_t d, int translate,
                              uint64_t *out) {
    uint64_t v = 0;
    for (uint64_t j = 0; j < d; ++j) {
        uint8_t dense = suffix[j];
        if (translate) {
            dense = idx->io_to_dense[suffix[j]];
            if (dense == 0) return GDXO_PANIC_INVALID_SYMBOL; /* alphabet.rs:195-198 */
        }
        v += (uint64_t)(uint8_t)(dense - 1) * idx->factors[j]; /* sentinel is not in the table */
    }
    *out = v;
    return GDXO_OK;
}

/* lookup_table.rs:64-66,215-222 */
static int lookup_idx(const gdxo_index *idx, uint64_t depth, uint64_t i, uint64_t *start,
                      uint64_t *end) {
    const iarray *t = &idx->tables[depth];
    if (i >= t->len / 2) return GDXO_PANIC_LOOKUP_OOB;
    *start = iarray_get(t, 2 * i);
    *end = iarray_get(t, 2 * i + 1);
    return GDXO_OK;
}

/* lib.rs:273-275 */
static inline uint64_t lf_mapping_step(const gdxo_index *idx, uint8_t symbol, uint64_t i) {
    return idx->count[symbol] + gdxo_rank_query(idx->rank, symbol, i);
}

/* cursor.rs:40-51 */
static inline void extend_front_dense(const gdxo_index *idx, uint8_t symbol, uint64_t *start,
                                      uint64_t *end) {
    if (*start != *end) {
        uint64_t s = lf_mapping_step(idx, symbol, *start);
        uint64_t e = lf_mapping_step(idx, symbol, *end);
        *start = s;
        *end = e;
    }
}

/* lib.rs:217-235 (translate = 1) and lib.rs:248-271 (translate = 0) */
static int cursor_for_query_impl(const gdxo_index *idx, const uint8_t *q, uint64_t m, int translate,
                                 uint64_t *start, uint64_t *end) {
    uint64_t depth = m < max_depth(idx) ? m : max_depth(idx); /* lib.rs:277-281 */
    uint64_t suffix_idx = m - depth;
    uint64_t li;
    int rc = compute_lookup_idx(idx, q + suffix_idx, depth, translate, &li);
    if (rc) return rc;
    rc = lookup_idx(idx, depth, li, start, end);
    if (rc) return rc;
    for (uint64_t r = suffix_idx; r-- > 0;) {
        uint8_t symbol = q[r];
        if (translate) {
            symbol = idx->io_to_dense[q[r]];
            if (symbol == 0) return GDXO_PANIC_INVALID_SYMBOL;
        }
        extend_front_dense(idx, symbol, start, end);
        if (*end - *start == 0) break;
    }
    return GDXO_OK;
}

int gdxo_cursor_for_query(const gdxo_index *idx, const uint8_t *q, uint64_t m, uint64_t *start,
                          uint64_t *end) {
    return cursor_for_query_impl(idx, q, m, 1, start, end);
}

int gdxo_extend_query_front(const gdxo_index *idx, uint8_t io_symbol, uint64_t *start,
                            uint64_t *end) {
    uint8_t symbol = idx->io_to_dense[io_symbol];
    if (symbol == 0) return GDXO_PANIC_INVALID_SYMBOL;
    extend_front_dense(idx, symbol, start, end);
    return GDXO_OK;
}

typedef struct {
    gdxo_index *idx;
    uint32_t depth;
    uint64_t num_values;
    int rc;
} table_fill_ctx;

/* lookup_table.rs:225-258 fill_table: every dense k-mer of length `depth` over symbols 1..=ns is
 * searched with the tables built so far (max_depth == depth-1 at this point) */
static void table_fill_range(uint64_t b, uint64_t e, int tid, void *vctx) {
    (void)tid;
    table_fill_ctx *c = (table_fill_ctx *)vctx;
    gdxo_index *idx = c->idx;
    uint8_t query[64];
    const uint64_t ns = idx->num_searchable;
    for (uint64_t v = b; v < e; ++v) {
        uint64_t t = v;
        for (uint32_t j = 0; j < c->depth; ++j) {
            query[j] = (uint8_t)(t % ns + 1); /* +1 to offset sentinel */
            t /= ns;
        }
        uint64_t s = 0, en = 0;
        int rc = cursor_for_query_impl(idx, query, c->depth, 0, &s, &en);
        if (rc) c->rc = rc;
        iarray_set(&idx->tables[c->depth], 2 * v, s);
        iarray_set(&idx->tables[c->depth], 2 * v + 1, en);
    }
}

/* lookup_table.rs:163-181 fill_lookup_tables */
static int fill_lookup_tables(gdxo_index *idx, uint32_t depth_max, int nthreads) {
    const uint64_t ns = idx->num_searchable;
    idx->factors = (uint64_t *)calloc(depth_max + 1, 8);
    idx->tables = (iarray *)calloc(depth_max + 1, sizeof(iarray));
    if (!idx->factors || !idx->tables) return GDXO_ERR_ALLOC;
    idx->tables_cap = depth_max + 1;
    uint64_t f = 1;
    for (uint32_t d = 0; d <= depth_max; ++d) {
        idx->factors[d] = f;
        f *= ns;
    }
    idx->n_tables = 0;
    for (uint32_t d = 0; d <= depth_max; ++d) {
        uint64_t num_values = idx->factors[d];
        if (iarray_alloc(&idx->tables[d], 2 * num_values, idx->storage)) return GDXO_ERR_ALLOC;
        if (d == 0) {
            iarray_set(&idx->tables[0], 0, 0);
            iarray_set(&idx->tables[0], 1, idx->n); /* :205-209 */
        } else {
            table_fill_ctx ctx = {idx, d, num_values, 0};
            parallel_ranges(nthreads, num_values, 4096, table_fill_range, &ctx);
            if (ctx.rc) return ctx.rc;
        }
        idx->n_tables = d + 1; /* tables.push: max_depth() grows only after the table is complete */
    }
    return GDXO_OK;
}

/* ---- construction ----------------------------------------------------------------------------- */

static int finish_index(gdxo_index *idx, uint32_t lookup_depth, int nthreads) {
    if (tree_build(&idx->tree, idx->sentinels, idx->ntexts)) return GDXO_ERR_ALLOC;
    return fill_lookup_tables(idx, lookup_depth, nthreads);
}

static int check_config(uint32_t sigma, uint32_t num_searchable, uint32_t sampling_rate,
                        uint64_t ntexts) {
    if (sampling_rate == 0) return GDXO_PANIC_BAD_CONFIG;                 /* config.rs:28 */
    if (sigma < 2 || sigma > 256) return GDXO_PANIC_BAD_CONFIG;           /* alphabet.rs:166-174 */
    if (num_searchable < 1 || num_searchable > sigma - 1) return GDXO_PANIC_BAD_CONFIG; /* :186-189 */
    if (ntexts == 0) return GDXO_PANIC_BAD_CONFIG; /* construction/mod.rs:300 "at least one texts" */
    return GDXO_OK;
}

int gdxo_build(const uint8_t *texts, const uint64_t *text_offsets, uint64_t ntexts,
               const uint8_t io_to_dense[256], uint32_t sigma, uint32_t num_searchable,
               uint32_t sampling_rate, uint32_t lookup_depth, int storage, gdxo_index **out) {
    *out = NULL;
    int rc = check_config(sigma, num_searchable, sampling_rate, ntexts);
    if (rc) return rc;
    gdxo_index *idx = (gdxo_index *)calloc(1, sizeof(*idx));
    if (!idx) return GDXO_ERR_ALLOC;
    memcpy(idx->io_to_dense, io_to_dense, 256);
    idx->sigma = sigma;
    idx->num_searchable = num_searchable;
    idx->storage = storage;
    idx->sampling_rate = sampling_rate;
    idx->ntexts = ntexts;

    /* construction/mod.rs:255-308 create_concatenated_densely_encoded_text */
    const uint64_t total = text_offsets[ntexts] - text_offsets[0];
    const uint64_t n = total + ntexts;
    idx->n = n;
    idx->text = (uint8_t *)calloc(n ? n : 1, 1);
    idx->sentinels = (uint64_t *)calloc(ntexts, 8);
    if (!idx->text || !idx->sentinels) { gdxo_free(idx); return GDXO_ERR_ALLOC; }
    uint64_t w = 0;
    for (uint64_t t = 0; t < ntexts; ++t) {
        for (uint64_t p = text_offsets[t]; p < text_offsets[t + 1]; ++p) {
            uint8_t d = io_to_dense[texts[p]];
            if (d == 0) { gdxo_free(idx); return GDXO_PANIC_INVALID_SYMBOL; }
            idx->text[w++] = d;
            idx->freq[d]++;
        }
        idx->sentinels[t] = w; /* :267-274 */
        idx->text[w++] = 0;
    }
    idx->freq[0] = ntexts; /* :302 */
    if (n > storage_max(storage)) { gdxo_free(idx); return GDXO_PANIC_TEXT_TOO_LONG; } /* :34 */

    /* construction/mod.rs:318-336 frequency_table_to_count: exclusive prefix sums, sigma+1 entries */
    idx->count = (uint64_t *)calloc(sigma + 1, 8);
    uint64_t sum = 0;
    for (uint32_t s = 0; s <= sigma; ++s) {
        idx->count[s] = sum;
        sum += idx->freq[s];
    }

    idx->sa = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
    idx->bwt = (uint8_t *)malloc(n ? n : 1);
    if (!idx->sa || !idx->bwt || build_suffix_array(idx->text, (int64_t)n, sigma, idx->sa)) {
        gdxo_free(idx);
        return GDXO_ERR_ALLOC;
    }
    /* bwt.rs:93-116: BWT[i] = T[SA[i]-1], SA[i]==0 -> T[n-1]; border lookup for every BWT[i]==0 */
    idx->border_rows = (uint64_t *)calloc(ntexts, 8);
    idx->border_pos = (uint64_t *)calloc(ntexts, 8);
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t ti = (uint64_t)idx->sa[i];
        uint64_t src = ti > 0 ? ti : n;
        idx->bwt[i] = idx->text[src - 1];
        if (idx->bwt[i] == 0) {
            idx->border_rows[idx->n_border] = i;
            idx->border_pos[idx->n_border] = ti;
            idx->n_border++;
        }
    }
    /* sampled_suffix_array.rs:27-54: keep SA[i] for i % s == 0 */
    uint64_t ns = div_ceil(n, sampling_rate);
    if (iarray_alloc(&idx->samples, ns, storage)) { gdxo_free(idx); return GDXO_ERR_ALLOC; }
    for (uint64_t i = 0, k = 0; i < n; i += sampling_rate, ++k)
        iarray_set(&idx->samples, k, (uint64_t)idx->sa[i]);

    idx->rank = rank_construct_mt(idx->bwt, n, sigma, storage, 1);
    if (!idx->rank) { gdxo_free(idx); return GDXO_ERR_ALLOC; }
    rc = finish_index(idx, lookup_depth, 1);
    if (rc) { gdxo_free(idx); return rc; }
    *out = idx;
    return GDXO_OK;
}

int gdxo_from_parts(const uint8_t *bwt, uint64_t n, const uint8_t io_to_dense[256], uint32_t sigma,
                    uint32_t num_searchable, const uint64_t *count, const uint64_t *sentinel_indices,
                    uint64_t ntexts, const uint64_t *sampled_sa, uint64_t n_samples,
                    uint32_t sampling_rate, const uint64_t *border_rows, const uint64_t *border_pos,
                    uint64_t n_border, uint32_t lookup_depth, int storage, int nthreads,
                    gdxo_index **out) {
    *out = NULL;
    int rc = check_config(sigma, num_searchable, sampling_rate, ntexts);
    if (rc) return rc;
    if (n > storage_max(storage)) return GDXO_PANIC_TEXT_TOO_LONG;
    gdxo_index *idx = (gdxo_index *)calloc(1, sizeof(*idx));
    if (!idx) return GDXO_ERR_ALLOC;
    memcpy(idx->io_to_dense, io_to_dense, 256);
    idx->sigma = sigma;
    idx->num_searchable = num_searchable;
    idx->storage = storage;
    idx->sampling_rate = sampling_rate;
    idx->ntexts = ntexts;
    idx->n = n;
    idx->count = (uint64_t *)calloc(sigma + 1, 8);
    idx->sentinels = (uint64_t *)calloc(ntexts, 8);
    if (!idx->count || !idx->sentinels) { gdxo_free(idx); return GDXO_ERR_ALLOC; }
    memcpy(idx->count, count, 8 * (sigma + 1));
    memcpy(idx->sentinels, sentinel_indices, 8 * ntexts);
    if (sampled_sa) {
        if (iarray_alloc(&idx->samples, n_samples, storage)) { gdxo_free(idx); return GDXO_ERR_ALLOC; }
        for (uint64_t k = 0; k < n_samples; ++k) iarray_set(&idx->samples, k, sampled_sa[k]);
    }
    if (n_border) {
        idx->border_rows = (uint64_t *)calloc(n_border, 8);
        idx->border_pos = (uint64_t *)calloc(n_border, 8);
        /* insertion sort by row (there is one entry per text) */
        for (uint64_t i = 0; i < n_border; ++i) {
            uint64_t j = i;
            while (j > 0 && idx->border_rows[j - 1] > border_rows[i]) {
                idx->border_rows[j] = idx->border_rows[j - 1];
                idx->border_pos[j] = idx->border_pos[j - 1];
                --j;
            }
            idx->border_rows[j] = border_rows[i];
            idx->border_pos[j] = border_pos[i];
        }
        idx->n_border = n_border;
    }
    idx->rank = rank_construct_mt(bwt, n, sigma, storage, nthreads);
    if (!idx->rank) { gdxo_free(idx); return GDXO_ERR_ALLOC; }
    rc = finish_index(idx, lookup_depth, nthreads);
    if (rc) { gdxo_free(idx); return rc; }
    *out = idx;
    return GDXO_OK;
}

void gdxo_free(gdxo_index *idx) {
    if (!idx) return;
    free(idx->count);
    gdxo_rank_free(idx->rank);
    free(idx->samples.p);
    free(idx->border_rows);
    free(idx->border_pos);
    free(idx->tree.nodes);
    free(idx->sentinels);
    if (idx->tables)
        for (uint32_t d = 0; d < idx->tables_cap; ++d) free(idx->tables[d].p);
    free(idx->tables);
    free(idx->factors);
    free(idx->text);
    free(idx->sa);
    free(idx->bwt);
    free(idx);
}

/* ---- introspection ---------------------------------------------------------------------------- */
uint64_t gdxo_text_len(const gdxo_index *idx) { return idx->n; }
uint64_t gdxo_num_texts(const gdxo_index *idx) { return idx->ntexts; }
const uint8_t *gdxo_dense_text(const gdxo_index *idx) { return idx->text; }
const int64_t *gdxo_suffix_array(const gdxo_index *idx) { return idx->sa; }
const uint8_t *gdxo_bwt(const gdxo_index *idx) { return idx->bwt; }
const uint64_t *gdxo_count_array(const gdxo_index *idx) { return idx->count; }
const uint64_t *gdxo_sentinel_indices(const gdxo_index *idx) { return idx->sentinels; }
const uint64_t *gdxo_frequency_table(const gdxo_index *idx) { return idx->freq; }
uint64_t gdxo_num_border(const gdxo_index *idx) { return idx->n_border; }
const uint64_t *gdxo_border_rows(const gdxo_index *idx) { return idx->border_rows; }
const uint64_t *gdxo_border_pos(const gdxo_index *idx) { return idx->border_pos; }
uint64_t gdxo_num_samples(const gdxo_index *idx) { return idx->samples.len; }
uint64_t gdxo_sample(const gdxo_index *idx, uint64_t k) { return iarray_get(&idx->samples, k); }
const uint64_t *gdxo_blocks(const gdxo_index *idx, uint64_t *len) {
    *len = idx->rank->n_blocks_words;
    return idx->rank->blocks;
}
const uint16_t *gdxo_block_offsets(const gdxo_index *idx, uint64_t *len) {
    *len = idx->rank->n_block_offsets;
    return idx->rank->block_offsets;
}
uint64_t gdxo_superblock_offset(const gdxo_index *idx, uint64_t k) {
    return iarray_get(&idx->rank->superblock_offsets, k);
}
uint64_t gdxo_num_superblock_offsets(const gdxo_index *idx) {
    return idx->rank->superblock_offsets.len;
}
uint64_t gdxo_lookup_table_len(const gdxo_index *idx, uint32_t depth) {
    return depth < idx->n_tables ? idx->tables[depth].len / 2 : 0;
}
void gdxo_lookup_entry(const gdxo_index *idx, uint32_t depth, uint64_t i, uint64_t *start,
                       uint64_t *end) {
    *start = iarray_get(&idx->tables[depth], 2 * i);
    *end = iarray_get(&idx->tables[depth], 2 * i + 1);
}

/* ------------------------------------------------------------------------------------------------
 * batched search: BatchComputedCursors<.., 64>  (batch_computed_cursors.rs)
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    const uint8_t *q[BATCH_SIZE]; /* Buffers::queries        (:204) */
    uint64_t qlen[BATCH_SIZE];
    uint64_t start[BATCH_SIZE];   /* Buffers::intervals      (:203) */
    uint64_t end[BATCH_SIZE];
    uint32_t at_idx[BATCH_SIZE];  /* Buffers::query_at_idx   (:205) */
    uint8_t symbols[BATCH_SIZE];  /* Buffers::symbols        (:206) */
} batch_buffers;

static inline void batch_swap(batch_buffers *b, uint32_t i, uint32_t j) {
    const uint8_t *tq = b->q[i]; b->q[i] = b->q[j]; b->q[j] = tq;
    uint64_t t = b->qlen[i]; b->qlen[i] = b->qlen[j]; b->qlen[j] = t;
    t = b->start[i]; b->start[i] = b->start[j]; b->start[j] = t;
    t = b->end[i]; b->end[i] = b->end[j]; b->end[j] = t;
    uint32_t a = b->at_idx[i]; b->at_idx[i] = b->at_idx[j]; b->at_idx[j] = a;
}

/* batch_computed_cursors.rs:131-158 */
static void move_finished_queries_to_end(batch_buffers *b, uint64_t next_idx, uint32_t *unfinished) {
    uint32_t i = 0;
    while (i < *unfinished) {
        if (b->qlen[i] > next_idx && b->start[i] != b->end[i]) {
            ++i;
            continue;
        }
        uint32_t j = *unfinished - 1;
        batch_swap(b, i, j);
        *unfinished -= 1;
    }
}

/* batch_computed_cursors.rs:36-73 compute_next_batch for queries [q0, q0+bs) */
static int compute_batch(const gdxo_index *idx, const uint8_t *qbytes, const uint64_t *qoffsets,
                         uint64_t q0, uint32_t bs, uint64_t *starts, uint64_t *ends) {
    batch_buffers b;
    const uint64_t D = max_depth(idx);
    for (uint32_t i = 0; i < bs; ++i) { /* :41-47 */
        b.q[i] = qbytes + qoffsets[q0 + i];
        b.qlen[i] = qoffsets[q0 + i + 1] - qoffsets[q0 + i];
        b.at_idx[i] = i;
    }
    /* :75-96 batched_lookup_jumps */
    uint64_t depths[BATCH_SIZE], idxs[BATCH_SIZE];
    for (uint32_t i = 0; i < bs; ++i) {
        depths[i] = b.qlen[i] < D ? b.qlen[i] : D;
        int rc = compute_lookup_idx(idx, b.q[i] + (b.qlen[i] - depths[i]), depths[i], 1, &idxs[i]);
        if (rc) return rc;
    }
    for (uint32_t i = 0; i < bs; ++i) { /* lookup_table.rs:131-140 */
        int rc = lookup_idx(idx, depths[i], idxs[i], &b.start[i], &b.end[i]);
        if (rc) return rc;
    }
    uint64_t next_idx = D; /* :52 */
    uint32_t unfinished = bs;
    move_finished_queries_to_end(&b, next_idx, &unfinished);
    while (unfinished > 0) { /* :62-70 */
        /* :98-129 batched_lf_mappings */
        for (uint32_t i = 0; i < unfinished; ++i) {
            uint8_t s = idx->io_to_dense[b.q[i][b.qlen[i] - next_idx - 1]];
            if (s == 0) return GDXO_PANIC_INVALID_SYMBOL;
            b.symbols[i] = s;
        }
        gdxo_rank_batch(idx->rank, b.symbols, b.start, b.end, unfinished);
        for (uint32_t i = 0; i < unfinished; ++i) {
            b.start[i] += idx->count[b.symbols[i]];
            b.end[i] += idx->count[b.symbols[i]];
        }
        next_idx += 1;
        move_finished_queries_to_end(&b, next_idx, &unfinished);
    }
    /* :160-172 move_queries_back_to_initial_order == scatter by query_at_idx */
    for (uint32_t i = 0; i < bs; ++i) {
        starts[q0 + b.at_idx[i]] = b.start[i];
        ends[q0 + b.at_idx[i]] = b.end[i];
    }
    return GDXO_OK;
}

int gdxo_cursors_many(const gdxo_index *idx, const uint8_t *qbytes, const uint64_t *qoffsets,
                      uint64_t nq, uint64_t *starts, uint64_t *ends, uint64_t *panic_query) {
    for (uint64_t q0 = 0; q0 < nq; q0 += BATCH_SIZE) {
        uint32_t bs = (uint32_t)(nq - q0 < BATCH_SIZE ? nq - q0 : BATCH_SIZE);
        int rc = compute_batch(idx, qbytes, qoffsets, q0, bs, starts, ends);
        if (rc) {
            if (panic_query) *panic_query = q0;
            return rc;
        }
    }
    return GDXO_OK;
}

typedef struct {
    const gdxo_index *idx;
    const uint8_t *qbytes;
    const uint64_t *qoffsets;
    uint64_t *starts, *ends;
    int rc;
    uint64_t panic_query;
    pthread_mutex_t mu;
} many_ctx;

static void cursors_range(uint64_t b, uint64_t e, int tid, void *vctx) {
    (void)tid;
    many_ctx *c = (many_ctx *)vctx;
    uint64_t pq = 0;
    int rc = gdxo_cursors_many(c->idx, c->qbytes, c->qoffsets + b, e - b, c->starts + b, c->ends + b,
                               &pq);
    if (rc) {
        pthread_mutex_lock(&c->mu);
        if (!c->rc || b + pq < c->panic_query) {
            c->rc = rc;
            c->panic_query = b + pq;
        }
        pthread_mutex_unlock(&c->mu);
    }
}

static int cursors_many_mt(const gdxo_index *idx, const uint8_t *qbytes, const uint64_t *qoffsets,
                           uint64_t nq, int nthreads, uint64_t *starts, uint64_t *ends,
                           uint64_t *panic_query) {
    many_ctx c = {idx, qbytes, qoffsets, starts, ends, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    parallel_ranges(nthreads, nq, BATCH_SIZE, cursors_range, &c);
    if (c.rc && panic_query) *panic_query = c.panic_query;
    return c.rc;
}

int gdxo_count_many(const gdxo_index *idx, const uint8_t *qbytes, const uint64_t *qoffsets,
                    uint64_t nq, int nthreads, uint64_t *counts, uint64_t *panic_query) {
    uint64_t *ends = (uint64_t *)malloc(8 * (size_t)(nq ? nq : 1));
    if (!ends) return GDXO_ERR_ALLOC;
    int rc = cursors_many_mt(idx, qbytes, qoffsets, nq, nthreads, counts, ends, panic_query);
    if (!rc)
        for (uint64_t i = 0; i < nq; ++i) counts[i] = ends[i] - counts[i]; /* cursor.rs:61-63 */
    free(ends);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * locate
 * ---------------------------------------------------------------------------------------------- */

static uint64_t border_lookup(const gdxo_index *idx, uint64_t row) {
    uint64_t lo = 0, hi = idx->n_border;
    while (lo < hi) {
        uint64_t mid = (lo + hi) / 2;
        if (idx->border_rows[mid] < row) lo = mid + 1; else hi = mid;
    }
    return idx->border_pos[lo]; /* the reference indexes the HashMap and would panic if absent */
}

/* sampled_suffix_array.rs:110-138 recover_range, one row */
static uint64_t recover_row(const gdxo_index *idx, uint64_t i) {
    uint64_t steps = 0;
    while (i % idx->sampling_rate != 0) {
        uint8_t c = gdxo_rank_symbol_at(idx->rank, i);
        if (c == 0) return border_lookup(idx, i) + steps; /* :121-126 */
        i = lf_mapping_step(idx, c, i);
        steps += 1;
    }
    return iarray_get(&idx->samples, i / idx->sampling_rate) + steps;
}

/* lib.rs:187-197 + text_id_search_tree.rs:35-48 */
int gdxo_locate_interval(const gdxo_index *idx, uint64_t start, uint64_t end, gdxo_hit *out) {
    for (uint64_t i = start; i < end; ++i) {
        uint64_t pos = recover_row(idx, i);
        uint64_t id = tree_lookup_text_id(&idx->tree, pos);
        out[i - start].text_id = id;
        out[i - start].position = id == 0 ? pos : pos - idx->sentinels[id - 1] - 1;
    }
    return GDXO_OK;
}

typedef struct {
    const gdxo_index *idx;
    const uint64_t *starts, *ends, *hit_offsets;
    gdxo_hit *hits;
} locate_ctx;

static void locate_range(uint64_t b, uint64_t e, int tid, void *vctx) {
    (void)tid;
    locate_ctx *c = (locate_ctx *)vctx;
    for (uint64_t q = b; q < e; ++q)
        gdxo_locate_interval(c->idx, c->starts[q], c->ends[q], c->hits + c->hit_offsets[q]);
}

int gdxo_locate_many(const gdxo_index *idx, const uint8_t *qbytes, const uint64_t *qoffsets,
                     uint64_t nq, int nthreads, uint64_t *hit_offsets, gdxo_hit **hits,
                     uint64_t *panic_query) {
    *hits = NULL;
    uint64_t *starts = (uint64_t *)malloc(8 * (size_t)(nq ? nq : 1));
    uint64_t *ends = (uint64_t *)malloc(8 * (size_t)(nq ? nq : 1));
    if (!starts || !ends) { free(starts); free(ends); return GDXO_ERR_ALLOC; }
    int rc = cursors_many_mt(idx, qbytes, qoffsets, nq, nthreads, starts, ends, panic_query);
    if (!rc) {
        uint64_t total = 0;
        for (uint64_t q = 0; q < nq; ++q) {
            hit_offsets[q] = total;
            total += ends[q] - starts[q];
        }
        hit_offsets[nq] = total;
        gdxo_hit *h = (gdxo_hit *)malloc(sizeof(gdxo_hit) * (size_t)(total ? total : 1));
        if (!h) rc = GDXO_ERR_ALLOC;
        else {
            locate_ctx c = {idx, starts, ends, hit_offsets, h};
            parallel_ranges(nthreads, nq, BATCH_SIZE, locate_range, &c);
            *hits = h;
        }
    }
    free(starts);
    free(ends);
    return rc;
}

void gdxo_free_hits(gdxo_hit *hits) { free(hits); }

/* ------------------------------------------------------------------------------------------------
 * the other TextWithRankSupport variants: Condensed<Block512>, Flat<Block64>, Flat<Block512>
 * (lib.rs:104-113; block.rs:66-192; flat.rs; condensed.rs generic over B).  Restated by definition, in
 * the reference's own array layouts, so that the product's ingestion of those arrays can be tested.
 * ---------------------------------------------------------------------------------------------- */
struct gdxo_vrank {
    int variant;              /* GDXO_RANK_CONDENSED / GDXO_RANK_FLAT */
    uint32_t block_bits, W;   /* Block::NUM_BITS (block.rs:29), NUM_U64 (:33) */
    uint64_t text_len;
    uint32_t sigma, planes;
    uint64_t used_bits;       /* positions per block: NUM_BITS (condensed.rs:39-41) or NUM_BITS - 16 (flat.rs:43-45) */
    uint64_t superblock_size; /* 65536 (condensed.rs:34) or (65536 / used) * used (flat.rs:76-77) */
    uint64_t units;           /* blocks per position group: planes (condensed) or sigma (flat) */
    uint64_t *blocks;         /* [group][unit][W] */
    uint64_t n_block_words;
    uint16_t *block_offsets;  /* condensed only: [group][symbol] */
    uint64_t n_block_offsets;
    iarray superblock_offsets;
};

static inline void vblock_set_bit(uint64_t *blk, uint64_t idx) { blk[idx / 64] |= 1ull << (idx % 64); } /* block.rs:104-108 */
static inline uint32_t vblock_get_bit(const uint64_t *blk, uint64_t idx) { return (uint32_t)((blk[idx / 64] >> (idx % 64)) & 1u); }
/* block.rs:110-120 + BLOCK512_MASKS (:194-226): ones in bits [0, idx) */
static inline uint64_t vblock_count_ones_before(const uint64_t *blk, uint32_t W, uint64_t idx) {
    uint64_t sum = 0;
    for (uint32_t w = 0; w < W; ++w) {
        uint64_t mask = idx >= 64ull * (w + 1) ? ~0ull : (idx > 64ull * w ? ~(~0ull << (idx - 64ull * w)) : 0ull);
        sum += (uint64_t)__builtin_popcountll(blk[w] & mask);
    }
    return sum;
}

gdxo_vrank *gdxo_vrank_construct(const uint8_t *text, uint64_t n, uint32_t sigma, int storage, int variant,
                                 uint32_t block_bits) {
    if (sigma < 2 || (block_bits != 64 && block_bits != 512) || (variant != GDXO_RANK_CONDENSED && variant != GDXO_RANK_FLAT))
        return NULL;
    gdxo_vrank *r = (gdxo_vrank *)calloc(1, sizeof(*r));
    if (!r) return NULL;
    r->variant = variant;
    r->block_bits = block_bits;
    r->W = block_bits / 64;
    r->text_len = n;
    r->sigma = sigma;
    r->planes = ilog2_ceil_for_nonzero(sigma);
    r->used_bits = variant == GDXO_RANK_FLAT ? block_bits - 16 : block_bits;
    r->superblock_size = variant == GDXO_RANK_FLAT ? (65536 / r->used_bits) * r->used_bits : 65536;
    r->units = variant == GDXO_RANK_FLAT ? sigma : r->planes;
    const uint64_t len = n + 1; /* condensed.rs:69, flat.rs:73: rank(idx) is defined for idx <= text_len */
    const uint64_t groups = div_ceil(len, r->used_bits), nsb = div_ceil(len, r->superblock_size);
    r->n_block_words = groups * r->units * r->W;
    r->blocks = (uint64_t *)calloc(r->n_block_words ? r->n_block_words : 1, 8);
    if (variant == GDXO_RANK_CONDENSED) {
        r->n_block_offsets = groups * sigma;
        r->block_offsets = (uint16_t *)calloc(r->n_block_offsets ? r->n_block_offsets : 1, 2);
    }
    uint64_t *in_sb = (uint64_t *)calloc(sigma, 8), *total = (uint64_t *)calloc(sigma, 8);
    if (!r->blocks || iarray_alloc(&r->superblock_offsets, nsb * sigma, storage) || !in_sb || !total ||
        (variant == GDXO_RANK_CONDENSED && !r->block_offsets)) {
        free(in_sb);
        free(total);
        gdxo_vrank_free(r);
        return NULL;
    }
    /* position by position: offsets are "how many c since the superblock began" at every group start, the
     * superblock table "how many c before the superblock" (condensed.rs:104-116, flat.rs:107-118) */
    for (uint64_t g = 0; g < groups; ++g) {
        const uint64_t p0 = g * r->used_bits;
        if (p0 % r->superblock_size == 0) {
            for (uint32_t c = 0; c < sigma; ++c) {
                iarray_set(&r->superblock_offsets, (p0 / r->superblock_size) * sigma + c, total[c]);
                in_sb[c] = 0;
            }
        }
        for (uint32_t c = 0; c < sigma; ++c) {
            if (variant == GDXO_RANK_CONDENSED) r->block_offsets[g * sigma + c] = (uint16_t)in_sb[c];
            else r->blocks[(g * sigma + c) * r->W] = in_sb[c]; /* block.rs:122-124 integrate_block_offset */
        }
        for (uint64_t j = 0; j < r->used_bits && p0 + j < n; ++j) {
            const uint8_t sym = text[p0 + j];
            in_sb[sym] += 1;
            total[sym] += 1;
            if (variant == GDXO_RANK_CONDENSED) {
                for (uint32_t p = 0; p < r->planes; ++p)
                    if ((sym >> p) & 1u) vblock_set_bit(r->blocks + (g * r->planes + p) * r->W, j); /* condensed.rs:393-396 */
            } else {
                vblock_set_bit(r->blocks + (g * sigma + sym) * r->W, j + 16); /* flat.rs:291 */
            }
        }
    }
    free(in_sb);
    free(total);
    return r;
}

void gdxo_vrank_free(gdxo_vrank *r) {
    if (!r) return;
    free(r->blocks);
    free(r->block_offsets);
    free(r->superblock_offsets.p);
    free(r);
}

uint64_t gdxo_vrank_query(const gdxo_vrank *r, uint8_t symbol, uint64_t idx) {
    const uint64_t sb = iarray_get(&r->superblock_offsets, (idx / r->superblock_size) * r->sigma + symbol);
    const uint64_t g = idx / r->used_bits, in_block = idx % r->used_bits;
    if (r->variant == GDXO_RANK_FLAT) { /* flat.rs:221-246 */
        uint64_t blk[8];
        memcpy(blk, r->blocks + (g * r->sigma + symbol) * r->W, 8 * r->W);
        const uint64_t off = blk[0] & 0xffffull; /* block.rs:126-134 extract_block_offset_and_then_zeroize_it */
        blk[0] &= ~0xffffull;
        return sb + off + vblock_count_ones_before(blk, r->W, in_block + 16);
    }
    /* condensed.rs:291-341 */
    uint64_t acc[8];
    for (uint32_t w = 0; w < r->W; ++w) acc[w] = ~0ull;
    uint32_t s = symbol;
    for (uint32_t p = 0; p < r->planes; ++p, s >>= 1) {
        const uint64_t *blk = r->blocks + (g * r->planes + p) * r->W;
        for (uint32_t w = 0; w < r->W; ++w) acc[w] &= (s & 1u) ? blk[w] : ~blk[w];
    }
    return sb + r->block_offsets[g * r->sigma + symbol] + vblock_count_ones_before(acc, r->W, in_block);
}

uint8_t gdxo_vrank_symbol_at(const gdxo_vrank *r, uint64_t idx) {
    const uint64_t g = idx / r->used_bits, in_block = idx % r->used_bits;
    if (r->variant == GDXO_RANK_FLAT) { /* flat.rs:248-267 */
        for (uint32_t c = 0; c < r->sigma; ++c)
            if (vblock_get_bit(r->blocks + (g * r->sigma + c) * r->W, in_block + 16)) return (uint8_t)c;
        return 0xff; /* unreachable!() in the reference */
    }
    uint32_t sym = 0; /* condensed.rs:343-362 */
    for (uint32_t p = 0; p < r->planes; ++p) sym |= vblock_get_bit(r->blocks + (g * r->planes + p) * r->W, in_block) << p;
    return (uint8_t)sym;
}

const uint64_t *gdxo_vrank_blocks(const gdxo_vrank *r, uint64_t *len_words) {
    *len_words = r->n_block_words;
    return r->blocks;
}
const uint16_t *gdxo_vrank_block_offsets(const gdxo_vrank *r, uint64_t *len) {
    *len = r->n_block_offsets;
    return r->block_offsets;
}
uint64_t gdxo_vrank_num_superblock_offsets(const gdxo_vrank *r) { return r->superblock_offsets.len; }
uint64_t gdxo_vrank_superblock_offset(const gdxo_vrank *r, uint64_t k) { return iarray_get(&r->superblock_offsets, k); }
uint64_t gdxo_vrank_superblock_size(const gdxo_vrank *r) { return r->superblock_size; }

/* ------------------------------------------------------------------------------------------------
 * independent check of an index that was built from parts (gdx_oracle.h: gdxo_verify_against_text)
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    const gdxo_index *idx;
    const uint8_t *T;
    uint64_t n;
    uint64_t *pos;     /* boundary positions, ascending */
    uint64_t *row;     /* their rows, ~0 = not known */
    uint64_t nb;
    uint64_t *bitmap;  /* one bit per row */
    uint64_t violations, visited;
    pthread_mutex_t mu;
} verify_ctx;

/* order of two suffixes of T by plain comparison; the end of the text is smaller than any symbol
 * (the libsais convention, construction/mod.rs:88-103) */
static const uint8_t *g_sort_text;
static uint64_t g_sort_n;
static int suffix_cmp(const void *pa, const void *pb) {
    uint64_t a = *(const uint64_t *)pa, b = *(const uint64_t *)pb;
    if (a == b) return 0;
    while (a < g_sort_n && b < g_sort_n) {
        if (g_sort_text[a] != g_sort_text[b]) return g_sort_text[a] < g_sort_text[b] ? -1 : 1;
        ++a;
        ++b;
    }
    return a == g_sort_n ? -1 : 1;
}

/* row of the suffix starting at p by backward search of the sentinel-free symbols that follow
 * (lib.rs:248-271 over dense symbols); ~0 when they are not unique before the next sentinel */
static uint64_t row_by_search(const gdxo_index *idx, const uint8_t *T, uint64_t n, uint64_t p) {
    for (uint64_t L = 32; L <= (1u << 16); L *= 2) {
        uint64_t end = p + L < n ? p + L : n, stop = p;
        while (stop < end && T[stop] != 0) ++stop;
        if (stop == p) return ~0ull;
        uint64_t s = 0, e = idx->n;
        for (uint64_t r = stop; r-- > p && s != e;) extend_front_dense(idx, T[r], &s, &e);
        if (e - s == 1) return s;
        if (stop < end || end == n) return ~0ull; /* ran into a sentinel: longer patterns do not exist */
    }
    return ~0ull;
}

static void verify_search_range(uint64_t b, uint64_t e, int tid, void *vctx) {
    (void)tid;
    verify_ctx *c = (verify_ctx *)vctx;
    for (uint64_t k = b; k < e; ++k)
        if (c->row[k] == ~0ull && c->T[c->pos[k]] != 0) c->row[k] = row_by_search(c->idx, c->T, c->n, c->pos[k]);
}

static void verify_walk_range(uint64_t b, uint64_t e, int tid, void *vctx) {
    (void)tid;
    verify_ctx *c = (verify_ctx *)vctx;
    const gdxo_index *idx = c->idx;
    uint64_t bad = 0, seen = 0;
    for (uint64_t k = b; k < e; ++k) {
        if (c->row[k] == ~0ull) continue; /* no row of its own: covered by the walk from the boundary above */
        uint64_t p = c->pos[k], row = c->row[k];
        /* the next boundary below with a row of its own ends this walk */
        uint64_t kb = k;
        while (kb > 0 && c->row[kb - 1] == ~0ull) --kb;
        const int has_lower = kb > 0;
        const uint64_t lower_pos = has_lower ? c->pos[kb - 1] : 0, lower_row = has_lower ? c->row[kb - 1] : 0;
        for (;;) {
            if (row >= c->n) { ++bad; break; }
            uint64_t old = __atomic_fetch_or(&c->bitmap[row >> 6], 1ull << (row & 63), __ATOMIC_RELAXED);
            if (old & (1ull << (row & 63))) ++bad; /* a second suffix claims this row */
            ++seen;
            if (idx->samples.p && row % idx->sampling_rate == 0 &&
                iarray_get(&idx->samples, row / idx->sampling_rate) != p)
                ++bad; /* sampled_suffix_array.rs:27-54 */
            const uint8_t sym = gdxo_rank_symbol_at(idx->rank, row);
            const uint8_t want = p > 0 ? c->T[p - 1] : c->T[c->n - 1]; /* bwt.rs:96-105 */
            if (sym != want) { ++bad; break; }
            if (sym == 0) { /* a text starts here: bwt.rs:108-116 */
                if (idx->n_border && border_lookup(idx, row) != p) ++bad;
                break;
            }
            const uint64_t next = lf_mapping_step(idx, sym, row);
            if (has_lower && p - 1 == lower_pos) {
                if (next != lower_row) ++bad; /* the chain of walks must close on the searched rows */
                break;
            }
            row = next;
            --p;
        }
    }
    pthread_mutex_lock(&c->mu);
    c->violations += bad;
    c->visited += seen;
    pthread_mutex_unlock(&c->mu);
}

static int u64_cmp(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

int gdxo_verify_against_text(const gdxo_index *idx, const uint8_t *dense_text, uint64_t n, int nthreads,
                             uint64_t *violations, uint64_t *rows_visited) {
    if (violations) *violations = 0;
    if (rows_visited) *rows_visited = 0;
    if (n != idx->n || n == 0) {
        if (violations) *violations = 1;
        return GDXO_OK;
    }
    const uint64_t seg = n / 8192 > 4096 ? n / 8192 : 4096;
    const uint64_t ncand = n / seg + 1;
    verify_ctx c;
    memset(&c, 0, sizeof c);
    c.idx = idx;
    c.T = dense_text;
    c.n = n;
    c.pos = (uint64_t *)malloc(8 * (size_t)(ncand + idx->ntexts));
    c.row = (uint64_t *)malloc(8 * (size_t)(ncand + idx->ntexts));
    c.bitmap = (uint64_t *)calloc((size_t)(n / 64 + 1), 8);
    uint64_t *starts = (uint64_t *)malloc(8 * (size_t)idx->ntexts);
    if (!c.pos || !c.row || !c.bitmap || !starts) {
        free(c.pos); free(c.row); free(c.bitmap); free(starts);
        return GDXO_ERR_ALLOC;
    }
    pthread_mutex_init(&c.mu, NULL);
    /* rows of the sentinel suffixes, straight from the definition of the suffix order: the last sentinel is
     * followed by the end of the text and sorts first, the others sort by the text that follows them */
    uint64_t nstarts = 0;
    for (uint64_t t = 0; t + 1 < idx->ntexts; ++t) starts[nstarts++] = idx->sentinels[t] + 1;
    g_sort_text = dense_text;
    g_sort_n = n;
    qsort(starts, (size_t)nstarts, 8, suffix_cmp);
    uint64_t nb = 0;
    for (uint64_t k = 0; k < nstarts; ++k) {
        c.pos[nb] = starts[k] - 1;
        c.row[nb++] = 1 + k;
    }
    c.pos[nb] = n - 1;
    c.row[nb++] = 0;
    for (uint64_t k = 1; k * seg < n; ++k) {
        if (dense_text[k * seg] == 0) continue;
        c.pos[nb] = k * seg;
        c.row[nb++] = ~0ull;
    }
    /* sort boundaries by position, carrying the rows along (pack both into pairs) */
    uint64_t *pairs = (uint64_t *)malloc(16 * (size_t)nb);
    if (!pairs) { free(c.pos); free(c.row); free(c.bitmap); free(starts); return GDXO_ERR_ALLOC; }
    for (uint64_t k = 0; k < nb; ++k) { pairs[2 * k] = c.pos[k]; pairs[2 * k + 1] = c.row[k]; }
    qsort(pairs, (size_t)nb, 16, u64_cmp);
    for (uint64_t k = 0; k < nb; ++k) { c.pos[k] = pairs[2 * k]; c.row[k] = pairs[2 * k + 1]; }
    free(pairs);
    c.nb = nb;
    parallel_ranges(nthreads, nb, 1, verify_search_range, &c);
    /* the topmost boundary must own a row (it is the last sentinel, n - 1): every position is then covered */
    parallel_ranges(nthreads, nb, 1, verify_walk_range, &c);
    if (c.visited != n) c.violations += 1; /* some row was never reached */
    if (violations) *violations = c.violations;
    if (rows_visited) *rows_visited = c.visited;
    pthread_mutex_destroy(&c.mu);
    free(c.pos); free(c.row); free(c.bitmap); free(starts);
    return GDXO_OK;
}
