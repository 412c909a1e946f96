/* The crate's examples/basic_usage.rs against the C ABI (include/genedex_b200.h), plain C.
 *   gcc -Iinclude examples/basic_usage.c -Lgenedex_b200/csrc -lgenedex_b200 -Wl,-rpath,$PWD/genedex_b200/csrc
 * Needs a CUDA device to run. */
#include <stdio.h>
#include <string.h>

#include "genedex_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        gdx_status st_ = (call);                                                 \
        if (st_ != GDX_OK) {                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", #call, (int)st_, gdx_last_error_message()); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(void) {
    /* alphabet::ascii_dna_with_n(): A C G T searchable, N not (src/alphabet.rs:255-258) */
    gdx_alphabet alphabet;
    memset(&alphabet, 0, sizeof alphabet);
    const char *groups[5] = {"Aa", "Cc", "Gg", "Tt", "Nn"};
    for (int g = 0; g < 5; ++g)
        for (const char *c = groups[g]; *c; ++c) alphabet.io_to_dense[(unsigned char)*c] = (uint8_t)(g + 1);
    alphabet.num_dense_symbols = 6;
    alphabet.num_searchable_dense_symbols = 4;

    /* texts = [b"aACGT", b"acGtn"], FmIndexConfig::<i32>::new().suffix_array_sampling_rate(2) */
    const uint8_t texts[] = "aACGTacGtn";
    const uint64_t text_offsets[3] = {0, 5, 10};
    gdx_config config = {GDX_I32, 2, 0, 1, GDX_CONSTRUCT_AUTO, -1, 0, 0};
    gdx_index *index = NULL;
    CHECK(gdx_index_build(texts, text_offsets, 2, &alphabet, &config, &index));

    /* index.count_many / locate_many(["AC", "CG", "GT", "GTN"]) */
    const uint8_t qbytes[] = "ACCGGTGTN";
    const uint64_t qoffsets[5] = {0, 2, 4, 6, 9};
    gdx_queries queries = {qbytes, qoffsets, 0, 4, GDX_QUERIES_IO_BYTES, 0};
    uint64_t counts[4], hit_offsets[5], num_hits = 0;
    gdx_hit *hits = NULL;
    CHECK(gdx_count_many(index, &queries, counts));
    CHECK(gdx_locate_many(index, &queries, hit_offsets, &hits, &num_hits));
    for (int q = 0; q < 4; ++q) {
        printf("query %d: count %llu\n", q, (unsigned long long)counts[q]);
        for (uint64_t k = hit_offsets[q]; k < hit_offsets[q + 1]; ++k)
            printf("  found in text %llu at position %llu\n", (unsigned long long)hits[k].text_id,
                   (unsigned long long)hits[k].position);
    }
    gdx_free_hits(index, hits);
    gdx_index_destroy(index);
    return !(counts[0] == 2 && counts[1] == 2 && counts[2] == 2 && counts[3] == 1 && num_hits == 7);
}
