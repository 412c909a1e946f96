// The reference's known-answer tests written against the C++ mirror (include/genedex_b200.hpp);
// reads like tests/fmindex.rs and the examples of the crate.  Needs a GPU to run.
#include <algorithm>
#include <cstdio>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "genedex_b200.hpp"

#include <tuple>

#define REQUIRE(cond)                                                        \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
            return 1;                                                        \
        }                                                                    \
    } while (0)

using namespace gdx;
using HitSet = std::set<Hit>;

static HitSet set_of(const std::vector<Hit> &v) { return HitSet(v.begin(), v.end()); }

template <class I>
static int single_text() {  // tests/fmindex.rs:7-80
    std::vector<std::string> texts{"cccaaagggttt"};
    FmIndex index = FmIndexConfig<I>().lookup_table_depth(0).suffix_array_sampling_rate(3).construct_index(
        texts, alphabet::ascii_dna());
    REQUIRE(set_of(index.locate("gg")) == (HitSet{{0, 6}, {0, 7}}));
    REQUIRE(set_of(index.locate("c")) == (HitSet{{0, 0}, {0, 1}, {0, 2}}));
    REQUIRE(index.locate("ta").empty());
    return 0;
}

int main() {
    if (single_text<I32>() || single_text<U32>() || single_text<I64>()) return 1;
    {  // tests/fmindex.rs:82-126
        std::vector<std::string> texts{"cccaaagggttt", "acgtacgtacgt"};
        FmIndex index = FmIndexConfig<U32>().lookup_table_depth(4).suffix_array_sampling_rate(3).construct_index(
            texts, alphabet::ascii_dna());
        REQUIRE(set_of(index.locate("gg")) == (HitSet{{0, 6}, {0, 7}}));
        REQUIRE(set_of(index.locate("gt")) == (HitSet{{0, 8}, {1, 2}, {1, 6}, {1, 10}}));
        std::vector<std::string> qs{"gg", "gt", "tc"};
        auto many = index.locate_many(qs);
        for (size_t i = 0; i < many.size(); ++i) {
            std::printf("query %zu:", i);
            for (auto h : many[i]) std::printf(" (%zu,%zu)", h.text_id, h.position);
            std::printf("\n");
        }
        REQUIRE(many.size() == 3);
        REQUIRE(set_of(many[0]) == (HitSet{{0, 6}, {0, 7}}));
        REQUIRE(set_of(many[1]) == (HitSet{{0, 8}, {1, 2}, {1, 6}, {1, 10}}));
        REQUIRE(many[2].empty());
        REQUIRE((index.count_many(qs) == std::vector<size_t>{2, 4, 0}));
        REQUIRE(index.num_texts() == 2 && index.total_text_len() == 26);
    }
    {  // examples/basic_usage.rs:8-16 and examples/cursor.rs:6-24
        std::vector<std::string> texts{"aACGT", "acGtn"};
        FmIndex index = FmIndexConfig<I32>().suffix_array_sampling_rate(2).construct_on_device().construct_index(
            texts, alphabet::ascii_dna_with_n());
        REQUIRE(index.count("GT") == 2);
        std::vector<std::string> t3{"AaACGT", "AacGtn", "GTGTGT"};
        FmIndex idx3 = FmIndexConfig<I32>().construct_index(t3, alphabet::ascii_dna_with_n());
        Cursor cursor = idx3.cursor_for_query("GT");
        REQUIRE(cursor.count() == 5);
        cursor.extend_query_front('C');
        REQUIRE(cursor.count() == 2);
        REQUIRE(set_of(cursor.locate()) == (HitSet{{0, 3}, {1, 2}}));
        std::vector<std::string> qs{"GT", "CGT"};
        auto cs = idx3.cursors_for_many_queries(qs);
        REQUIRE(cs[0].count() == 5 && cs[1].interval() == cursor.interval());
    }
    {  // alphabet.rs:195-198: invalid symbols panic -> exception with the query index
        std::vector<std::string> texts{"ACGTACGT"};
        FmIndex index = FmIndexConfig<I32>().construct_index(texts, alphabet::ascii_dna());
        std::vector<std::string> qs{"ACG", "AXG"};
        bool thrown = false;
        try {
            index.count_many(qs);
        } catch (const InvalidSymbol &e) {
            thrown = e.query == 1;
        }
        REQUIRE(thrown);
    }
    {  // one batch over two replicas (both on the current device here; other devices work the same way)
        std::vector<std::string> texts{"cccaaagggttt", "acgtacgtacgt"};
        FmIndex index = FmIndexConfig<U32>().suffix_array_sampling_rate(3).construct_index(texts, alphabet::ascii_dna());
        ReplicaSet replicas(index, {index.info().device});
        REQUIRE(replicas.size() == 2);
        std::vector<std::string> qs{"gg", "gt", "tc", "acgt", "cc"};
        REQUIRE((replicas.count_many(qs) == index.count_many(qs)));
        auto a = replicas.locate_many(qs), b = index.locate_many(qs);
        REQUIRE(a.size() == b.size());
        for (size_t i = 0; i < a.size(); ++i) REQUIRE(set_of(a[i]) == set_of(b[i]) && a[i].size() == b[i].size());
    }
    std::printf("cpp api ok\n");
    return 0;
}
