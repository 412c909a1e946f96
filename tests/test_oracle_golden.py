"""Pins the CPU oracle against every known-answer test the reference holds for the search path.

Each test names the reference test it replays (file:line under /root/reference).  These run on CPU.
"""
import numpy as np
import pytest

from oracle import oracle as O

RANDOM_LENGTH_512 = [  # tests/text_with_rank_support.rs:94-116, alphabet_size 27
    2, 2, 25, 0, 15, 19, 23, 7, 18, 13, 1, 20, 16, 14, 19, 15, 3, 4, 13, 17, 12, 22, 21, 8, 5,
    11, 13, 25, 2, 21, 16, 22, 23, 19, 3, 13, 23, 19, 18, 20, 13, 23, 2, 2, 17, 6, 9, 5, 19,
    26, 4, 18, 20, 17, 18, 1, 20, 26, 13, 3, 15, 17, 7, 2, 26, 12, 11, 25, 18, 25, 17, 24, 8,
    14, 15, 3, 14, 9, 11, 26, 12, 18, 21, 8, 7, 22, 7, 9, 10, 2, 14, 9, 4, 21, 13, 4, 4, 7, 0,
    24, 4, 4, 7, 10, 2, 3, 11, 12, 16, 9, 5, 6, 10, 25, 21, 6, 16, 3, 23, 5, 4, 15, 14, 1, 12,
    15, 3, 24, 2, 25, 9, 1, 18, 21, 15, 13, 1, 22, 6, 10, 15, 14, 16, 16, 13, 24, 5, 2, 21, 16,
    6, 19, 22, 6, 24, 23, 26, 26, 19, 13, 26, 0, 23, 6, 24, 13, 2, 20, 10, 15, 13, 22, 25, 3,
    11, 14, 5, 0, 13, 15, 12, 22, 7, 14, 14, 23, 20, 14, 21, 12, 10, 15, 19, 23, 2, 16, 14, 13,
    8, 0, 18, 3, 23, 10, 6, 2, 19, 9, 11, 19, 9, 22, 1, 11, 5, 12, 21, 19, 26, 18, 15, 25, 12,
    18, 10, 22, 13, 5, 22, 23, 5, 19, 6, 19, 19, 7, 8, 2, 26, 18, 1, 21, 20, 15, 4, 24, 16, 5,
    5, 4, 15, 3, 23, 21, 23, 3, 6, 15, 23, 6, 7, 1, 25, 0, 22, 10, 3, 10, 7, 15, 26, 1, 22, 9,
    11, 22, 1, 8, 19, 10, 25, 3, 2, 14, 19, 23, 22, 15, 11, 5, 0, 21, 5, 6, 25, 0, 21, 26, 21,
    5, 11, 8, 9, 10, 8, 20, 5, 0, 2, 15, 12, 24, 6, 6, 16, 16, 21, 4, 5, 12, 4, 12, 5, 23, 22,
    25, 12, 12, 5, 7, 16, 13, 3, 19, 26, 17, 15, 0, 10, 4, 3, 3, 19, 11, 5, 20, 24, 1, 8, 6,
    26, 25, 12, 15, 25, 0, 7, 25, 1, 12, 2, 26, 25, 2, 2, 4, 18, 10, 0, 9, 21, 10, 22, 1, 0,
    22, 11, 7, 4, 4, 9, 14, 10, 19, 22, 23, 18, 18, 9, 5, 25, 3, 9, 10, 13, 3, 16, 12, 5, 7,
    14, 17, 24, 21, 14, 0, 13, 26, 21, 26, 25, 4, 26, 2, 23, 14, 10, 26, 3, 26, 21, 2, 24, 19,
    17, 11, 26, 9, 11, 11, 17, 14, 9, 2, 21, 8, 26, 22, 7, 11, 19, 7, 17, 17, 16, 11, 17, 22,
    20, 4, 14, 6, 17, 5, 18, 8, 17, 13, 4, 3, 18, 7, 17, 26, 9, 14, 22, 13, 23, 25, 12, 3, 7,
    8, 17, 12, 14, 10, 8, 17, 26, 22, 12, 20, 13, 25, 23, 9, 20, 7, 6, 11, 15, 26, 15, 1, 21,
    12, 0, 9, 0, 9, 19, 10, 19, 26, 26, 21, 7, 18, 6, 14,
]


def _dna_index(storage):  # tests/fmindex.rs:7-13 create_index
    return O.OracleIndex.build([b"cccaaagggttt"], O.ALPHABETS["ascii_dna"](), storage=storage,
                               sampling_rate=3, lookup_depth=0)


@pytest.mark.parametrize("storage", ["i32", "u32"])
def test_basic_search(storage):  # tests/fmindex.rs:20-43
    assert set(_dna_index(storage).locate(b"gg")) == {(0, 6), (0, 7)}


@pytest.mark.parametrize("storage", ["i32", "u32"])
def test_text_front_search(storage):  # tests/fmindex.rs:45-72
    assert set(_dna_index(storage).locate(b"c")) == {(0, 0), (0, 1), (0, 2)}


@pytest.mark.parametrize("storage", ["i32", "u32"])
def test_search_no_wrapping(storage):  # tests/fmindex.rs:74-80
    assert _dna_index(storage).locate(b"ta") == []


def test_search_multitext():  # tests/fmindex.rs:82-126
    idx = O.OracleIndex.build([b"cccaaagggttt", b"acgtacgtacgt"], O.ALPHABETS["ascii_dna"](),
                              storage="u32", sampling_rate=3, lookup_depth=4)
    assert set(idx.locate(b"gg")) == {(0, 6), (0, 7)}
    assert set(idx.locate(b"gt")) == {(0, 8), (1, 2), (1, 6), (1, 10)}
    assert [set(h) for h in idx.locate_many([b"gg", b"gt"])] == [{(0, 6), (0, 7)},
                                                                  {(0, 8), (1, 2), (1, 6), (1, 10)}]


def test_u8_alphabet():  # tests/fmindex.rs:128-154
    texts = [bytes([0, 4, 3, 2, 1, 5, 8, 6, 7, 8]), bytes([5, 7, 3, 4, 2, 1, 5, 8]), b""]
    idx = O.OracleIndex.build(texts, O.u8_until(8), storage="u32", sampling_rate=3, lookup_depth=4)
    assert set(idx.locate(bytes([1, 5, 8]))) == {(0, 4), (1, 5)}
    assert idx.num_texts == 3


def test_basic_usage_example():  # examples/basic_usage.rs:8-16, lib.rs:15-32, README.md:33
    idx = O.OracleIndex.build([b"aACGT", b"acGtn"], O.ALPHABETS["ascii_dna_with_n"](), storage="i32",
                              sampling_rate=2, lookup_depth=0)
    assert idx.count(b"GT") == 2
    assert list(idx.count_many([b"AC", b"CG", b"GT", b"GTN"])) == [2, 2, 2, 1]


def test_cursor_example():  # examples/cursor.rs:6-24
    idx = O.OracleIndex.build([b"AaACGT", b"AacGtn", b"GTGTGT"], O.ALPHABETS["ascii_dna_with_n"](),
                              storage="i32")
    iv = idx.cursor_for_query(b"GT")
    assert iv[1] - iv[0] == 5
    iv = idx.extend_query_front(iv, ord("C"))
    assert iv[1] - iv[0] == 2
    assert set(idx.locate_interval(*iv)) == {(0, 3), (1, 2)}


def test_concat_text():  # src/construction/mod.rs:374-398
    idx = O.OracleIndex.build([b"cccaaagggttt", b"acgtacgtacgt"], O.ALPHABETS["ascii_dna"](), storage="i32")
    assert idx.dense_text().tolist() == [2, 2, 2, 1, 1, 1, 3, 3, 3, 4, 4, 4, 0,
                                         1, 2, 3, 4, 1, 2, 3, 4, 1, 2, 3, 4, 0]
    assert idx.sentinel_indices().tolist() == [12, 25]
    freq = idx.frequency_table()
    assert freq[:5].tolist() == [2, 6, 6, 6, 6] and freq[5:].sum() == 0
    # construction/mod.rs:318-336: exclusive prefix sums, sigma + 1 entries
    assert idx.count_array().tolist() == [0, 2, 8, 14, 20, 26]


def test_text_id_search_tree():  # src/text_id_search_tree.rs:160-172
    sent = [10, 21, 32, 50, 68, 140, 141]
    for pos, want in [(5, 0), (21, 1), (0, 0), (140, 5), (141, 6), (33, 3), (67, 4)]:
        assert O.tree_lookup(sent, pos) == want
    # lower_bound equivalence for every position
    for pos in range(0, 142):
        assert O.tree_lookup(sent, pos) == int(np.searchsorted(sent, pos, side="left"))


def _check_rank_against_naive(text, sigma, storage="i32"):
    # tests/text_with_rank_support.rs:46-67 test_against_naive
    r = O.OracleRank(text, sigma, storage)
    text = np.asarray(text, dtype=np.uint8)
    for i, s in enumerate(text):
        assert r.symbol_at(i) == s
    for symbol in range(sigma):
        occ = np.concatenate([[0], np.cumsum(text == symbol)]) if text.size else np.zeros(1, int)
        for idx in range(text.size + 1):
            assert r.rank(symbol, idx) == occ[idx], (symbol, idx)
    return r


def test_rank_empty():  # tests/text_with_rank_support.rs:77-83
    _check_rank_against_naive([], 2)


def test_rank_superblock_size_text_of_single_character():  # tests/text_with_rank_support.rs:85-89
    text = np.zeros(65536, dtype=np.uint8)
    r = O.OracleRank(text, 2, "u32")
    for idx in list(range(0, 300)) + list(range(65536 - 300, 65537)) + list(range(1000, 65536, 997)):
        assert r.rank(0, idx) == idx
        assert r.rank(1, idx) == 0


def test_rank_random_length_512():  # tests/text_with_rank_support.rs:91-119
    assert len(RANDOM_LENGTH_512) == 512
    _check_rank_against_naive(RANDOM_LENGTH_512, 27)
    _check_rank_against_naive(RANDOM_LENGTH_512, 27, "i64")


def test_rank_example():  # examples/text_with_rank_support.rs:8-22 (values hold for every variant)
    r = O.OracleRank([0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3], 4)
    assert r.symbol_at(4) == 1 and r.rank(1, 4) == 1


def test_rank_layout_sizes():  # condensed.rs:69-74: n+1 positions, b planes, sigma offsets
    n, sigma = 1000, 6
    idx = O.OracleIndex.build([b"A" * (n - 1)], O.ALPHABETS["ascii_dna_with_n"](), storage="i32")
    nblocks = (n + 1 + 63) // 64
    assert idx.blocks().size == nblocks * 3
    assert idx.block_offsets().size == nblocks * sigma
    assert idx.superblock_offsets().size == sigma


def _copied_and_recovered_equal(texts, sampling_rate):
    # src/sampled_suffix_array.rs:146-166 copied_and_recovered_array_must_equal
    alph = O.ALPHABETS["ascii_dna_with_n"]()
    n = sum(len(t) for t in texts)
    sampled = O.OracleIndex.build(texts, alph, "i32", sampling_rate=sampling_rate, lookup_depth=4)
    full = O.OracleIndex.build(texts, alph, "i32", sampling_rate=1, lookup_depth=4)
    # recover_range yields concatenated-text positions; compare through (text_id, position)
    assert sampled.locate_interval(0, n) == full.locate_interval(0, n)
    # and the full-rate one is the plain suffix array
    sa = full.suffix_array()
    got = full.locate_interval(0, full.text_len)
    sent = full.sentinel_indices()
    for row, (tid, pos) in enumerate(got):
        base = 0 if tid == 0 else int(sent[tid - 1]) + 1
        assert base + pos == sa[row]


def test_walking_over_text_borders():  # src/sampled_suffix_array.rs:168-179
    _copied_and_recovered_equal([bytes([65]), b"", bytes([78, 84, 78, 78, 84, 78, 78, 84, 78])], 5)


def test_regression_edge_inputs():
    # tests/fmindex.proptest-regressions:9 and proptest-regressions/sampled_suffix_array.txt:8
    alph = O.ALPHABETS["ascii_dna"]()
    idx = O.OracleIndex.build([b""], alph, "i32", sampling_rate=1, lookup_depth=0)
    assert idx.text_len == 1
    assert idx.locate(b"") == [(0, 0)]
    assert idx.count(b"A") == 0
    _copied_and_recovered_equal([b""], 1)
    # proptest-regressions/text_with_rank_support/mod.txt:7 (text=[], alphabet_size=1): the
    # reference asserts alphabet_size >= 2 (condensed.rs:64)
    with pytest.raises(O.OraclePanic):
        O.OracleRank([], 1)
    # proptest-regressions/construction/bwt.txt:7-8: BWT of [2,0] and of 128 zeros
    a = O.u8_until(3)  # io symbol x -> dense x+1 ; texts end with the sentinel themselves
    idx = O.OracleIndex.build([bytes([1])], a, "i32", sampling_rate=1)  # dense text [2,0]
    assert idx.dense_text().tolist() == [2, 0]
    assert idx.suffix_array().tolist() == [1, 0]
    assert idx.bwt().tolist() == [2, 0]
    rows, pos = idx.border()
    assert rows.tolist() == [1] and pos.tolist() == [0]
    idx = O.OracleIndex.build([b""] * 128, a, "i32", sampling_rate=1)  # dense text = 128 zeros
    assert idx.suffix_array().tolist() == list(range(127, -1, -1))
    assert idx.bwt().tolist() == [0] * 128
    rows, pos = idx.border()
    assert rows.tolist() == list(range(128)) and pos.tolist() == list(range(127, -1, -1))


def test_alphabet_presets():  # src/alphabet.rs:365-428 construct_alphabets
    want = {"ascii_dna": (5, 4), "ascii_dna_with_n": (6, 4), "ascii_dna_iupac": (16, 15),
            "ascii_dna_iupac_as_dna_with_n": (6, 4), "ascii_amino_acid": (23, 22),
            "ascii_amino_acid_iupac": (28, 27), "ascii_printable": (96, 95)}
    for name, (sigma, ns) in want.items():
        a = O.ALPHABETS[name]()
        assert (a.sigma, a.num_searchable) == (sigma, ns), name
    for m in range(1, 255):
        a = O.u8_until(m)
        assert (a.sigma, a.num_searchable) == (m + 2, m + 1)


def test_invalid_symbol_panics_lazily():
    # alphabet.rs:195-198 panic; batch_computed_cursors.rs:84-87,106-113 translate lazily
    idx = O.OracleIndex.build([b"ACGTACGT"], O.ALPHABETS["ascii_dna"](), "i32")
    with pytest.raises(O.OraclePanic) as e:
        idx.count(b"AXGT"[1:2])
    assert e.value.code == O.PANIC_INVALID_SYMBOL
    # "TTX..": the search dies at "TT" (absent) before it would reach X when read right-to-left
    assert idx.count(b"XTT") == 0
    assert list(idx.count_many([b"XTT", b"ACG"])) == [0, 2]
    with pytest.raises(O.OraclePanic):
        idx.count_many([b"ACG", b"XGT"])


# ---- the other three rank variants (lib.rs:104-113), same tests as the reference runs for all four:
#      tests/text_with_rank_support.rs:69-119 test_different_block_sizes_against_naive ------------------------
VARIANTS = [("condensed", 64, "i32"), ("condensed", 512, "u32"), ("flat", 64, "i64"), ("flat", 512, "i32")]


def _check_variant_against_naive(text, sigma, variant, block_bits, storage):
    r = O.OracleVariantRank(text, sigma, storage, variant, block_bits)
    for i, s in enumerate(text):
        assert r.symbol_at(i) == s
    for symbol in range(sigma):
        count = 0
        for idx in range(len(text) + 1):
            assert r.rank(symbol, idx) == count, (variant, block_bits, symbol, idx)
            if idx < len(text) and text[idx] == symbol:
                count += 1


@pytest.mark.parametrize("variant,block_bits,storage", VARIANTS)
def test_rank_variants_reference_kats(variant, block_bits, storage):
    _check_variant_against_naive([], 2, variant, block_bits, storage)                 # :77-83 empty
    _check_variant_against_naive(RANDOM_LENGTH_512, 27, variant, block_bits, storage)  # :91-119 random_length_512
    r = O.OracleVariantRank([0] * 65536, 2, storage, variant, block_bits)              # :85-89 superblock_size_text...
    for idx in list(range(0, 65537, 97)) + [65535, 65536]:
        assert r.rank(0, idx) == idx and r.rank(1, idx) == 0


def test_rank_variant_layouts():
    """array shapes of the reference's own constructors (condensed.rs:69-77, flat.rs:73-83; block.rs:3,29-33)"""
    rng = np.random.default_rng(1)
    idx = O.OracleIndex.build([bytes(rng.choice(list(b"ACGTN"), 70_000).astype(np.uint8))],
                              O.ALPHABETS["ascii_dna_with_n"](), "u32", 4, 0)
    text = list(idx.bwt())
    c64 = O.OracleVariantRank(text, 6, "u32", "condensed", 64)
    assert np.array_equal(c64.blocks(), idx.blocks())  # == the pinned Block64 restatement inside the index
    assert np.array_equal(c64.block_offsets(), idx.block_offsets())
    assert np.array_equal(c64.superblock_offsets(), idx.superblock_offsets())
    n1 = len(text) + 1
    c512 = O.OracleVariantRank(text, 6, "u32", "condensed", 512)
    assert c512.blocks().size == -(-n1 // 512) * 3 * 8 and c512.block_offsets().size == -(-n1 // 512) * 6
    f64 = O.OracleVariantRank(text, 6, "u32", "flat", 64)
    assert f64.superblock_size() == (65536 // 48) * 48 and f64.blocks().size == -(-n1 // 48) * 6
    assert f64.superblock_offsets().size == -(-n1 // f64.superblock_size()) * 6 and f64.block_offsets().size == 0
    f512 = O.OracleVariantRank(text, 6, "u32", "flat", 512)
    assert f512.superblock_size() == (65536 // 496) * 496 and f512.blocks().size == -(-n1 // 496) * 6 * 8
