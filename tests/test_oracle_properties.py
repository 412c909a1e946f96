"""Replays the reference's property tests against the CPU oracle with hypothesis (CPU only).

tests/fmindex.rs:264-315, tests/text_with_rank_support.rs:121-135,
src/text_with_rank_support/mod.rs:194-246, src/sampled_suffix_array.rs:181-195.
"""
import random

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import oracle as O

SETTINGS = dict(deadline=None, suppress_health_check=list(HealthCheck))

dna_text = st.lists(st.sampled_from(b"ACGT"), min_size=0, max_size=1499).map(bytes)
dna_n_text = st.lists(st.sampled_from(b"ACGTN"), min_size=0, max_size=1499).map(bytes)


def _sample_existing(rng, texts, max_extent=200, k=20):  # tests/fmindex.rs:156-188 QuerySampler
    out = []
    for _ in range(k):
        tid = rng.randrange(len(texts))
        t = texts[tid]
        if not t:
            break
        pos = rng.randrange(len(t))
        extent = rng.randrange(min(max_extent, len(t) - pos + 1))
        out.append(((tid, pos), t[pos:pos + extent]))
    return out


@settings(max_examples=40, **SETTINGS)
@given(texts=st.lists(dna_text, min_size=1, max_size=4), s=st.integers(1, 64), depth=st.integers(0, 5),
       seed=st.integers(0, 2 ** 32))
def test_correctness_random_texts(texts, s, depth, seed):  # tests/fmindex.rs:264-315
    rng = random.Random(seed)
    existing = _sample_existing(rng, texts)
    rand_q = [bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(20))) for _ in range(100)]
    naive = [O.naive_search(texts, q) for q in rand_q]
    for storage, alph in (("i32", "ascii_dna"), ("u32", "ascii_dna_with_n"),
                          ("i64", "ascii_dna_iupac_as_dna_with_n")):
        idx = O.OracleIndex.build(texts, O.ALPHABETS[alph](), storage, sampling_rate=s, lookup_depth=depth)
        many = idx.locate_many([q for _, q in existing])  # tests/fmindex.rs:229-262 run_queries
        assert len(many) == len(existing)
        for (hit, q), m in zip(existing, many):
            assert hit in set(idx.locate(q))
            assert hit in set(m)
        many = idx.locate_many(rand_q)
        counts = idx.count_many(rand_q)
        for q, want, m, c in zip(rand_q, naive, many, counts):
            single = idx.locate(q)
            assert set(single) == want
            assert m == single  # same SA-row order, not only the same set
            assert c == len(want) == idx.count(q)


@settings(max_examples=60, **SETTINGS)
@given(data=st.data())
def test_rank_correctness_random_texts(data):  # tests/text_with_rank_support.rs:121-135
    max_symbol = data.draw(st.integers(1, 255))
    text = data.draw(st.lists(st.integers(0, max_symbol), min_size=0, max_size=999))
    sigma = max_symbol + 1
    storage = data.draw(st.sampled_from(["i32", "u32", "i64"]))
    r = O.OracleRank(text, sigma, storage)
    t = np.asarray(text, dtype=np.int64)
    for i, sym in enumerate(text):
        assert r.symbol_at(i) == sym
    for symbol in sorted(set(text) | {0, max_symbol}):
        occ = np.concatenate([[0], np.cumsum(t == symbol)])
        for idx in range(len(text) + 1):
            assert r.rank(symbol, idx) == occ[idx]


@settings(max_examples=60, **SETTINGS)
@given(data=st.data())
def test_replace_many_intervals_same_as_rank(data):  # text_with_rank_support/mod.rs:194-225,241-245
    max_char = data.draw(st.integers(2, 31))
    text = data.draw(st.lists(st.integers(0, max_char), min_size=0, max_size=999))
    r = O.OracleRank(text, max_char + 1, "u32")
    rng = random.Random(data.draw(st.integers(0, 2 ** 32)))
    for _ in range(20):
        n = rng.randrange(1, 65)
        starts = [rng.randrange(len(text) + 1) for _ in range(n)]  # start > end is allowed here
        ends = [rng.randrange(len(text) + 1) for _ in range(n)]
        symbols = [rng.randrange(max_char + 1) for _ in range(n)]
        a, b = r.rank_batch(symbols, starts, ends)
        assert a.tolist() == [r.rank(c, i) for c, i in zip(symbols, starts)]
        assert b.tolist() == [r.rank(c, i) for c, i in zip(symbols, ends)]


@settings(max_examples=60, **SETTINGS)
@given(texts=st.lists(dna_n_text, min_size=1, max_size=4), s=st.integers(1, 8))
def test_sampled_suffix_array_recovery(texts, s):  # src/sampled_suffix_array.rs:181-195
    alph = O.ALPHABETS["ascii_dna_with_n"]()
    n = sum(len(t) for t in texts)
    sampled = O.OracleIndex.build(texts, alph, "i32", sampling_rate=s, lookup_depth=4)
    full = O.OracleIndex.build(texts, alph, "i32", sampling_rate=1, lookup_depth=4)
    assert sampled.locate_interval(0, n) == full.locate_interval(0, n)


@settings(max_examples=80, **SETTINGS)
@given(data=st.data())
def test_suffix_array_is_the_unique_one(data):
    # construction/mod.rs:88-103 calls libsais; the SA is unique, so compare with a plain sort.
    # Low-entropy and periodic texts stress the prefix-doubling rounds.
    kind = data.draw(st.sampled_from(["random", "runs", "periodic"]))
    if kind == "random":
        texts = data.draw(st.lists(st.lists(st.sampled_from(b"ACGTN"), max_size=300).map(bytes),
                                   min_size=1, max_size=5))
    elif kind == "runs":
        parts = data.draw(st.lists(st.tuples(st.sampled_from(b"ACGTN"), st.integers(1, 120)),
                                   min_size=0, max_size=12))
        texts = [b"".join(bytes([c]) * k for c, k in parts), b"", b"NNNNNNNN"]
    else:
        unit = data.draw(st.lists(st.sampled_from(b"AC"), min_size=1, max_size=5).map(bytes))
        texts = [unit * data.draw(st.integers(1, 150)), unit * 3]
    idx = O.OracleIndex.build(texts, O.ALPHABETS["ascii_dna_with_n"](), "i32", sampling_rate=1)
    dense = idx.dense_text().tolist()
    assert idx.suffix_array().tolist() == O.naive_suffix_array(dense)
    # bwt.rs:96-105
    sa = idx.suffix_array()
    n = len(dense)
    assert idx.bwt().tolist() == [dense[(p - 1) % n] for p in sa.tolist()]


@settings(max_examples=40, **SETTINGS)
@given(texts=st.lists(dna_text, min_size=1, max_size=3), depth=st.integers(0, 5))
def test_lookup_tables_hold_the_search_result(texts, depth):  # lookup_table.rs:163-258
    with_tables = O.OracleIndex.build(texts, O.ALPHABETS["ascii_dna"](), "u32", lookup_depth=depth)
    plain = O.OracleIndex.build(texts, O.ALPHABETS["ascii_dna"](), "u32", lookup_depth=0)
    for d in range(depth + 1):
        table = with_tables.lookup_table(d)
        assert table.shape[0] == 4 ** d
        for i in range(0, 4 ** d, max(1, 4 ** d // 64)):
            q = bytes(b"ACGT"[(i // 4 ** j) % 4] for j in range(d))  # first symbol = lowest digit
            s, e = plain.cursor_for_query(q)
            assert e - s == table[i][1] - table[i][0]
            if e > s:
                assert (s, e) == tuple(table[i])


def test_from_parts_equals_build():
    rng = random.Random(7)
    texts = [bytes(rng.choice(b"ACGTN") for _ in range(5000)), b"", bytes(rng.choice(b"ACGT") for _ in range(3000))]
    alph = O.ALPHABETS["ascii_dna_with_n"]()
    a = O.OracleIndex.build(texts, alph, "u32", sampling_rate=4, lookup_depth=3)
    rows, pos = a.border()
    b = O.OracleIndex.from_parts(a.bwt(), alph, a.count_array(), a.sentinel_indices(), a.samples(), 4,
                                 rows, pos, lookup_depth=3, storage="u32", nthreads=3)
    assert np.array_equal(a.blocks(), b.blocks())
    assert np.array_equal(a.block_offsets(), b.block_offsets())
    assert np.array_equal(a.superblock_offsets(), b.superblock_offsets())
    qs = [bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(12))) for _ in range(300)]
    assert a.locate_many(qs) == b.locate_many(qs, nthreads=4)
    assert a.count_many(qs).tolist() == b.count_many(qs, nthreads=3).tolist()


# ---- gdxo_verify_against_text: the check that makes bench.py's CPU arm independent of whoever built its BWT ----
def _parts_index(texts, alph, s=4, corrupt=None):
    oa = O.ALPHABETS[alph]()
    full = O.OracleIndex.build(texts, oa, "u32", sampling_rate=s, lookup_depth=0)
    bwt = full.bwt().copy()
    samples = full.samples().copy()
    rows, pos = full.border()
    if corrupt == "bwt":  # swap two different neighbouring symbols: same counts, wrong order
        i = next(i for i in range(len(bwt) - 1) if bwt[i] != bwt[i + 1] and bwt[i] and bwt[i + 1])
        bwt[i], bwt[i + 1] = bwt[i + 1], bwt[i]
    if corrupt == "sample":
        samples[len(samples) // 2] += 1
    if corrupt == "border":
        pos = pos.copy()
        pos[0] += 1
    part = O.OracleIndex.from_parts(bwt, oa, full.count_array(), full.sentinel_indices(), samples, s, rows, pos,
                                    lookup_depth=0, storage="u32", nthreads=2)
    return part, full.dense_text()


@pytest.mark.parametrize("texts", [
    [b"cccaaagggttt"], [b"ACGT" * 300], [b"", b"ACGTN", b"", b"ACGTN", b"GATTACA" * 40],
    [bytes(random.Random(3).choice(b"ACGTN") for _ in range(30_000)), bytes(random.Random(4).choice(b"ACGT") for _ in range(9_000))],
])
def test_verify_against_text_accepts_the_true_index(texts):
    part, dense = _parts_index(texts, "ascii_dna_with_n")
    bad, seen = part.verify_against_text(dense, nthreads=3)
    assert (bad, seen) == (0, len(dense))


@pytest.mark.parametrize("corrupt", ["bwt", "sample", "border"])
def test_verify_against_text_rejects_a_wrong_index(corrupt):
    rng = random.Random(9)
    texts = [bytes(rng.choice(b"ACGT") for _ in range(20_000)), bytes(rng.choice(b"ACGTN") for _ in range(5_000))]
    part, dense = _parts_index(texts, "ascii_dna_with_n", corrupt=corrupt)
    bad, _ = part.verify_against_text(dense, nthreads=2)
    assert bad > 0
    other = np.array(dense).copy()  # the right index against another text
    other[100] = 1 + (other[100] % 4)
    good, _ = _parts_index(texts, "ascii_dna_with_n")
    assert good.verify_against_text(other, nthreads=2)[0] > 0


# ---- tests/text_with_rank_support.rs:121-136 correctness_random_texts, for all four variants (:69-75) ----------
@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.integers(1, 255).flatmap(lambda mx: st.tuples(st.lists(st.integers(0, mx), max_size=300), st.just(mx + 1))))
def test_rank_variants_against_naive(case):
    text, sigma = case
    for variant, bits, storage in (("condensed", 64, "i32"), ("condensed", 512, "u32"), ("flat", 64, "i64"),
                                   ("flat", 512, "i32")):
        r = O.OracleVariantRank(text, sigma, storage, variant, bits)
        occ = [0] * sigma
        for idx in range(len(text) + 1):
            for symbol in {0, sigma - 1, text[idx - 1] if idx else 0, text[idx] if idx < len(text) else 0}:
                assert r.rank(symbol, idx) == occ[symbol], (variant, bits, symbol, idx)
            if idx < len(text):
                assert r.symbol_at(idx) == text[idx]
                occ[text[idx]] += 1
