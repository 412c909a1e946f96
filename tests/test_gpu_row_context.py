"""Row context table accelerator (gdx_index_set_row_context_table): SA[row] + the 45 text symbols in front of it in
one 16-byte entry.  It may only change how fast a one-row interval is verified, never a result: every case is compared
with the CPU oracle with the table on and off, and the cases are chosen to sit on the edges of its fast path -- context
windows cut short by text borders, by the start of the text and by `N`, queries with more symbols left than an entry
holds, queries longer than the staged 64 symbols, invalid bytes left of the verified part (lazy error), and the same
through the host packer (2-bit packed kernel)."""
import random

import numpy as np
import pytest

from oracle import oracle as O
import gdx_testutil as util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gdx():
    import genedex_b200
    assert genedex_b200._lib.load().gdx_device_count() >= 1
    return genedex_b200


def _edge_texts(rng):
    """Many short texts (borders every few symbols up to a few hundred), one long text with N runs of every length."""
    texts = [bytes(rng.choice(b"ACGT") for _ in range(m)) for m in (1, 2, 3, 7, 15, 16, 17, 44, 45, 46, 47, 63, 64, 65, 90, 91, 200)]
    long = bytearray(rng.choice(b"ACGT") for _ in range(30_000))
    p = 50
    for run in list(range(1, 12)) + [40, 44, 45, 46, 60, 100]:
        long[p:p + run] = b"N" * run
        p += run + rng.choice((3, 17, 44, 45, 46, 47, 80))
    texts.append(bytes(long))
    # a repeat-rich stretch: intervals stay wide for a while, so the verification starts with few symbols left
    unit = bytes(rng.choice(b"ACGT") for _ in range(37))
    texts.append(b"".join(bytes(b if rng.random() > 0.03 else rng.choice(b"ACGT") for b in unit) for _ in range(120)))
    return texts


def _edge_queries(rng, texts):
    qs = []
    for t in texts:
        n = len(t)
        for _ in range(60 if n < 1000 else 1500):
            m = rng.choice((1, 2, 5, 16, 20, 33, 45, 46, 47, 50, 61, 62, 63, 64, 65, 66, 80, 130))
            if m > n:
                m = n
            p = rng.randrange(0, n - m + 1)
            q = bytearray(t[p:p + m])
            kind = rng.randrange(8)
            if kind == 0 and m > 1:      # a mismatch somewhere in the part the text comparison sees
                j = rng.randrange(m)
                q[j] = rng.choice(b"ACGT")
            elif kind == 1:              # reaches over the start of the text / the border in front of it
                q = bytearray(rng.choice(b"ACGT") for _ in range(rng.randrange(1, 4))) + q
            elif kind == 2 and m > 2:    # an invalid byte: reported only if the search gets there
                q[rng.randrange(m)] = ord("x")
            qs.append(bytes(q))
    rng.shuffle(qs)
    return qs


@pytest.mark.parametrize("depth,seed", [(0, 0), (0, 8), (3, 0), (0, 12)])
def test_results_do_not_depend_on_the_row_context_table(gdx, depth, seed):
    rng = random.Random(100 + depth + seed)
    texts = _edge_texts(rng)
    oidx, pidx = util.build_pair(gdx, texts, "ascii_dna_with_n", "u32", 4, depth)
    pidx.set_seed_table_depth(seed)
    qs = [q for q in _edge_queries(rng, texts)]
    valid = [q for q in qs if b"x" not in q and not (depth and b"N" in q[-depth:])]
    for on in (True, False, True):
        pidx.set_row_context_table(on)
        assert pidx.info().row_context_entry_bytes == (16 if on else 0)
        util.assert_same_results(oidx, pidx, valid)
        # lazily reported invalid bytes: the same queries fail / pass with and without the table
        for q in [q for q in qs if b"x" in q][:150]:
            try:
                want = int(oidx.count_many([q])[0])     # the batched path: lazy translation
            except O.OraclePanic:
                want = None
            if want is None:
                with pytest.raises(gdx.InvalidSymbolError):
                    pidx.count_many([q])
            else:
                assert pidx.count_many([q]) == [want]


def test_row_context_with_the_packed_kernel_and_large_batches(gdx):
    """> 1 MB of queries: the host packer engages, k_search<PACKED> compares the register-held tail with the entry."""
    rng = random.Random(7)
    nrng = np.random.default_rng(7)
    n = 1_500_000
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[nrng.integers(0, 4, n)].copy()
    for s in nrng.integers(0, n - 200, 300):
        text[s:s + int(nrng.integers(1, 120))] = ord("N")
    texts = [text[:900_000].tobytes(), text[900_000:900_040].tobytes(), text[900_040:].tobytes()]
    oidx, pidx = util.build_pair(gdx, texts, "ascii_dna_with_n", "u32", 8, 0)
    qs = []
    for i in range(120_000):
        m = rng.choice((12, 30, 45, 46, 50, 60, 64, 65, 100))
        p = rng.randrange(0, n - m)
        q = bytearray(text[p:p + m].tobytes())
        if i % 3 == 0:
            q[rng.randrange(m)] = rng.choice(b"ACGT")
        qs.append(bytes(q))
    data, off = O.pack(qs)
    want_c = oidx.count_many_packed(data, off)
    want_o, want_h = oidx.locate_many_packed(data, off)
    for on in (True, False):
        pidx.set_row_context_table(on)
        assert np.array_equal(pidx.count_many_packed(data, off), want_c)
        got_o, got_h = pidx.locate_many_packed(data, off)
        assert np.array_equal(got_o, want_o) and np.array_equal(got_h, want_h)
        st = pidx.stats()
        assert st.packed_queries > 0


def test_row_context_policy(gdx):
    texts = [b"ACGTACGTTTGACA" * 50]
    never = gdx.FmIndexConfig("u32").row_context_table(False).construct_index(texts, gdx.alphabet.ascii_dna())
    assert never.info().row_context_entry_bytes == 0
    auto = gdx.FmIndexConfig("u32").construct_index(texts, gdx.alphabet.ascii_dna())
    assert auto.info().row_context_entry_bytes == 16      # ample memory: built
    # does not apply: more than 4 searchable symbols
    iupac = gdx.FmIndexConfig("u32").construct_index([b"ACGTRYKM" * 20], gdx.alphabet.ascii_dna_iupac())
    assert iupac.info().row_context_entry_bytes == 0
    with pytest.raises(gdx.GenedexError):
        iupac.set_row_context_table(True)
    # 64-bit storage of a short text is still a text shorter than 2^32: allowed
    wide = gdx.FmIndexConfig("i64").construct_index(texts, gdx.alphabet.ascii_dna())
    wide.set_row_context_table(True)
    assert wide.count_many([b"TTGACA", b"ACGTACGTTTGACAACGT", b"GGGG"]) == [50, 49, 0]
