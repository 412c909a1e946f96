"""debug: locate on a packed batch that holds invalid symbols"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import genedex_b200 as gdx
from oracle import oracle as O
import gdx_testutil as util

rng = np.random.default_rng(21)
n = 2_000_000
text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
texts = [text[:1_200_000].tobytes(), text[1_200_000:].tobytes()]
pidx = gdx.FmIndexConfig("u32").construct_index(texts, gdx.alphabet.ascii_dna_with_n())
oidx = O.OracleIndex.build(texts, util.oracle_alphabet("ascii_dna_with_n"), "u32", 4, 0)
t0 = texts[0]
qs = [t0[1000 + 40 * i: 1050 + 40 * i] for i in range(60_000)]
for variant in ("clean", "one_invalid_last", "absent_tail"):
    q2 = list(qs)
    if variant == "one_invalid_last":
        q2[41_234] = q2[41_234][:-1] + b"!"
    if variant == "absent_tail":
        q2[59_999] = b"ACGT" * 12 + b"!!"
    data, off = O.pack(q2)
    nq = len(q2)
    s, e = np.zeros(nq, dtype=np.uint64), np.zeros(nq, dtype=np.uint64)
    try:
        pidx.cursors_many_packed(data, off, out=(s, e))
        print(variant, "cursors ok")
    except gdx.InvalidSymbolError as ex:
        print(variant, "cursors raised, query", ex.query)
    w = (e - s).astype(np.int64)
    print("  widths: min", w.min(), "max", w.max(), "sum", w.sum(), "n_big", int((w > 100).sum()), "first big", np.flatnonzero(w > 100)[:5])
    try:
        o, h = pidx.locate_many_packed(data, off)
        print(variant, "locate ok hits", h.shape)
    except Exception as ex:
        print(variant, "locate raised", type(ex).__name__, str(ex)[:100])
