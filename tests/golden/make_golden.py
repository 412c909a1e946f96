"""Writes tests/golden/*.json.

reference_kats.json   the reference's own known-answer vectors for the search path, transcribed by hand
                      from its tests/examples (each case cites file:line under /root/reference).  The
                      Rust reference cannot be run here, so nothing in this file is machine-generated.
oracle_vectors.json   seeded random cases answered by the CPU oracle (which is pinned by the KATs above
                      and the reference's property tests); regenerate with  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def lat(b: bytes) -> str:
    return b.decode("latin1")


REFERENCE_KATS = [
    {"cite": "tests/fmindex.rs:7-80 (create_index, basic_search, text_front_search, search_no_wrapping)",
     "alphabet": "ascii_dna", "storages": ["i32", "u32"], "sampling_rate": 3, "lookup_depth": 0,
     "texts": ["cccaaagggttt"],
     "queries": [{"q": "gg", "hits": [[0, 6], [0, 7]]}, {"q": "c", "hits": [[0, 0], [0, 1], [0, 2]]},
                 {"q": "ta", "hits": []}]},
    {"cite": "tests/fmindex.rs:82-126 (search_multitext)",
     "alphabet": "ascii_dna", "storages": ["u32"], "sampling_rate": 3, "lookup_depth": 4,
     "texts": ["cccaaagggttt", "acgtacgtacgt"],
     "queries": [{"q": "gg", "hits": [[0, 6], [0, 7]]}, {"q": "gt", "hits": [[0, 8], [1, 2], [1, 6], [1, 10]]}]},
    {"cite": "tests/fmindex.rs:128-154 (u8_alphabet)",
     "alphabet": "u8_until_8", "storages": ["u32"], "sampling_rate": 3, "lookup_depth": 4,
     "texts": [lat(bytes([0, 4, 3, 2, 1, 5, 8, 6, 7, 8])), lat(bytes([5, 7, 3, 4, 2, 1, 5, 8])), ""],
     "queries": [{"q": lat(bytes([1, 5, 8])), "hits": [[0, 4], [1, 5]]}]},
    {"cite": "examples/basic_usage.rs:8-16, src/lib.rs:15-32, README.md:33",
     "alphabet": "ascii_dna_with_n", "storages": ["i32"], "sampling_rate": 2, "lookup_depth": 0,
     "texts": ["aACGT", "acGtn"], "queries": [{"q": "GT", "count": 2}]},
    {"cite": "examples/cursor.rs:6-24 (cursor GT -> 5, extend_query_front C -> 2)",
     "alphabet": "ascii_dna_with_n", "storages": ["i32"], "sampling_rate": 4, "lookup_depth": 0,
     "texts": ["AaACGT", "AacGtn", "GTGTGT"], "queries": [{"q": "GT", "count": 5}, {"q": "CGT", "count": 2}]},
    {"cite": "tests/fmindex.proptest-regressions:9 (texts = [[]], sampling rate 1); naive_search tests/fmindex.rs:207-227",
     "alphabet": "ascii_dna", "storages": ["i32", "u32", "i64"], "sampling_rate": 1, "lookup_depth": 0,
     "texts": [""], "queries": [{"q": "", "hits": [[0, 0]]}, {"q": "A", "hits": []}]},
]


def oracle_vectors():
    from oracle import oracle as O
    rng = random.Random(20261017)
    cases = []
    for alph, storage, s, depth in (("ascii_dna_with_n", "u32", 4, 0), ("ascii_dna", "i32", 7, 3),
                                    ("protein20", "u32", 4, 2), ("ascii_dna_iupac", "i64", 5, 1)):
        oa = O.ALPHABETS[alph]()
        syms = bytes(oa.dense_to_io)
        search = syms[: oa.num_searchable]
        texts = [bytes(rng.choice(syms) for _ in range(rng.randrange(0, 400))) for _ in range(3)]
        qs = []
        for _ in range(40):
            t = rng.choice([x for x in texts if x] or [b"A"])
            p = rng.randrange(len(t))
            q = t[p:p + rng.randrange(0, 30)]
            if depth > 0 and any(c not in search for c in q):
                continue
            qs.append(q)
        qs += [bytes(rng.choice(search) for _ in range(rng.randrange(0, 12))) for _ in range(20)]
        idx = O.OracleIndex.build(texts, oa, storage, s, depth)
        starts, ends = idx.cursors_many(qs)
        hits = idx.locate_many(qs)
        cases.append({"alphabet": alph, "storage": storage, "sampling_rate": s, "lookup_depth": depth,
                      "texts": [lat(t) for t in texts],
                      "queries": [{"q": lat(q), "interval": [int(a), int(b)], "hits": [list(h) for h in hs]}
                                  for q, a, b, hs in zip(qs, starts, ends, hits)]})
    return cases


if __name__ == "__main__":
    json.dump(REFERENCE_KATS, open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1)
    json.dump(oracle_vectors(), open(os.path.join(HERE, "oracle_vectors.json"), "w"))
    print("wrote", os.listdir(HERE))
