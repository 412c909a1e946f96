"""The committed golden fixtures (tests/golden/*.json) against the oracle (CPU) and the CUDA path (GPU)."""
import json
import os

import pytest

import gdx_testutil as util
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLDEN, "reference_kats.json")))
VECTORS = json.load(open(os.path.join(GOLDEN, "oracle_vectors.json")))


def b(s):
    return s.encode("latin1")


def _check_kat(index_for, case):
    for storage in case["storages"]:
        idx = index_for(case, storage)
        for q in case["queries"]:
            if "hits" in q:
                got = idx.locate(b(q["q"]))
                got = {(h.text_id, h.position) if hasattr(h, "text_id") else tuple(h) for h in got}
                assert got == {tuple(h) for h in q["hits"]}, (case["cite"], q["q"])
                assert idx.count(b(q["q"])) == len(q["hits"])
            else:
                assert idx.count(b(q["q"])) == q["count"], (case["cite"], q["q"])


@pytest.mark.parametrize("case", KATS, ids=[c["cite"].split(" ")[0] for c in KATS])
def test_reference_kats_oracle(case):
    _check_kat(lambda c, st: O.OracleIndex.build([b(t) for t in c["texts"]], util.oracle_alphabet(c["alphabet"]), st,
                                                 c["sampling_rate"], c["lookup_depth"]), case)


def test_oracle_vectors_are_reproducible():
    # the committed oracle vectors equal what the current oracle answers (guards against silent drift)
    for case in VECTORS:
        idx = O.OracleIndex.build([b(t) for t in case["texts"]], util.oracle_alphabet(case["alphabet"]),
                                  case["storage"], case["sampling_rate"], case["lookup_depth"])
        qs = [b(q["q"]) for q in case["queries"]]
        starts, ends = idx.cursors_many(qs)
        assert [[int(a), int(e)] for a, e in zip(starts, ends)] == [q["interval"] for q in case["queries"]]
        assert [[list(h) for h in hs] for hs in idx.locate_many(qs)] == [q["hits"] for q in case["queries"]]


@pytest.mark.gpu
@pytest.mark.parametrize("case", KATS, ids=[c["cite"].split(" ")[0] for c in KATS])
def test_reference_kats_cuda(case):
    import genedex_b200 as gdx

    def build(c, st):
        return (gdx.FmIndexConfig(st).suffix_array_sampling_rate(c["sampling_rate"])
                .lookup_table_depth(c["lookup_depth"])
                .construct_index([b(t) for t in c["texts"]], util.product_alphabet(gdx, c["alphabet"])))
    _check_kat(build, case)


@pytest.mark.gpu
@pytest.mark.parametrize("on_device", [False, True])
def test_oracle_vectors_cuda(on_device):
    import genedex_b200 as gdx
    for case in VECTORS:
        cfg = (gdx.FmIndexConfig(case["storage"]).suffix_array_sampling_rate(case["sampling_rate"])
               .lookup_table_depth(case["lookup_depth"]).construct_on_device(on_device, verify=True))
        idx = cfg.construct_index([b(t) for t in case["texts"]], util.product_alphabet(gdx, case["alphabet"]))
        qs = [b(q["q"]) for q in case["queries"]]
        cursors = idx.cursors_for_many_queries(qs)
        assert [list(c.interval) for c in cursors] == [q["interval"] for q in case["queries"]]
        got = [[[h.text_id, h.position] for h in hs] for hs in idx.locate_many(qs)]
        assert got == [q["hits"] for q in case["queries"]]
        assert idx.count_many(qs) == [len(q["hits"]) for q in case["queries"]]
