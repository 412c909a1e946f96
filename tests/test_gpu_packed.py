"""GPU parity of the transfer-side changes of round 2: host 2-bit packing of IO-byte batches (+ the exception
path for bytes without a code), pre-packed input, uint32 results, and the sharded entry points -- all against
the CPU oracle, bit-exact, through the C ABI."""
import ctypes as C
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


@pytest.fixture(scope="module")
def dna_case(gdx):
    """2 Mbp two-text DNA index with N runs + 400 k queries: most are plain ACGT windows / random strings,
    a few per cent hold N (valid, not searchable) -- enough bytes (> 1 MB) for the packer to engage."""
    rng = np.random.default_rng(21)
    n = 2_000_000
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
    for s in rng.integers(0, n - 3000, 40):
        text[s:s + int(rng.integers(1, 3000))] = ord("N")
    texts = [text[:1_200_000].tobytes(), text[1_200_000:].tobytes()]
    oa = util.oracle_alphabet("ascii_dna_with_n")
    oidx = O.OracleIndex.build(texts, oa, "u32", sampling_rate=4, lookup_depth=0)
    pidx = gdx.FmIndexConfig("u32").construct_index(texts, gdx.alphabet.ascii_dna_with_n())
    qs = []
    prng = random.Random(5)
    for i in range(400_000):
        kind = i % 8 if i % 32 else 7   # ~3 % of the queries are random strings over ACGTN: exceptions of the packer
        kind = 5 if kind == 7 and i % 32 else kind
        m = prng.randrange(6, 90) if kind == 7 else prng.randrange(20, 60)
        if kind < 5:
            p = prng.randrange(0, n - 100)
            qs.append(text[p:p + m].tobytes())      # may run into an N stretch: exception query
        elif kind == 5:
            qs.append(bytes(prng.choice(b"ACGT") for _ in range(m)))
        elif kind == 6:
            q = bytearray(text[(p := prng.randrange(0, n - 100)):p + m].tobytes())
            if q:
                q[prng.randrange(len(q))] = prng.choice(b"ACGTacgt")
            qs.append(bytes(q))
        else:
            qs.append(bytes(prng.choice(b"ACGTN") for _ in range(m)))
    data, off = O.pack(qs)
    return dict(oidx=oidx, pidx=pidx, data=data, off=off, qs=qs, texts=texts)


def test_packed_host_path_matches_oracle(gdx, dna_case):
    c = dna_case
    oidx, pidx, data, off = c["oidx"], c["pidx"], c["data"], c["off"]
    nq = off.size - 1
    want_s, want_e = oidx.cursors_many_packed(data, off)
    got = pidx.count_many_packed(data, off)
    st = pidx.stats()
    assert np.array_equal(got, want_e - want_s)
    n_exc = sum(1 for q in c["qs"] if b"N" in q)
    assert st.packed_queries == nq, "every chunk should have crossed PCIe packed"
    assert st.exception_queries == n_exc and n_exc > 1000
    assert st.h2d_bytes < 0.6 * data.size, "packing must cut the PCIe bytes (2-bit symbols + offsets + exception queries)"
    gs, ge = pidx.cursors_many_packed(data, off)
    assert np.array_equal(gs, want_s) and np.array_equal(ge, want_e)
    sub = 120_000
    ooff, ohits = oidx.locate_many_packed(data[: int(off[sub])], off[: sub + 1])
    poff, phits = pidx.locate_many_packed(data[: int(off[sub])], off[: sub + 1])
    assert np.array_equal(ooff, poff) and np.array_equal(ohits, phits)
    assert pidx.stats().packed_queries == sub


def test_uint32_results(gdx, dna_case):
    c = dna_case
    oidx, pidx, data, off = c["oidx"], c["pidx"], c["data"], c["off"]
    nq = off.size - 1
    want_s, want_e = oidx.cursors_many_packed(data, off)
    c32 = np.full(nq, 0xDEADBEEF, dtype=np.uint32)
    pidx.count_many_packed(data, off, out=c32)
    assert np.array_equal(c32.astype(np.uint64), want_e - want_s)
    s32, e32 = np.zeros(nq, dtype=np.uint32), np.zeros(nq, dtype=np.uint32)
    pidx.cursors_many_packed(data, off, out=(s32, e32))
    assert np.array_equal(s32.astype(np.uint64), want_s) and np.array_equal(e32.astype(np.uint64), want_e)
    few = 100  # small batch: direct D2H into the caller's uint32 array
    c_few = np.zeros(few, dtype=np.uint32)
    pidx.count_many_packed(data[: int(off[few])], off[: few + 1], out=c_few)
    assert np.array_equal(c_few.astype(np.uint64), (want_e - want_s)[:few])


def test_packed_batch_reports_invalid_symbols_lazily(gdx, dna_case):
    """an invalid byte is an exception of the packer; the query is re-run from its IO bytes, so the reference's
    lazy panic (batch_computed_cursors.rs:84-87,106-113) and the index of the first offending query survive"""
    c = dna_case
    oidx, pidx = c["oidx"], c["pidx"]
    text = c["texts"][0]
    qs = [text[1000 + 19 * i: 1050 + 19 * i] for i in range(60_000)]  # 3 MB of length-50 windows inside text 0
    assert all(len(q) == 50 for q in qs)
    qs = [q if b"N" not in q else b"ACGTACGT" for q in qs]
    absent = b"ACGT" * 12 + b"!!"          # '!' is left of where the interval of this random 48-mer is long empty...
    qs[30_000] = b"!" + b"ACGT" * 12 + b"A"   # ... so an invalid byte at the far left of an absent query is never reached
    data, off = O.pack(qs)
    want = oidx.count_many_packed(data, off)
    assert want[30_000] == 0
    assert np.array_equal(pidx.count_many_packed(data, off), want)
    assert pidx.stats().exception_queries == 1
    qs[41_234] = qs[41_234][:-1] + b"!"      # the rightmost symbol is always translated
    qs[50_001] = qs[50_001][:-3] + b"!" + qs[50_001][-2:]
    qs[59_999] = absent
    data, off = O.pack(qs)
    with pytest.raises(O.OraclePanic):
        oidx.count_many_packed(data, off)
    for call in (lambda: pidx.count_many_packed(data, off), lambda: pidx.cursors_many_packed(data, off),
                 lambda: pidx.locate_many_packed(data, off)):
        with pytest.raises(gdx.InvalidSymbolError) as ei:
            call()
        assert ei.value.query == 41_234


def test_mostly_unencodable_batch_falls_back_to_io_bytes(gdx, dna_case):
    c = dna_case
    oidx, pidx = c["oidx"], c["pidx"]
    prng = random.Random(9)
    qs = [bytes(prng.choice(b"ACGTNN") for _ in range(40)) for _ in range(60_000)]
    data, off = O.pack(qs)
    want = oidx.count_many_packed(data, off)
    assert np.array_equal(pidx.count_many_packed(data, off), want)
    st = pidx.stats()
    assert st.packed_queries == 0 and st.exception_queries == 0


def test_prepacked_input_host_and_device(gdx, dna_case):
    import torch
    c = dna_case
    oidx, pidx = c["oidx"], c["pidx"]
    text = c["texts"][0]
    prng = random.Random(3)
    qs = []
    while len(qs) < 150_000:
        p = prng.randrange(0, len(text) - 80)
        q = text[p:p + prng.randrange(9, 70)]   # (not shorter: a length-0 query alone has 2 M hits)
        if b"N" not in q:
            qs.append(q)
    data, off = O.pack(qs)
    nq = len(qs)
    want_s, want_e = oidx.cursors_many_packed(data, off)
    packed, first_bad = pidx.pack_queries_2bit(data, off)
    assert first_bad is None and packed.size == (int(off[-1]) + 3) // 4 + (-((int(off[-1]) + 3) // 4)) % 4
    enc = gdx._lib.GDX_QUERIES_PACKED_2BIT
    assert np.array_equal(pidx.count_many_packed(packed, off, encoding=enc), want_e - want_s)
    st = pidx.stats()
    assert st.packed_queries == nq and st.exception_queries == 0 and st.h2d_bytes < data.size // 2
    gs, ge = pidx.cursors_many_packed(packed, off, encoding=enc)
    assert np.array_equal(gs, want_s) and np.array_equal(ge, want_e)
    ooff, ohits = oidx.locate_many_packed(data, off)
    poff, phits = pidx.locate_many_packed(packed, off, encoding=enc)
    assert np.array_equal(ooff, poff) and np.array_equal(ohits, phits)
    # a batch with an unencodable byte is refused by the packer's report, not silently mis-searched
    bad = bytearray(data.tobytes())
    bad[int(off[77]) + 1] = ord("N")
    _, first_bad = pidx.pack_queries_2bit(np.frombuffer(bytes(bad), dtype=np.uint8), off)
    assert first_bad == 77
    # fixed-length form, odd length: sub-batches of the stream do not start on byte boundaries
    m = 37
    fixed = [text[p:p + m] for p in (prng.randrange(0, len(text) - m) for _ in range(400_000))]
    fixed = [q for q in fixed if b"N" not in q][:120_001]
    fdata = np.frombuffer(b"".join(fixed), dtype=np.uint8)
    foff = np.arange(len(fixed) + 1, dtype=np.uint64) * m
    fwant = oidx.count_many_packed(fdata, foff)
    fpacked, _ = pidx.pack_queries_2bit(fdata, None, m, len(fixed))
    assert np.array_equal(pidx.count_many_packed(fpacked, None, m, len(fixed), encoding=enc), fwant)
    # device-resident packed batch
    lib = gdx._lib.load()
    d_p = torch.from_numpy(fpacked).cuda()
    d_c = torch.zeros(len(fixed), dtype=torch.int64, device="cuda")
    d_err = torch.full((1,), -1, dtype=torch.int64, device="cuda")
    dq = gdx._lib.gdx_queries(d_p.data_ptr(), None, m, len(fixed), enc, 0)
    assert lib.gdx_count_many_device(pidx.handle, C.byref(dq), d_c.data_ptr(), d_err.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    assert int(d_err.item()) == -1
    assert np.array_equal(d_c.cpu().numpy().astype(np.uint64), fwant)


def test_sharded_calls_on_one_device(gdx, dna_case):
    """gdx_*_many_sharded with three replicas that all live on device 0 (replicate onto the source device =
    device-to-device copies): the range cutting, per-shard threads, input-order results, error merging and the
    segmented hit arrays are exercised without a second GPU (tests/test_gpu_multi.py does the real thing)."""
    from genedex_b200.replicate import ReplicaSet, shard_range
    c = dna_case
    oidx, pidx, data, off = c["oidx"], c["pidx"], c["data"], c["off"]
    nq = off.size - 1
    rs = ReplicaSet.replicate(pidx, [0, 0])
    assert len(rs.replicas) == 3 and all(r.info().device == 0 for r in rs.replicas)
    want_s, want_e = oidx.cursors_many_packed(data, off)
    got = rs.count_many_packed(data, off)
    assert np.array_equal(got, want_e - want_s)
    st = rs.stats()
    assert st.shards == 3 and st.queries == nq
    gs, ge = rs.cursors_many_packed(data, off)
    assert np.array_equal(gs, want_s) and np.array_equal(ge, want_e)
    sub = 150_001
    sdata, soff = data[: int(off[sub])], off[: sub + 1]
    ooff, ohits = oidx.locate_many_packed(sdata, soff)
    hit_off, views, first, release = rs.locate_many_view(sdata, soff)
    try:
        assert np.array_equal(hit_off, ooff)
        assert int(first[-1]) == ohits.shape[0]
        assert np.array_equal(np.concatenate(views), ohits)
        for k in range(3):
            b, e = shard_range(sub, k, 3)
            assert int(hit_off[b]) == int(first[k]) and int(hit_off[e]) == int(first[k + 1])
    finally:
        release()
    # a process that owns only the middle shard writes only that range
    mid = ReplicaSet([rs.replicas[1]], first_shard=1, n_shards=3)
    out = np.full(nq, 12345, dtype=np.uint64)
    mid.count_many_packed(data, off, out=out)
    b, e = shard_range(nq, 1, 3)
    assert np.array_equal(out[b:e], (want_e - want_s)[b:e])
    assert np.all(out[:b] == 12345) and np.all(out[e:] == 12345)
    # errors: the first offending query of the whole batch is reported, whatever shard it is in
    qs = list(c["qs"][:90_000])
    qs[70_000] = b"AC!T"
    qs[80_000] = b"AC!T"
    bdata, boff = O.pack(qs)
    with pytest.raises(gdx.InvalidSymbolError) as ei:
        rs.count_many_packed(bdata, boff)
    assert ei.value.query == 70_000


def test_a_result_that_cannot_fit_is_refused_and_leaves_the_index_usable(gdx, dna_case):
    """30 000 empty queries match every row (lib.rs:202-210: the empty cursor is (0, n)): 6 * 10^10 hits.  The
    library says GDX_ERR_OOM instead of dying, and the next call is not poisoned by the failed allocation."""
    c = dna_case
    pidx, oidx = c["pidx"], c["oidx"]
    qs = [b""] * 30_000 + [b"ACGT"]
    data, off = O.pack(qs)
    n = pidx.total_text_len()
    assert pidx.count_many_packed(data, off).tolist() == [n] * 30_000 + [int(oidx.count(b"ACGT"))]
    with pytest.raises(MemoryError):
        pidx.locate_many_packed(data, off)
    data2, off2 = O.pack([b"ACGTAC", b"TTTTTTTTTTTTTTTTTTTTTTTTTTTTTT"])
    ooff, ohits = oidx.locate_many_packed(data2, off2)
    poff, phits = pidx.locate_many_packed(data2, off2)
    assert np.array_equal(ooff, poff) and np.array_equal(ohits, phits)


def test_compact_locate_results(gdx, dna_case):
    """gdx_locate_many_compact / _sharded_compact: u32 hits per query + gdx_hit32 hits = the same hits in the same order"""
    from genedex_b200.replicate import ReplicaSet, shard_range
    c = dna_case
    oidx, pidx, data, off = c["oidx"], c["pidx"], c["data"], c["off"]
    sub = 150_001
    sdata, soff = data[: int(off[sub])], off[: sub + 1]
    ooff, ohits = oidx.locate_many_packed(sdata, soff)
    counts, hits, release = pidx.locate_many_compact_view(sdata, soff)
    try:
        assert counts.dtype == np.uint32 and hits.dtype == np.uint32
        assert np.array_equal(counts.astype(np.uint64), ooff[1:] - ooff[:-1])
        assert np.array_equal(hits.astype(np.uint64), ohits)
    finally:
        release()
    rs = ReplicaSet.replicate(pidx, [0])
    counts2, views, release2 = rs.locate_many_compact_view(sdata, soff)
    try:
        assert np.array_equal(counts2.astype(np.uint64), ooff[1:] - ooff[:-1])
        assert np.array_equal(np.concatenate(views).astype(np.uint64), ohits)
        b, e = shard_range(sub, 1, 2)
        assert views[0].shape[0] == int(ooff[b]) and views[1].shape[0] == int(ooff[e] - ooff[b])
    finally:
        release2()
    few, fhits, frel = pidx.locate_many_compact_view(*O.pack([b"ACGTAC", b"", b"TTTTTTTTTTTTTTTTTTTTTTTTT"][::2]))
    o2, h2 = oidx.locate_many_packed(*O.pack([b"ACGTAC", b"TTTTTTTTTTTTTTTTTTTTTTTTT"]))
    assert np.array_equal(few.astype(np.uint64), o2[1:] - o2[:-1]) and np.array_equal(fhits.astype(np.uint64), h2)
    frel()


def test_offsets_that_leave_the_batch_are_refused(gdx, dna_case):
    """A C caller can hand over offsets a Rust slice-of-slices cannot express (decreasing, or past the bytes).  Nothing is
    read outside the batch: the kernel checks every query against the extent of the uploaded chunk, the host the chunk
    boundaries; the call fails with GDX_ERR_BAD_ARG naming the query, and the index stays usable."""
    c = dna_case
    pidx, oidx, data, off = c["pidx"], c["oidx"], c["data"], c["off"]
    lib = gdx._lib.load()
    small, small_off = O.pack([b"ACGTACGT", b"ACG", b"TTGACA", b"GATTACA"])
    for bad_at, value in ((2, 3), (2, 10**12)):          # decreasing / far outside (the total stays plausible)
        broken = small_off.copy()
        broken[bad_at] = value
        with pytest.raises(gdx.GenedexError) as e:
            pidx.count_many_packed(small, broken)
        assert e.value.status == gdx._lib.GDX_ERR_BAD_ARG and not isinstance(e.value, gdx.InvalidSymbolError)
        assert int(lib.gdx_last_error_query()) in (bad_at - 1, bad_at)
    # the large-batch forms: packed chunks with 32-bit chunk-relative offsets, and the same with packing off
    for where in (7, 123_457, off.size - 3):
        broken = off.copy()
        broken[where] = broken[where + 1] + 5            # query `where` gets a negative length ...
        with pytest.raises(gdx.GenedexError) as e:
            pidx.count_many_packed(data, broken)
        assert e.value.status == gdx._lib.GDX_ERR_BAD_ARG
        assert int(lib.gdx_last_error_query()) in (where - 1, where)
        with pytest.raises(gdx.GenedexError):
            pidx.locate_many_packed(data, broken)
    assert np.array_equal(pidx.count_many_packed(data, off), oidx.count_many_packed(data, off))
