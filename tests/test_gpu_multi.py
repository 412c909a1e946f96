"""Multi-GPU tests (skipped below two devices; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
replicas made by the library's own NCCL broadcast, one batch sharded over them by the library, results against
the CPU oracle; and the one-process-per-GPU form (gdx_index_broadcast + n_local = 1 shards) under mp.spawn."""
import os
import random
import socket
import sys

import numpy as np
import pytest

from oracle import oracle as O
import gdx_testutil as util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    import genedex_b200
    return genedex_b200._lib.load().gdx_device_count()


def _case(n=1_500_000, nq=300_000, seed=31):
    rng = np.random.default_rng(seed)
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
    text[40_000:47_000] = ord("N")
    texts = [text[: n // 3].tobytes(), text[n // 3:].tobytes()]
    prng = random.Random(seed)
    qs = []
    for i in range(nq):
        if i % 3:
            p = prng.randrange(0, n - 80)
            qs.append(text[p:p + prng.randrange(15, 60)].tobytes())
        else:
            qs.append(bytes(prng.choice(b"ACGT") for _ in range(prng.randrange(8, 24))))
    return texts, qs


@pytest.mark.skipif(_ndev() < 2, reason="needs at least two GPUs")
def test_replicate_over_nccl_and_sharded_search():
    import genedex_b200 as gdx
    from genedex_b200.replicate import ReplicaSet, replicate_transport, shard_range
    ndev = _ndev()
    texts, qs = _case()
    oa = util.oracle_alphabet("ascii_dna_with_n")
    oidx = O.OracleIndex.build(texts, oa, "u32", sampling_rate=4, lookup_depth=0)
    pidx = gdx.FmIndexConfig("u32").device(0).construct_index(texts, gdx.alphabet.ascii_dna_with_n())
    rs = ReplicaSet.replicate(pidx, list(range(1, ndev)))
    assert replicate_transport() == "nccl", "the replicas must travel through the library's ncclBroadcast"
    assert [r.info().device for r in rs.replicas] == list(range(ndev))
    for r in rs.replicas[1:]:  # every replica derived its own accelerators, as the policy in the image says
        assert r.info().dense_suffix_array_bytes == pidx.info().dense_suffix_array_bytes
        assert r.info().seed_table_depth == pidx.info().seed_table_depth
    data, off = O.pack(qs)
    nq = len(qs)
    want_s, want_e = oidx.cursors_many_packed(data, off)
    got = rs.count_many_packed(data, off)
    assert np.array_equal(got, want_e - want_s)
    assert rs.stats().shards == ndev
    gs, ge = rs.cursors_many_packed(data, off)
    assert np.array_equal(gs, want_s) and np.array_equal(ge, want_e)
    ooff, ohits = oidx.locate_many_packed(data, off, nthreads=0)
    hit_off, views, first, release = rs.locate_many_view(data, off)
    try:
        assert np.array_equal(hit_off, ooff) and np.array_equal(np.concatenate(views), ohits)
    finally:
        release()
    # every replica on its own answers the whole batch identically (the image arrived intact)
    for r in rs.replicas[1:]:
        b, e = shard_range(nq, 0, 7)
        sub_off = off[b:e + 1]
        assert np.array_equal(r.count_many_packed(data, sub_off), (want_e - want_s)[b:e])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, tmpdir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import genedex_b200 as gdx
    from genedex_b200.replicate import ReplicaSet, broadcast_index, gather_in_order, replicate_transport, shard_range, torch_share_id
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # only carries the 128-byte unique id + the gather
    try:
        texts, qs = _case(n=600_000, nq=90_000, seed=77)
        qs = [q if b"N" not in q else b"ACGTACGTAC" for q in qs]  # (N cannot index a depth-3 lookup table: DESIGN.md)
        alphabet = gdx.alphabet.ascii_dna_with_n()
        pidx = None
        if rank == 0:
            pidx = gdx.FmIndexConfig("u32").device(0).lookup_table_depth(3).construct_index(texts, alphabet)
        local = broadcast_index(pidx, alphabet, rank, world, rank, torch_share_id(rank))
        assert replicate_transport() == "nccl"
        assert local.info().device == rank and local.info().lookup_table_depth == 3
        data, off = O.pack(qs)
        nq = len(qs)
        rs = ReplicaSet([local], first_shard=rank, n_shards=world)
        counts = np.zeros(nq, dtype=np.uint64)
        rs.count_many_packed(data, off, out=counts)
        b, e = shard_range(nq, rank, world)
        assert not counts[:b].any() and not counts[e:].any()
        got = gather_in_order(counts[b:e], nq, rank, world)
        if rank == 0:
            oidx = O.OracleIndex.build(texts, util.oracle_alphabet("ascii_dna_with_n"), "u32", 4, 3)
            assert np.array_equal(got, oidx.count_many_packed(data, off, nthreads=0))
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ndev() < 2, reason="needs at least two GPUs")
def test_one_process_per_gpu_broadcast_and_shards(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
