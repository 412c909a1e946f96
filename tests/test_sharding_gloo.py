"""world_size-2 gloo test (CPU) of the multi-process plumbing of the N > 1 path: query sharding by the
library's contiguous ranges (gdx_shard_range), the out-of-band hand-over of the 128-byte NCCL unique id
that gdx_index_broadcast needs (torch_share_id; the broadcast itself needs GPUs: tests/test_gpu_multi.py),
and the order-preserving gather of per-rank results.  The per-rank "search" here is the CPU oracle --
the product's search needs a GPU -- so this checks the host logic around the C ABI only."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from genedex_b200 import replicate as R
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the unique id of the root reaches every rank unchanged
        uid = bytes(np.random.default_rng(123).integers(0, 256, 128, dtype=np.uint8))
        got_uid = R.torch_share_id(rank)(uid if rank == 0 else None)
        assert got_uid == uid
        # 2. every rank searches its contiguous shard on its own replica, rank 0 gathers in order
        trng = np.random.default_rng(7)
        text = np.frombuffer(b"ACGT", dtype=np.uint8)[trng.integers(0, 4, 50_000)].tobytes()
        idx = O.OracleIndex.build([text], O.ALPHABETS["ascii_dna"](), "u32", 4, 3)  # the "replica"
        nq = 1001
        queries = [text[p:p + 12] for p in trng.integers(0, 49_000, nq)]
        b, e = R.shard_range(nq, rank, world)
        local = idx.count_many(queries[b:e])
        got = R.gather_in_order(local, nq, rank, world)
        if rank == 0:
            want = idx.count_many(queries)
            assert np.array_equal(got, want)
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    sys.path.insert(0, ROOT)
    from genedex_b200.replicate import shard_range
    for n in (0, 1, 7, 8, 1001, 60_000_000):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_broadcast_shard_gather(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
