"""Multi-GPU layer of the host mirror: everything here is a thin wrapper over the library's own entry points
(include/genedex_b200.h): gdx_index_replicate / gdx_index_broadcast (one ncclBroadcast of the device image),
gdx_shard_range and gdx_*_many_sharded (one batch cut into contiguous ranges, one replica per range, results in
input order, no collective on the query path -- SURVEY 8e).

  single process, several GPUs     ReplicaSet.replicate(index, devices).count_many_packed(...)
  one process per GPU (torchrun)   broadcast_index(index_or_None, rank, world, device, share_id) and then
                                   ReplicaSet([local], first_shard=rank, n_shards=world)
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import numpy as np

from . import _lib
from .index import FmIndex, _check, _queries_struct


def shard_range(n_items: int, shard: int, n_shards: int) -> tuple[int, int]:
    """gdx_shard_range: contiguous, balanced, order-preserving split of [0, n_items)."""
    b, e = C.c_uint64(), C.c_uint64()
    _lib.load().gdx_shard_range(n_items, shard, n_shards, C.byref(b), C.byref(e))
    return int(b.value), int(e.value)


def replicate_transport() -> str:
    """What the last gdx_index_replicate / gdx_index_broadcast of this thread used: "nccl" or "peer"."""
    return _lib.load().gdx_replicate_transport().decode()


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * _lib.GDX_NCCL_UNIQUE_ID_BYTES)()
    _check(_lib.load().gdx_nccl_unique_id(buf))
    return bytes(buf)


def broadcast_index(index: FmIndex | None, alphabet, rank: int, world: int, device: int,
                    share_id: Callable[[bytes | None], bytes], root: int = 0) -> FmIndex:
    """Every rank calls this; the root passes its FmIndex (others None) and gets it back, the others get a
    replica on `device`.  share_id(id_or_None) must return the root's NCCL unique id on every rank (e.g. a
    torch.distributed / MPI broadcast of 128 bytes) -- the only out-of-band step; header and image travel
    through the library's own NCCL communicator."""
    uid = share_id(nccl_unique_id() if rank == root else None)
    assert len(uid) == _lib.GDX_NCCL_UNIQUE_ID_BYTES
    h = C.c_void_p()
    idbuf = (C.c_uint8 * len(uid)).from_buffer_copy(uid)
    _check(_lib.load().gdx_index_broadcast(index.handle if index is not None else None, idbuf, rank, world, root,
                                            device, C.byref(h)))
    if rank == root:
        return index
    return FmIndex(h, alphabet)


def torch_share_id(rank: int, root: int = 0, group=None) -> Callable[[bytes | None], bytes]:
    """share_id for broadcast_index on top of an initialised torch.distributed process group."""
    def share(uid):
        import torch.distributed as dist
        box = [uid]
        dist.broadcast_object_list(box, src=root, group=group)
        return box[0]
    return share


class ReplicaSet:
    """The replicas this process drives + which shards of a batch they own.  The *_packed methods take ONE batch
    (all of it) and fill the owned ranges of the result arrays in place, in input order."""

    def __init__(self, replicas: Sequence[FmIndex], first_shard: int = 0, n_shards: int | None = None):
        assert len(replicas) > 0
        self.replicas = list(replicas)
        self.first_shard = first_shard
        self.n_shards = len(self.replicas) if n_shards is None else n_shards
        self._lib = _lib.load()
        self._handles = (C.c_void_p * len(self.replicas))(*[r.handle for r in self.replicas])

    @classmethod
    def replicate(cls, index: FmIndex, devices: Sequence[int]) -> "ReplicaSet":
        """`index` + one new replica on every device of `devices` (gdx_index_replicate: one grouped ncclBroadcast)."""
        lib = _lib.load()
        n = len(devices)
        out = (C.c_void_p * max(n, 1))()
        devs = (C.c_int32 * max(n, 1))(*devices)
        _check(lib.gdx_index_replicate(index.handle, devs, n, out))
        return cls([index] + [FmIndex(C.c_void_p(out[i]), index.alphabet()) for i in range(n)])

    def owned_range(self, nq: int) -> tuple[int, int]:
        return (shard_range(nq, self.first_shard, self.n_shards)[0],
                shard_range(nq, self.first_shard + len(self.replicas) - 1, self.n_shards)[1])

    def count_many_packed(self, data, offsets=None, fixed_len: int = 0, nq: int | None = None, out=None,
                          encoding: int = _lib.GDX_QUERIES_IO_BYTES) -> np.ndarray:
        nq = (offsets.size - 1) if offsets is not None else nq
        counts = out if out is not None else np.zeros(max(nq, 1), dtype=np.uint64)
        q = _queries_struct(data, offsets, fixed_len, nq, encoding)
        _check(self._lib.gdx_count_many_sharded(self._handles, len(self.replicas), self.first_shard, self.n_shards,
                                                C.byref(q), counts.ctypes.data))
        return counts[:nq]

    def cursors_many_packed(self, data, offsets=None, fixed_len: int = 0, nq: int | None = None, out=None,
                            encoding: int = _lib.GDX_QUERIES_IO_BYTES):
        nq = (offsets.size - 1) if offsets is not None else nq
        starts, ends = out if out is not None else (np.zeros(max(nq, 1), dtype=np.uint64),
                                                    np.zeros(max(nq, 1), dtype=np.uint64))
        q = _queries_struct(data, offsets, fixed_len, nq, encoding)
        _check(self._lib.gdx_cursors_many_sharded(self._handles, len(self.replicas), self.first_shard, self.n_shards,
                                                  C.byref(q), starts.ctypes.data, ends.ctypes.data))
        return starts[:nq], ends[:nq]

    def locate_many_view(self, data, offsets=None, fixed_len: int = 0, nq: int | None = None, hit_offsets=None,
                         encoding: int = _lib.GDX_QUERIES_IO_BYTES):
        """-> (hit_offsets, [hits view per local shard], shard_first_hit, release).  The hits of query i of local
        shard k are views[k][hit_offsets[i] - shard_first_hit[k] : hit_offsets[i + 1] - shard_first_hit[k]]."""
        nq = (offsets.size - 1) if offsets is not None else nq
        if hit_offsets is None:
            hit_offsets = np.zeros(nq + 1, dtype=np.uint64)
        n_local = len(self.replicas)
        hp = (C.c_void_p * n_local)()
        first = (C.c_uint64 * (n_local + 1))()
        q = _queries_struct(data, offsets, fixed_len, nq, encoding)
        _check(self._lib.gdx_locate_many_sharded(self._handles, n_local, self.first_shard, self.n_shards, C.byref(q),
                                                 hit_offsets.ctypes.data, hp, first))
        views = []
        for k in range(n_local):
            n = int(first[k + 1] - first[k])
            if n:
                buf = (C.c_uint64 * (2 * n)).from_address(hp[k])
                views.append(np.frombuffer(buf, dtype=np.uint64).reshape(n, 2))
            else:
                views.append(np.zeros((0, 2), dtype=np.uint64))
        ptrs = [hp[k] for k in range(n_local)]

        def release():
            for r, ptr in zip(self.replicas, ptrs):
                if ptr:
                    self._lib.gdx_free_hits(r.handle, C.c_void_p(ptr))
        return hit_offsets, views, np.array(list(first), dtype=np.uint64), release

    def locate_many_compact_view(self, data, offsets=None, fixed_len: int = 0, nq: int | None = None, hit_counts=None,
                                 encoding: int = _lib.GDX_QUERIES_IO_BYTES):
        """gdx_locate_many_sharded_compact -> (hit_counts uint32[nq] (owned ranges filled), [uint32 hits view per local
        shard], release)."""
        nq = (offsets.size - 1) if offsets is not None else nq
        if hit_counts is None:
            hit_counts = np.zeros(max(nq, 1), dtype=np.uint32)
        n_local = len(self.replicas)
        hp = (C.c_void_p * n_local)()
        nh = (C.c_uint64 * n_local)()
        q = _queries_struct(data, offsets, fixed_len, nq, encoding)
        _check(self._lib.gdx_locate_many_sharded_compact(self._handles, n_local, self.first_shard, self.n_shards, C.byref(q),
                                                         hit_counts.ctypes.data, hp, nh))
        views = []
        for k in range(n_local):
            n = int(nh[k])
            if n:
                buf = (C.c_uint32 * (2 * n)).from_address(hp[k])
                views.append(np.frombuffer(buf, dtype=np.uint32).reshape(n, 2))
            else:
                views.append(np.zeros((0, 2), dtype=np.uint32))
        ptrs = [hp[k] for k in range(n_local)]

        def release():
            for r, ptr in zip(self.replicas, ptrs):
                if ptr:
                    self._lib.gdx_free_hits(r.handle, C.c_void_p(ptr))
        return hit_counts[:nq], views, release

    def locate_many_packed(self, data, offsets=None, fixed_len: int = 0, nq: int | None = None,
                           encoding: int = _lib.GDX_QUERIES_IO_BYTES):
        """Concatenating convenience form (copies): -> (hit_offsets of the owned range rebased to 0, hits[n, 2])."""
        hit_offsets, views, first, release = self.locate_many_view(data, offsets, fixed_len, nq, None, encoding)
        try:
            hits = np.concatenate(views) if views else np.zeros((0, 2), dtype=np.uint64)
        finally:
            release()
        return hit_offsets, hits

    def stats(self) -> _lib.gdx_stats:
        return self.replicas[0].stats()


def gather_in_order(local: np.ndarray, n_total: int, rank: int, world: int, group=None) -> np.ndarray | None:
    """Collect per-rank result slices (split by shard_range) on rank 0 in input order (torch.distributed)."""
    import torch
    import torch.distributed as dist
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(e - b for b, e in sizes)
    t = torch.zeros(pad, dtype=torch.int64)
    t[: local.size] = torch.from_numpy(local.astype(np.int64))
    out = [torch.zeros(pad, dtype=torch.int64) for _ in range(world)] if rank == 0 else None
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
        out = [o.cuda() for o in out] if out is not None else None
    dist.gather(t, out, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate([o.cpu().numpy()[: e - b] for o, (b, e) in zip(out, sizes)]).astype(local.dtype)
