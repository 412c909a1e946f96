"""Multi-GPU plumbing around the C ABI: one process per GPU, full index replica per GPU, queries
sharded by contiguous ranges, no collective on the query path.

Only torch.distributed plumbing lives here (broadcast of the device image from rank 0, gather of
per-rank results); the search itself is always the C ABI on the local GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class CudaBytes:
    """Zero-copy torch view of a raw device allocation: torch.as_tensor(CudaBytes(ptr, n), device=...)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced, order-preserving split of [0, n_items) over `world` ranks."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_bytes(buf, src: int = 0, group=None, chunk: int = 1 << 30):
    """Broadcast a 1-D uint8 torch tensor in chunks (NCCL over NVLink for cuda tensors, gloo on CPU)."""
    import torch.distributed as dist
    n = buf.numel()
    for b in range(0, n, chunk):
        dist.broadcast(buf[b:min(n, b + chunk)], src, group=group)
    return buf


def replicate_index(index, alphabet, device, rank: int, group=None):
    """Rank 0 passes its FmIndex (others pass None); every rank returns an FmIndex on `device`.

    Rank 0 exports (header, image pointer); the image goes out with one chunked NCCL broadcast; the
    peers adopt header + bytes without copying (the tensor is kept alive by the returned index)."""
    import torch
    import torch.distributed as dist

    from .index import FmIndex
    lib = _lib.load()
    hbytes = int(lib.gdx_index_header_bytes())
    hdr = torch.zeros(hbytes, dtype=torch.uint8, device=device)
    size = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == 0:
        hbuf = (C.c_uint8 * hbytes)()
        img, nbytes = C.c_void_p(), C.c_uint64()
        rc = lib.gdx_index_export(index.handle, hbuf, C.byref(img), C.byref(nbytes))
        assert rc == 0, lib.gdx_last_error_message()
        hdr.copy_(torch.frombuffer(bytearray(hbuf), dtype=torch.uint8))
        size[0] = nbytes.value
    dist.broadcast(hdr, 0, group=group)
    dist.broadcast(size, 0, group=group)
    n = int(size.item())
    if rank == 0:
        image = torch.as_tensor(CudaBytes(img.value, n), device=device)
    else:
        image = torch.empty(n, dtype=torch.uint8, device=device)
    broadcast_bytes(image, 0, group)
    if rank == 0:
        return index
    torch.cuda.synchronize(device)
    hb = (C.c_uint8 * hbytes).from_buffer_copy(hdr.cpu().numpy().tobytes())
    h = C.c_void_p()
    rc = lib.gdx_index_adopt_image(hb, image.data_ptr(), device.index, 0, C.byref(h))
    assert rc == 0, lib.gdx_last_error_message()
    return FmIndex(h, alphabet, keepalive=image)


def gather_in_order(local: np.ndarray, n_total: int, rank: int, world: int, group=None) -> np.ndarray | None:
    """Collect per-rank result slices (split by shard_range) on rank 0 in input order."""
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
