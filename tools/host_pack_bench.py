"""Host packer throughput by thread count (gdx_pack_symbols = the stage of gdx_count_many that packs IO bytes to
2 bits) next to a plain parallel memcpy of the same bytes, to tell a compute-bound packer from a memory-bound box.

  python tools/host_pack_bench.py [bytes]
"""
import ctypes as C
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import genedex_b200 as gdx  # noqa: E402
from genedex_b200.index import _alphabet_struct  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000_000
    lib = gdx._lib.load()
    print(subprocess.run("lscpu | egrep 'Model name|Socket|Core|Thread|^CPU\\(s\\)|NUMA|L3|Flags' | cut -c1-400", shell=True,
                         capture_output=True, text=True).stdout)
    a = _alphabet_struct(gdx.alphabet.ascii_dna_with_n())
    rng = np.random.default_rng(1)
    data = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)].copy()
    out = np.zeros(n // 4 + 8, dtype=np.uint8)
    ne = C.c_uint64()
    ncpu = len(os.sched_getaffinity(0))
    for threads in [1, 2, 4, 8, 12, 16, 24, 32]:
        if threads > ncpu:
            break
        lib.gdx_host_pool_resize(threads)
        best = 1e9
        for _ in range(4):
            t = time.perf_counter()
            lib.gdx_pack_symbols(C.byref(a), data.ctypes.data, n, out.ctypes.data, None, 0, C.byref(ne))
            best = min(best, time.perf_counter() - t)
        # plain copy of the same source bytes with the same number of python threads (numpy releases the GIL)
        dst = np.empty(n, dtype=np.uint8)
        parts = np.array_split(np.arange(0, n + 1, max(1, n // threads))[: threads + 1], 1)[0]
        bounds = list(parts[:-1]) + [n]
        cbest = 1e9
        for _ in range(3):
            ths = [threading.Thread(target=lambda lo, hi: np.copyto(dst[lo:hi], data[lo:hi]), args=(int(bounds[i]), int(bounds[i + 1])))
                   for i in range(len(bounds) - 1)]
            t = time.perf_counter()
            [x.start() for x in ths]
            [x.join() for x in ths]
            cbest = min(cbest, time.perf_counter() - t)
        del dst
        print(f"threads {threads:2d}: pack {n / best / 1e9:6.1f} GB/s of IO bytes ({best * 1e3:7.1f} ms)   "
              f"memcpy {n / cbest / 1e9:6.1f} GB/s read + the same written", flush=True)


def latency(n=134_217_728, iters=300):
    """distribution of the duration of one chunk-sized packing job (the pipeline waits for every one of them)"""
    lib = gdx._lib.load()
    a = _alphabet_struct(gdx.alphabet.ascii_dna_with_n())
    rng = np.random.default_rng(2)
    data = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)].copy()
    out = np.zeros(n // 4 + 8, dtype=np.uint8)
    ne = C.c_uint64()
    ncpu = len(os.sched_getaffinity(0))
    for threads in sorted({ncpu, ncpu - 1, ncpu - 2, max(1, ncpu // 2)}, reverse=True):
        lib.gdx_host_pool_resize(threads)
        ts = []
        for _ in range(iters):
            t = time.perf_counter()
            lib.gdx_pack_symbols(C.byref(a), data.ctypes.data, n, out.ctypes.data, None, 0, C.byref(ne))
            ts.append((time.perf_counter() - t) * 1e3)
        ts.sort()
        print(f"chunk of {n >> 20} MiB, {threads:2d} threads: median {ts[len(ts) // 2]:.2f} ms  p90 {ts[int(0.9 * len(ts))]:.2f}  "
              f"p99 {ts[int(0.99 * len(ts))]:.2f}  max {ts[-1]:.2f} ms", flush=True)


def knobs(n=1_500_000_000):
    """A/B of the packer's prefetch / streaming-store knobs (read once per process: one subprocess per setting)"""
    code = ("import sys, time, ctypes as C, numpy as np; sys.path.insert(0, %r); import genedex_b200 as gdx\n"
            "from genedex_b200.index import _alphabet_struct\n"
            "lib = gdx._lib.load(); a = _alphabet_struct(gdx.alphabet.ascii_dna_with_n()); n = %d\n"
            "rng = np.random.default_rng(1); data = np.frombuffer(b'ACGT', dtype=np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)].copy()\n"
            "out = np.zeros(n // 4 + 64, dtype=np.uint8); ne = C.c_uint64(); best = 1e9\n"
            "for _ in range(5):\n"
            "    t = time.perf_counter(); lib.gdx_pack_symbols(C.byref(a), data.ctypes.data, n, out.ctypes.data, None, 0, C.byref(ne)); best = min(best, time.perf_counter() - t)\n"
            "print('%%.1f GB/s' %% (n / best / 1e9))" % (ROOT, n))
    for env in ({}, {"GDX_PACK_STREAM": "0"}, {"GDX_PACK_PREFETCH": "0"}, {"GDX_PACK_PREFETCH": "0", "GDX_PACK_STREAM": "0"},
                {"GDX_PACK_PREFETCH": "1024"}, {"GDX_PACK_PREFETCH": "4096"}, {"GDX_PACK_PREFETCH": "8192"},
                {"GDX_PACK_ISA": "avx2"}, {"GDX_PACK_ISA": "vbmi"},
                {"GDX_HOST_THREADS": "1"}, {"GDX_HOST_THREADS": "1", "GDX_PACK_PREFETCH": "0", "GDX_PACK_STREAM": "0"}):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
        print(env or "default", out.stdout.strip(), out.stderr.strip()[-200:], flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "knobs":
        knobs()
    elif len(sys.argv) > 1 and sys.argv[1] == "latency":
        latency()
    else:
        main()
