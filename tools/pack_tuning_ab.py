"""ABAB comparison of the host packer's tuning knobs inside ONE process (hosts are noisy: alternate the settings, report
medians): the packer alone on 1.5 GB, and gdx_count_many end to end on a 30 M-query batch from pinned buffers."""
import ctypes as C, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import genedex_b200 as gdx
from genedex_b200.index import _alphabet_struct

lib = gdx._lib.load()
SETTINGS = [(0, 0), (2048, 0), (0, 1), (2048, 1), (4096, 1), (4096, 0), (1024, 1)]
a = _alphabet_struct(gdx.alphabet.ascii_dna_with_n())
n = 1_500_000_000
rng = np.random.default_rng(1)
data = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)].copy()
out = np.zeros(n // 4 + 64, dtype=np.uint8)
ne = C.c_uint64()
res = {s: [] for s in SETTINGS}
for rnd in range(6):
    for s in SETTINGS:
        lib.gdx_host_pack_tuning(*s)
        t = time.perf_counter()
        lib.gdx_pack_symbols(C.byref(a), data.ctypes.data, n, out.ctypes.data, None, 0, C.byref(ne))
        res[s].append(n / (time.perf_counter() - t) / 1e9)
print("packer alone, GB/s of IO bytes (median of 6, alternating):")
for s in SETTINGS:
    print(f"  prefetch {s[0]:5d} stream {s[1]}: {statistics.median(res[s]):6.1f}  (min {min(res[s]):.1f} max {max(res[s]):.1f})", flush=True)
del data, out
dev = torch.device("cuda", 0)
nq, m = 30_000_000, 50
text = bench.make_text_on_device(3_100_000_000, 0.05, dev)
qn = torch.empty(nq * m, dtype=torch.uint8).pin_memory().numpy()
bench.fill_query_range(text, qn, np.zeros(nq, dtype=np.int64), 0, nq, m, dev)
host = text.cpu().numpy(); del text; torch.cuda.empty_cache()
idx = gdx.FmIndexConfig("u32").construct_on_device(True).construct_index_packed(host, np.array([0, host.size], dtype=np.uint64), gdx.alphabet.ascii_dna_with_n())
cn = torch.zeros(nq, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
for _ in range(3):
    idx.count_many_packed(qn, None, m, nq, out=cn)
res = {s: [] for s in SETTINGS}
for rnd in range(8):
    for s in SETTINGS:
        lib.gdx_host_pack_tuning(*s)
        t = time.perf_counter()
        idx.count_many_packed(qn, None, m, nq, out=cn)
        res[s].append((time.perf_counter() - t) * 1e3)
print("gdx_count_many end to end, 30 M queries, ms (median of 8, alternating):")
for s in SETTINGS:
    print(f"  prefetch {s[0]:5d} stream {s[1]}: {statistics.median(res[s]):6.2f} ms = {nq / statistics.median(res[s]) / 1e6:.2f} G q/s  (min {min(res[s]):.2f} max {max(res[s]):.2f})", flush=True)
