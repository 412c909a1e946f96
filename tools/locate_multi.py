"""The multi-hit locate workload of bench.py on its own (for ncu): 2 M length-14 queries on the 3.1 Gbp index,
~12 hits each, in ONE pipeline chunk (GDX_CHUNK_FIRST_MB=64), dense suffix array (argv[1] = dense) or the configured
sampled one (argv[1] = sampled; k_locate_walk_compact)."""
import os, sys, time
os.environ.setdefault("GDX_CHUNK_FIRST_MB", "64")
os.environ.setdefault("GDX_CHUNK_MAX_MB", "64")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import genedex_b200 as gdx
mode = sys.argv[1] if len(sys.argv) > 1 else "sampled"

dev = torch.device("cuda", 0)
n, mq, mm = 3_100_000_000, 2_000_000, 14
text = bench.make_text_on_device(n, 0.05, dev)
q = np.empty(mq * mm, dtype=np.uint8)
bench.fill_query_range(text, q, None, 0, mq, mm, dev, seed=bench.QUERY_SEED + 7)
host = text.cpu().numpy(); del text; torch.cuda.empty_cache()
idx = gdx.FmIndexConfig("u32").construct_on_device(True).construct_index_packed(host, np.array([0, host.size], dtype=np.uint64), gdx.alphabet.ascii_dna_with_n())
if mode == "sampled":
    idx.set_dense_suffix_array(False)
off = np.zeros(mq + 1, dtype=np.uint64)
for it in range(3):
    torch.cuda.nvtx.range_push("timed")
    t0 = time.perf_counter()
    _, hits, rel = idx.locate_many_view(q, None, mm, mq, hit_offsets=off)
    dt = (time.perf_counter() - t0) * 1e3
    torch.cuda.nvtx.range_pop()
    s = idx.stats(); rel()
    print(mode, "call", it, "ms", round(dt, 2), "hits", int(s.hits), "kernel_ms_locate", round(s.kernel_ms_locate, 3), "walk steps", int(s.locate_walk_steps), file=sys.stderr)
