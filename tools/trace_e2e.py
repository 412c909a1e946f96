"""Per-chunk timeline of gdx_count_many (GDX_TRACE=1) on a headline-sized batch."""
import os, sys, time
os.environ["GDX_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import genedex_b200 as gdx
args = bench.parse_args()
dev = torch.device("cuda", 0)
text = bench.make_text_on_device(args.text_len, args.n_fraction, dev)
q_dev, _ = bench.sample_queries_on_device(text, args.queries, args.query_len, bench.QUERY_SEED, dev)
q_host = torch.empty(q_dev.numel(), dtype=torch.uint8).pin_memory(); q_host.copy_(q_dev)
counts = torch.empty(args.queries, dtype=torch.int64).pin_memory()
text_host = text.cpu().numpy(); del text, q_dev; torch.cuda.empty_cache()
idx = gdx.FmIndexConfig("u32").construct_on_device(True).construct_index_packed(text_host, np.array([0, text_host.size], dtype=np.uint64), gdx.alphabet.ascii_dna_with_n())
qn, cn = q_host.numpy(), counts.numpy().view(np.uint64)
for i in range(4):
    t0 = time.perf_counter()
    idx.count_many_packed(qn, None, args.query_len, args.queries, out=cn)
    print("call", i, "ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
print("---- locate", file=sys.stderr)
hit_off = torch.empty(args.queries + 1, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
for i in range(4):
    t0 = time.perf_counter()
    _, hits, release = idx.locate_many_view(qn, None, args.query_len, args.queries, hit_offsets=hit_off)
    t1 = time.perf_counter()
    release()
    st = idx.stats()
    print("locate call", i, "ms", (t1 - t0) * 1e3, "kernel_ms_locate", st.kernel_ms_locate, "search", st.kernel_ms_search, file=sys.stderr)
