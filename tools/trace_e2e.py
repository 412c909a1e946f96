"""Per-chunk timeline of gdx_count_many / gdx_locate_many (GDX_TRACE=1), pinned buffers.  TRACE_QUERIES (default
30 M) queries; TRACE_LOCATE=0 skips the locate calls."""
import os, sys, time
os.environ["GDX_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import genedex_b200 as gdx
args = bench.parse_args()
nq = min(args.queries, int(os.environ.get("TRACE_QUERIES", 30_000_000)))
dev = torch.device("cuda", 0)
text = bench.make_text_on_device(args.text_len, args.n_fraction, dev)
qn = torch.empty(nq * args.query_len, dtype=torch.uint8).pin_memory().numpy()
bench.fill_query_range(text, qn, np.zeros(nq, dtype=np.int64), 0, nq, args.query_len, dev)
text_host = text.cpu().numpy(); del text; torch.cuda.empty_cache()
idx = gdx.FmIndexConfig("u32").construct_on_device(True).construct_index_packed(text_host, np.array([0, text_host.size], dtype=np.uint64), gdx.alphabet.ascii_dna_with_n())
cn = torch.zeros(nq, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
for i in range(3):
    t0 = time.perf_counter()
    idx.count_many_packed(qn, None, args.query_len, nq, out=cn)
    print("call", i, "ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
if os.environ.get("TRACE_LOCATE", "1") == "0":
    sys.exit(0)
print("---- locate", file=sys.stderr)
hit_off = np.zeros(nq + 1, dtype=np.uint64)
for i in range(3):
    t0 = time.perf_counter()
    _, hits, release = idx.locate_many_view(qn, None, args.query_len, nq, hit_offsets=hit_off)
    t1 = time.perf_counter()
    release()
    st = idx.stats()
    print("locate call", i, "ms", (t1 - t0) * 1e3, "kernel_ms_locate", st.kernel_ms_locate, "search", st.kernel_ms_search, file=sys.stderr)
