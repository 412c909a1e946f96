"""How much do pageable (ordinary malloc) caller buffers cost compared with pinned ones?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import genedex_b200 as gdx
args = bench.parse_args()
dev = torch.device("cuda", 0)
text = bench.make_text_on_device(args.text_len, args.n_fraction, dev)
q_dev, _ = bench.sample_queries_on_device(text, args.queries, args.query_len, bench.QUERY_SEED, dev)
q_pinned = torch.empty(q_dev.numel(), dtype=torch.uint8).pin_memory(); q_pinned.copy_(q_dev)
c_pinned = torch.empty(args.queries, dtype=torch.int64).pin_memory()
text_host = text.cpu().numpy(); del text, q_dev; torch.cuda.empty_cache()
idx = gdx.FmIndexConfig("u32").construct_on_device(True).construct_index_packed(text_host, np.array([0, text_host.size], dtype=np.uint64), gdx.alphabet.ascii_dna_with_n())
m, nq = args.query_len, args.queries
q_page = q_pinned.numpy().copy()
c_page = np.empty(nq, dtype=np.uint64)
def run(name, q, c, fn):
    for _ in range(2): fn(q, c)
    t0 = time.perf_counter()
    for _ in range(5): fn(q, c)
    print(f"{name}: {(time.perf_counter() - t0) * 200:.2f} ms per call", flush=True)
cnt = lambda q, c: idx.count_many_packed(q, None, m, nq, out=c)
run("count pinned in / pinned out", q_pinned.numpy(), c_pinned.numpy().view(np.uint64), cnt)
run("count pinned in / pageable out", q_pinned.numpy(), c_page, cnt)
run("count pageable in / pinned out", q_page, c_pinned.numpy().view(np.uint64), cnt)
run("count pageable in / pageable out", q_page, c_page, cnt)
run("cursors pinned in / pageable out (mirror default)", q_pinned.numpy(), None, lambda q, c: idx.cursors_many_packed(q, None, m, nq))
run("locate pinned in / pageable offsets (mirror default)", q_pinned.numpy(), None, lambda q, c: idx.locate_many_view(q, None, m, nq)[2]())
