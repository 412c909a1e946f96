"""A/B of the pipeline chunk size against the packer's store kind, end to end (gdx_count_many, pinned buffers, 60 M
length-50 queries on the 3.1 Gbp index).  Question: do small chunks whose packed staging stays in the host's L3
(regular stores, DMA reads served by the cache) beat 32 MB chunks written around the cache?

  python tools/chunk_ab.py            # parent: one child process per setting (the chunk size is read at load time)
"""
import json, os, subprocess, sys, time

SETTINGS = [(32, 1), (32, 0), (8, 1), (8, 0), (4, 0), (16, 0)]   # (GDX_CHUNK_MAX_MB, GDX_PACK_STREAM)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.argv = sys.argv[:1]
    import numpy as np, torch
    import bench
    import genedex_b200 as gdx
    args = bench.parse_args()
    nq = args.queries
    dev = torch.device("cuda", 0)
    text = bench.make_text_on_device(args.text_len, args.n_fraction, dev)
    qn = torch.empty(nq * args.query_len, dtype=torch.uint8).pin_memory().numpy()
    bench.fill_query_range(text, qn, np.zeros(nq, dtype=np.int64), 0, nq, args.query_len, dev)
    text_host = text.cpu().numpy(); del text; torch.cuda.empty_cache()
    idx = gdx.FmIndexConfig("u32").construct_on_device(True).construct_index_packed(
        text_host, np.array([0, text_host.size], dtype=np.uint64), gdx.alphabet.ascii_dna_with_n())
    cn = torch.zeros(nq, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
    ms = []
    for i in range(11):
        t0 = time.perf_counter()
        idx.count_many_packed(qn, None, args.query_len, nq, out=cn)
        ms.append((time.perf_counter() - t0) * 1e3)
    ms = sorted(ms[3:])
    st = idx.stats()
    print(json.dumps({"chunk_max_mb": int(os.environ["GDX_CHUNK_MAX_MB"]), "stream": int(os.environ["GDX_PACK_STREAM"]),
                      "ms_median": ms[len(ms) // 2], "ms_min": ms[0], "ms_max": ms[-1], "gqps": nq / ms[len(ms) // 2] / 1e6,
                      "packed_queries": int(st.packed_queries), "checksum": int(cn.sum())}))
    sys.exit(0)

for chunk, stream in SETTINGS:
    env = dict(os.environ, GDX_CHUNK_MAX_MB=str(chunk), GDX_PACK_STREAM=str(stream))
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=240)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    print(line[-1] if line else f"FAILED chunk {chunk} stream {stream}: {r.stderr[-400:]}", flush=True)
