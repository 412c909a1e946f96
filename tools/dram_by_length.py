"""DRAM bytes and time of k_search as a function of the query length (run under ncu):

  ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:k_search \
      --csv --log-file gpurun_out/dram_by_len.csv python tools/dram_by_length.py

Each launch processes 7.5 M queries of one length sampled from the 3.1 Gbp benchmark text; the
difference between successive lengths is the DRAM cost of the LF steps at that depth of the trie.
Also prints the live CUDA-event time of each launch (meaningless under ncu)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import genedex_b200 as gdx  # noqa: E402


def main():
    import torch
    n = int(os.environ.get("TEXT_LEN", 3_100_000_000))
    nq = int(os.environ.get("QUERIES", 7_500_000))
    lens = [int(x) for x in os.environ.get("LENS", "6,8,9,10,11,12,13,14,15,16,17,18,20,24,30,40,50").split(",")]
    dev = torch.device("cuda", 0)
    text = bench.make_text_on_device(n, 0.05, dev)
    qd = {m: bench.sample_queries_on_device(text, nq, m, 7 + m, dev)[0] for m in lens}
    text_host = text.cpu().numpy()
    del text
    torch.cuda.empty_cache()
    pidx = (gdx.FmIndexConfig("u32").suffix_array_sampling_rate(4).lookup_table_depth(0).construct_on_device(True)
            .construct_index_packed(text_host, np.array([0, n], dtype=np.uint64), gdx.alphabet.ascii_dna_with_n()))
    lib = gdx._lib.load()
    d_counts = torch.zeros(nq, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    reps = int(os.environ.get("REPS", 1))
    for m in lens:
        qs = gdx._lib.gdx_queries(qd[m].data_ptr(), None, m, nq)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(reps):
            assert lib.gdx_count_many_device(pidx.handle, C.byref(qs), d_counts.data_ptr(), None, stream) == 0
        torch.cuda.synchronize()
        e0.record()
        assert lib.gdx_count_many_device(pidx.handle, C.byref(qs), d_counts.data_ptr(), None, stream) == 0
        e1.record()
        torch.cuda.synchronize()
        print(f"m={m} ms={e0.elapsed_time(e1):.3f} mean_count={float(d_counts.double().mean()):.2f}", flush=True)


if __name__ == "__main__":
    main()
