"""Config C4 (20-letter protein alphabet, 500 M residues, 10 M length-12 queries) count launches on their own, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import genedex_b200 as gdx
import run_configs as rc
dev = torch.device("cuda", 0)
n, nq, m = 500_000_000, 10_000_000, 12
text = rc.make_protein_text(n, dev, 0x5EED0004)
host = text.cpu().numpy()
offs = np.array([0, n], dtype=np.int64)
q1, _, _ = rc.sample_windows(text, offs, nq // 2, m, 0x5EED0005, dev)
q = torch.cat([q1, rc.random_queries(nq - nq // 2, m, rc.PROTEIN, 0x5EED0006, dev)])
del text
idx = gdx.FmIndexConfig("u32").construct_on_device(True).construct_index_packed(host, offs.astype(np.uint64), gdx.Alphabet.from_io_symbols(rc.PROTEIN, 0))
torch.cuda.nvtx.range_push("timed")
ms, _ = rc.time_device_count(gdx, idx, q, m, nq, reps=3)
torch.cuda.nvtx.range_pop()
print("c4 count kernel ms", ms, file=sys.stderr)
