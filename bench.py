#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched FM-index search path (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # CPU reference arm (oracle port)

Workload at N=1 (BASELINE.json configs[1], SURVEY 8d "C2"): hg38-shaped synthetic 3.1 Gbp DNA single
text (`ascii_dna_with_n`, ~5 % N in runs <= 10 kbp, u32 storage, sampling rate 4, lookup depth 0 =
crate defaults), 7.5 M length-50 queries sampled from the text (375 MB), count.  One "step" = one
pass of count_many over the whole query batch.  For N > 1 every rank holds a full index replica
(built on rank 0, one NCCL broadcast of the device image) and its own 7.5 M queries: weak scaling,
no collective on the query path.

`value`   queries/s with the queries already resident in HBM (k_query_keys + radix sort + one k_search
          launch per step).  The index carries the library's default accelerators (packed text, dense
          suffix array: config.dense_suffix_array_bytes); GDX_DENSE_SA=0 / GDX_VERIFY=0 switch them off.
`e2e`     the same through the C ABI with pinned HOST buffers: H2D of the queries and D2H of the
          counts are inside the timed region (chunked 3-stream pipeline in libgenedex_b200).
`roofline` the k_search launch: algorithmic bytes (m + 2*R*steps + 16 per query for the LF steps
          executed, + R per walk step + 64 per text-verified query; R = 32 B rank record) / CUDA-event
          duration of the whole step, against the measured HBM copy peak; `traffic` = DRAM bytes of the
          launch from the committed ncu capture (profiles/r1_k_search*.txt).
`cpu_baseline` the oracle's port of the reference's 64-query batched search on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("GDX_ORACLE_NATIVE", "1")  # CPU baseline: oracle rebuilt with -march=native on this box

TEXT_SEED = 0x5EED0001
QUERY_SEED = 0x5EED0002
N_CODE = ord("N")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--text-len", type=int, default=3_100_000_000)
    ap.add_argument("--queries", type=int, default=7_500_000)
    ap.add_argument("--query-len", type=int, default=50)
    ap.add_argument("--lookup-depth", type=int, default=0)
    ap.add_argument("--sampling-rate", type=int, default=4)
    ap.add_argument("--n-fraction", type=float, default=0.05)
    ap.add_argument("--no-locate", action="store_true", help="skip the locate side measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline sample")
    return ap.parse_args()


# ---- synthetic data (torch on the GPU is plumbing here: RNG + gathers, nothing of the search path) ---
def make_text_on_device(n_symbols, n_fraction, device):
    """hg38-shaped DNA: iid ACGT with ~n_fraction N in runs of 1..10000; returns uint8 IO bytes."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(TEXT_SEED)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    text = lut[torch.randint(0, 4, (n_symbols,), generator=g, device=device, dtype=torch.uint8).long()] \
        if n_symbols <= (1 << 28) else None
    if text is None:  # chunked to keep the int64 index temporaries small
        text = torch.empty(n_symbols, dtype=torch.uint8, device=device)
        step = 1 << 28
        for b in range(0, n_symbols, step):
            e = min(n_symbols, b + step)
            text[b:e] = lut[torch.randint(0, 4, (e - b,), generator=g, device=device, dtype=torch.uint8).long()]
    max_run = min(10_000, max(1, n_symbols // 8))
    n_runs = int(n_fraction * n_symbols / (max_run / 2)) if n_fraction > 0 else 0
    if n_runs:
        cg = torch.Generator()
        cg.manual_seed(TEXT_SEED + 1)
        starts = torch.randint(0, n_symbols - max_run, (n_runs,), generator=cg).tolist()
        lens = torch.randint(1, max_run + 1, (n_runs,), generator=cg).tolist()
        for s, l in zip(starts, lens):
            text[s:s + l] = N_CODE
    return text


def sample_queries_on_device(text, nq, m, seed, device):
    """nq windows of length m sampled uniformly from the text, windows containing N rejected."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = text.numel()
    out = torch.empty((nq, m), dtype=torch.uint8, device=device)
    starts_out = torch.empty(nq, dtype=torch.int64, device=device)
    filled = 0
    ar = torch.arange(m, device=device)
    while filled < nq:
        want = min(nq - filled, 1 << 21)
        cand = torch.randint(0, n - m, (int(want * 1.25) + 16,), generator=g, device=device)
        win = text[cand[:, None] + ar[None, :]]
        ok = ~(win == N_CODE).any(dim=1)
        win, cand = win[ok][:want], cand[ok][:want]
        k = win.shape[0]
        out[filled:filled + k] = win
        starts_out[filled:filled + k] = cand
        filled += k
    return out.reshape(-1), starts_out


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML every 5 ms in a thread; falls
    back to `nvidia-smi -lms` when pynvml is unavailable)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.samples = []  # (time, sm_mhz, reasons bitmask)
        self.gpu_index = gpu_index
        self.stop_flag = False
        self.thread = None
        self.sm_max = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.gpu_index]) if vis and vis.split(",")[0].isdigit() else self.gpu_index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def pump():
                while not self.stop_flag:
                    try:
                        mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        try:
                            reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((time.time(), float(mhz), int(reasons)))
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        sel = [(m, r) for t, m, r in self.samples if t0 <= t <= t1]
        reasons = sorted(name for name, bit in self.BAD.items() if any(r & bit for _, r in sel))
        return {"sm_mhz": statistics.median([m for m, _ in sel]) if sel else None, "sm_max_mhz": self.sm_max,
                "reasons": reasons, "samples": len(sel)}


def ncu_traffic_bytes(args, verified, dense, seeded):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_search launch, read from the committed summary of
    the ncu --set full capture of this very workload (profiles/r1_k_search*.txt); None for any other workload."""
    default = (args.text_len == 3_100_000_000 and args.queries == 7_500_000 and args.query_len == 50
               and args.lookup_depth == 0 and args.sampling_rate == 4)
    if not default:
        return None
    name = {(True, True, True): "r1_k_search.txt", (True, True, False): "r1_k_search_dense_sa.txt",
            (True, False, False): "r1_k_search_sampled_sa.txt",
            (False, False, False): "r1_k_search_v1_lf_only_sorted.txt"}.get((bool(verified), bool(dense), bool(seeded)))
    if name is None:
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total = 0.0
    try:
        for line in open(os.path.join(ROOT, "profiles", name)):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(f[1]) * unit[f[2]]
    except Exception:
        return None
    return total or None


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads (and with them its pinned staging memory) to the NUMA node its GPU
    hangs off: with 8 ranks streaming 375 MB batches concurrently, remote-socket host memory halves the
    per-GPU PCIe rate.  Best effort; returns the node or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b2 = part.partition("-")
            cpus.update(range(int(a), int(b2 or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def workload_name(args):
    return (f"C2: hg38-shaped synthetic {args.text_len / 1e9:.2f} Gbp DNA single text (ascii_dna_with_n, "
            f"{args.n_fraction:.0%} N in runs<=10kbp, u32, s={args.sampling_rate}, D={args.lookup_depth}), "
            f"{args.queries / 1e6:.2f}M len-{args.query_len} queries sampled from the text, count")



def build_oracle_from_product(pidx, args, nthreads=0):
    """CPU oracle index (reference three-array layout) over the BWT of the device-built index."""
    from oracle import oracle as O
    bwt = pidx.download_bwt()
    count = pidx.count_array()
    n = pidx.total_text_len()
    oa = O.ALPHABETS["ascii_dna_with_n"]()
    return O.OracleIndex.from_parts(bwt, oa, count, np.array([n - 1], dtype=np.uint64), None, args.sampling_rate,
                                    None, None, lookup_depth=args.lookup_depth, storage="u32", nthreads=nthreads)


# ---- CPU reference arm -----------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch

    import genedex_b200 as gdx
    from oracle import oracle as O
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    text = make_text_on_device(args.text_len, args.n_fraction, dev)
    q_dev, _ = sample_queries_on_device(text, args.queries, args.query_len, QUERY_SEED, dev)
    q = q_dev.cpu().numpy()
    text_host = text.cpu().numpy()
    del text, q_dev
    torch.cuda.empty_cache()
    cfg = (gdx.FmIndexConfig("u32").suffix_array_sampling_rate(args.sampling_rate)
           .lookup_table_depth(args.lookup_depth).construct_on_device(True))
    pidx = cfg.construct_index_packed(text_host, np.array([0, text_host.size], dtype=np.uint64),
                                      gdx.alphabet.ascii_dna_with_n())
    del text_host
    cores = O.lib().gdxo_online_cores()
    oidx = build_oracle_from_product(pidx, args, nthreads=cores)
    del pidx
    m, nq = args.query_len, args.queries
    # calibrate, then size each step so that the whole run stays within ~2 minutes
    cal = min(nq, 100_000)
    off = np.arange(cal + 1, dtype=np.uint64) * m
    t0 = time.perf_counter()
    oidx.count_many_packed(q[: cal * m], off, nthreads=cores)
    rate = cal / (time.perf_counter() - t0)
    per_step = int(min(nq, max(64 * cores, rate * 120.0 / (args.steps + args.warmup))))
    off = np.arange(per_step + 1, dtype=np.uint64) * m
    times = []
    for it in range(args.warmup + args.steps):
        lo = (it * per_step) % max(1, nq - per_step + 1)
        t0 = time.perf_counter()
        oidx.count_many_packed(q[lo * m:(lo + per_step) * m], off, nthreads=cores)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = per_step / (ms * 1e-3)
    sample = f"{per_step} of the {nq} queries per step, {cores} threads, contiguous chunk per thread"
    emit_json({
        "impl": "reference", "metric": "len-50 count queries/s on 3.1 Gbp DNA index", "value": value,
        "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "index_built_by": "device construction (setup, untimed)"},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ---- this repo's arm ---------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import genedex_b200 as gdx
    lib = gdx._lib.load()
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(dev)
    m, nq = args.query_len, args.queries

    t_setup = time.perf_counter()
    text = make_text_on_device(args.text_len, args.n_fraction, dev)
    q_dev, q_starts = sample_queries_on_device(text, nq, m, QUERY_SEED + rank, dev)
    q_host = torch.empty(nq * m, dtype=torch.uint8).pin_memory()
    q_host.copy_(q_dev)
    counts_host = torch.empty(nq, dtype=torch.int64).pin_memory()
    starts_host = q_starts.cpu().numpy()
    text_host = text.cpu().numpy() if rank == 0 else None
    del text, q_dev, q_starts
    torch.cuda.empty_cache()
    t_data = time.perf_counter() - t_setup

    # index: built on rank 0 (device construction), one NCCL broadcast of the image to the replicas
    t0 = time.perf_counter()
    alphabet = gdx.alphabet.ascii_dna_with_n()
    if rank == 0:
        cfg = (gdx.FmIndexConfig("u32").suffix_array_sampling_rate(args.sampling_rate)
               .lookup_table_depth(args.lookup_depth).device(local_rank)
               .construct_on_device(True, verify=True))
        pidx = cfg.construct_index_packed(text_host, np.array([0, text_host.size], dtype=np.uint64), alphabet)
        del text_host
    t_build = time.perf_counter() - t0
    t_bcast = 0.0
    if world > 1:  # one NCCL broadcast of the device image GPU0 -> peers (genedex_b200/replicate.py)
        from genedex_b200.replicate import replicate_index
        t0 = time.perf_counter()
        pidx = replicate_index(pidx if rank == 0 else None, alphabet, dev, rank)
        torch.cuda.synchronize()
        t_bcast = time.perf_counter() - t0
    info = pidx.info()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: queries resident in HBM, one k_search launch per step --------------------------------
    d_q = q_host.to(dev, non_blocking=False)
    d_counts = torch.zeros(nq, dtype=torch.int64, device=dev)
    d_err = torch.full((1,), -1, dtype=torch.int64, device=dev)
    qs = gdx._lib.gdx_queries(d_q.data_ptr(), None, m, nq)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        rc = lib.gdx_count_many_device(pidx.handle, C.byref(qs), d_counts.data_ptr(), d_err.data_ptr(), stream)
        assert rc == 0, lib.gdx_last_error_message()

    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_region0 = time.time()
    evs[0].record()
    for i in range(args.steps):
        step_device()
        evs[i + 1].record()
    barrier()
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    kernel_ms = max_over_ranks(evs[0].elapsed_time(evs[-1]) / args.steps)
    assert int(d_err.item()) == -1
    counts_device_path = d_counts.cpu().numpy().astype(np.uint64)

    # ---- e2e: pinned host buffers through gdx_count_many (H2D + kernels + D2H inside the call) -------
    q_np, counts_np = q_host.numpy(), counts_host.numpy().view(np.uint64)
    for _ in range(max(args.warmup, 3)):
        pidx.count_many_packed(q_np, None, m, nq, out=counts_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pidx.count_many_packed(q_np, None, m, nq, out=counts_np)  # synchronous: returns with the counts on the host
    e2e_ms_local = (time.perf_counter() - t0) * 1e3 / args.steps  # this rank alone, before the barrier
    barrier()
    e2e_ms_total = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ms = max_over_ranks(e2e_ms_total)
    e2e_ms_ranks = [e2e_ms_local]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (e2e_ms_local, numa_node))
        e2e_ms_ranks = gathered
    t_region1 = time.time()
    st = pidx.stats()
    assert np.array_equal(counts_np, counts_device_path), "device-resident and host-buffer paths disagree"
    assert int(counts_np.min()) >= 1, "a query sampled from the text must occur at least once"
    clock_info = clocks.stop(t_region0, t_region1)

    # ---- locate (configs[2] side measurement): LF-walk + text-id mapping through the C ABI -----------
    locate = None
    if not args.no_locate:
        hit_off = torch.empty(nq + 1, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
        for _ in range(2):
            _, hits, release = pidx.locate_many_view(q_np, None, m, nq, hit_offsets=hit_off)
            release()
        barrier()
        t0 = time.perf_counter()
        reps = max(1, min(args.steps, 5))
        for _ in range(reps):
            _, hits, release = pidx.locate_many_view(q_np, None, m, nq, hit_offsets=hit_off)
            release()  # the pooled pinned buffer stays valid until the next locate call
        barrier()
        loc_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / reps)
        lst = pidx.stats()
        # size-independent property: every query sampled at position p is located at p
        first = hits[hit_off[:-1].astype(np.int64), 1].astype(np.int64)
        single = (hit_off[1:] - hit_off[:-1]) == 1
        assert np.array_equal(first[single], starts_host[single]), "locate: a unique hit is not at its origin"
        assert np.array_equal((hit_off[1:] - hit_off[:-1]).astype(np.uint64), counts_np)
        locate = {"value": world * nq / (loc_ms * 1e-3), "unit": "queries/s (e2e, host buffers)",
                  "hits_per_step": int(lst.hits), "walk_steps": int(lst.walk_steps),
                  "kernel_ms_locate": lst.kernel_ms_locate, "kernel_ms_search": lst.kernel_ms_search,
                  "ms_per_step": loc_ms}

    # ---- CPU baseline (rank 0, N = 1 only) + parity of the GPU counts on the same sample --------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = O.lib().gdxo_online_cores()
        oidx = build_oracle_from_product(pidx, args, nthreads=cores)
        cal = min(nq, 100_000)
        off = np.arange(cal + 1, dtype=np.uint64) * m
        t0 = time.perf_counter()
        oidx.count_many_packed(q_np[: cal * m], off, nthreads=cores)
        rate = cal / (time.perf_counter() - t0)
        sample_n = int(min(nq, max(cal, rate * args.cpu_seconds)))
        off = np.arange(sample_n + 1, dtype=np.uint64) * m
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            ocounts = oidx.count_many_packed(q_np[: sample_n * m], off, nthreads=cores)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        parity = bool(np.array_equal(ocounts, counts_np[:sample_n]))
        assert parity, "GPU counts differ from the CPU oracle on the baseline sample"
        cpu = {"value": sample_n / best, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"first {sample_n} of the {nq} queries, best of 2, contiguous chunk per thread",
               "gpu_counts_equal_oracle_on_sample": parity}
        del oidx

    if rank != 0:
        return
    peak, peak_kind = measured_peak_gbs()
    steps_exec = int(st.lf_steps)
    R = int(info.rank_record_bytes)
    # SURVEY 8d per query: m + [8 if D>0] + 2*R*steps + 16; a query finished by text verification adds its
    # walk (R per LF step), one SA-sample sector and one sector of packed text instead of further LF steps
    seed_depth = int(info.seed_table_depth)
    table_depth = max(args.lookup_depth, seed_depth if m >= seed_depth else 0)
    # the library skips the per-batch suffix sort when a lookup level already replaces the steps the sort
    # would let neighbouring threads share (api.cu plan_sort): ns^depth >= queries
    sorted_batch = not (table_depth > 0 and int(info.num_searchable_dense_symbols) ** table_depth >= nq)
    alg_bytes = (nq * (m + 16 + (8 if table_depth > 0 else 0)) + 2 * R * steps_exec
                 + R * int(st.walk_steps) + 64 * int(st.verified_queries))
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    value = world * nq / (kernel_ms * 1e-3)
    dense_bytes = int(info.dense_suffix_array_bytes)
    traffic = ncu_traffic_bytes(args, st.verified_queries > 0, dense_bytes > 0, int(info.seed_table_depth) > 0)
    random_access = None
    if traffic:
        # What bounds the kernel: DRAM serves a random access as a whole 128 B line on this part (ncu on
        # tools/gather_bench.py: 126 B of dram__bytes_read per random 32 B load, profiles/r1_gather_dram.txt),
        # and random line fetches saturate at ~43 G/s = 5.5 TB/s (profiles/r1_gather_ceiling.json).
        random_access = {"dram_lines_per_launch": traffic / 128, "lines_per_s": traffic / 128 / (kernel_ms * 1e-3),
                         "measured_random_line_ceiling_per_s": 43.0e9,
                         "frac_of_ceiling": traffic / 128 / (kernel_ms * 1e-3) / 43.0e9,
                         "note": "whole step in the denominator"}
    out = {
        "metric": "len-50 count queries/s on 3.1 Gbp DNA index",
        "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "queries_per_gpu": nq,
                   "l2": "inputs larger than L2: 1.55 GB rank records accessed at random, 375 MB of queries",
                   "index_image_bytes": int(info.image_bytes), "dense_suffix_array_bytes": dense_bytes,
                   "seed_table_depth": seed_depth, "seed_table_bytes": int(info.seed_table_bytes),
                   "rank_record_bytes": R,
                   "lf_steps_per_step": steps_exec, "verified_queries_per_step": int(st.verified_queries),
                   "verify_walk_steps_per_step": int(st.walk_steps), "step_ms_min_median_max": [round(min(step_ms), 3),
                                                                            round(statistics.median(step_ms), 3),
                                                                            round(max(step_ms), 3)],
                   "launches_per_step": ("k_query_keys + cub radix sort (suffix order) + k_search" if sorted_batch
                                         else "k_search (a lookup level of depth %d replaces the shared steps: no sort)" % table_depth), "setup_s": {"data": round(t_data, 2), "index_build": round(t_build, 2),
                                                                 "replicate": round(t_bcast, 2)}},
        "clocks": clock_info,
        "e2e": {"value": world * nq / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": nq * m,
                "d2h_bytes_per_step": nq * 8, "ms_per_step": e2e_ms, "kernel_ms_inside": st.kernel_ms_search, "ms_per_step_by_rank_and_numa_node": e2e_ms_ranks,
                "gpu_launches_per_step": int(st.kernel_launches)},
        # this repo's kernels per step: k_search (+ k_query_keys when the batch is sorted; cub's 5 radix-sort
        # launches are library code and not counted)
        "gpu_launches": (2 if sorted_batch else 1) * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_kind": peak_kind, "random_access": random_access, "kernel": "k_search<K32, VERIFY>" if st.verified_queries else "k_search<K32>",
                     "algorithmic_bytes_per_launch": alg_bytes,
                     # what DRAM actually moved (ncu `traffic`) over the same time: every small random read (8 B
                     # seed entry, 4 B suffix-array entry, <= 25 B of text, a 32 B rank record) costs a 128 B line
                     "dram_achieved_gbs": (traffic / (kernel_ms * 1e-3) / 1e9) if traffic else None,
                     "dram_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
                     "queries_per_s_at_survey_ceiling": 6545.3e9 / 3266.0,
                     "rank_queries_per_s": 2 * steps_exec / (kernel_ms * 1e-3)},
        "cpu_baseline": cpu,
        "locate": locate,
    }
    emit_json(out)


_JSON_FD = None


def emit_json(obj):
    """The ONE line of the contract goes to the real stdout; everything else written to fd 1 by libraries
    (NCCL prints its version banner there) was redirected to stderr in main()."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
