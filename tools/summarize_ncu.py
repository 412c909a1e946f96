"""Turn ncu outputs (gpurun_out/) into the small text summaries kept under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches.txt
  python tools/summarize_ncu.py report   gpurun_out/prof.ncu-rep profiles/r1_k_search.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_lookup_hit.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def launches(src, dst, exclude=None):
    import re
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if exclude and re.search(exclude, r[iname]):
            continue
        v = float(r[ival].replace(",", ""))
        a = agg.setdefault(r[iname], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {src}\n")
        if exclude:
            f.write(f"# launches matching /{exclude}/ left out (index construction during setup, not on the search path)\n")
        f.write(f"# per kernel: total {rows[1][iunit]}, share of all listed launches, launches, name\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t:16.0f} {100 * t / total:6.2f}% x{c:<5d} {n[:150]}\n")


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {src}\n")
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"## {name[:160]}\n")
            for h, u, v in zip(hdr, units, vals):
                if h in KEYS:
                    f.write(f"{h:80s} {v:>22s} {u}\n")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](*sys.argv[2:])
