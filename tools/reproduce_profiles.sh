#!/bin/bash
# The commands behind profiles/ (run under gpurun on one B200; outputs land in gpurun_out/, the summaries
# are then written with tools/summarize_ncu.py).  Round 1.
set -x
mkdir -p gpurun_out
# bench lines
python bench.py --impl reference --gpus 1 --steps 5 --warmup 2          > gpurun_out/bench_ref.json
python bench.py --steps 10 --warmup 3                                   > gpurun_out/bench_r1_latest.json
GDX_SEED_TABLE=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dense_sa_only.json
GDX_SEED_TABLE=0 GDX_DENSE_SA=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sampled_sa.json
GDX_SEED_TABLE=0 GDX_DENSE_SA=0 GDX_VERIFY=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lf_only.json
# other BASELINE configs at full size, with oracle parity
python tools/run_configs.py --out gpurun_out/configs_r1.jsonl
# random-access ceiling of the device, DRAM bytes per random load, DRAM cost of every search step
python tools/gather_bench.py gpurun_out/gather.json
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum \
    --clock-control none -k regex:k_gather --csv --log-file gpurun_out/exp_gather.csv python tools/gather_ncu.py
GDX_SEED_TABLE=0 GDX_DENSE_SA=0 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum \
    --clock-control none -k regex:k_search --csv --log-file gpurun_out/dram_by_len.csv python tools/dram_by_length.py
# ncu: launch list of the bench command, then the headline kernel in full
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|Device" -c 800 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline
ncu --set full --clock-control none --import-source on -k regex:k_search -s 2 -c 1 -o gpurun_out/prof_search \
    python bench.py --steps 2 --warmup 1 --no-locate --no-cpu-baseline
GDX_SEED_TABLE=0 ncu --set full --clock-control none --import-source on -k regex:k_search -s 2 -c 1 \
    -o gpurun_out/prof_search_dense_sa python bench.py --steps 2 --warmup 1 --no-locate --no-cpu-baseline
GDX_SEED_TABLE=0 GDX_DENSE_SA=0 ncu --set full --clock-control none --import-source on -k regex:k_search -s 2 -c 1 \
    -o gpurun_out/prof_search_sampled_sa python bench.py --steps 2 --warmup 1 --no-locate --no-cpu-baseline
ncu --set full --clock-control none --import-source on -k regex:k_locate_walk -c 1 -o gpurun_out/prof_walk \
    env GDX_VERIFY=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline
# summaries
python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches.txt
python tools/summarize_ncu.py report gpurun_out/prof_search.ncu-rep profiles/r1_k_search.txt
python tools/summarize_ncu.py report gpurun_out/prof_search_dense_sa.ncu-rep profiles/r1_k_search_dense_sa.txt
python tools/summarize_ncu.py report gpurun_out/prof_search_sampled_sa.ncu-rep profiles/r1_k_search_sampled_sa.txt
python tools/summarize_ncu.py report gpurun_out/prof_walk.ncu-rep profiles/r1_k_locate_walk.txt
# correctness tooling
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "kat or edge or verification"
python tools/fuzz_parity.py --seconds 300
