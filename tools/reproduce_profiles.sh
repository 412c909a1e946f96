#!/bin/bash
# The commands behind profiles/ (run under gpurun on one B200; outputs land in gpurun_out/, the summaries
# are then written with tools/summarize_ncu.py).  Round 1.
set -x
mkdir -p gpurun_out
# bench lines
python bench.py --impl reference --gpus 1 --steps 5 --warmup 2          > gpurun_out/bench_ref.json
python bench.py --steps 10 --warmup 3                                   > gpurun_out/bench_r1_latest.json
GDX_VERIFY=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline    > gpurun_out/bench_lf_only.json
# other BASELINE configs at full size, with oracle parity
python tools/run_configs.py --out gpurun_out/configs_r1.jsonl
# random-access ceiling of the device
python tools/gather_bench.py gpurun_out/gather.json
# ncu: launch list of the bench command, then the headline kernel in full
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|Device" -c 800 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline
ncu --set full --clock-control none --import-source on -k regex:k_search -s 2 -c 1 -o gpurun_out/prof_search \
    python bench.py --steps 2 --warmup 1 --no-locate --no-cpu-baseline
ncu --set full --clock-control none --import-source on -k regex:k_locate_walk -c 1 -o gpurun_out/prof_walk \
    env GDX_VERIFY=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline
# summaries
python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches.txt
python tools/summarize_ncu.py report gpurun_out/prof_search.ncu-rep profiles/r1_k_search.txt
python tools/summarize_ncu.py report gpurun_out/prof_walk.ncu-rep profiles/r1_k_locate_walk.txt
# correctness tooling
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "kat or edge or verification"
python tools/fuzz_parity.py --seconds 300
