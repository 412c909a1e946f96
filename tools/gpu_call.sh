#!/bin/bash
# gpurun call 2 of round 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_parity.py -m gpu -x -q -k "packed or uint32 or prepacked or sharded or unencodable or rank_variants" > gpurun_out/r2_t_packed.log 2>&1; echo "packed rc=$?"
tail -15 gpurun_out/r2_t_packed.log
timeout 600 python tools/host_pack_bench.py > gpurun_out/r2_host_pack_bench.txt 2>&1; cat gpurun_out/r2_host_pack_bench.txt
timeout 600 python tools/trace_e2e.py > /dev/null 2> gpurun_out/r2_trace_e2e.txt; tail -40 gpurun_out/r2_trace_e2e.txt
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench2.json')); print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','locate','no_accelerators')})[:5000])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_search|k_locate|k_expand|k_interval|k_add_base|k_query_keys|Radix|Scan" -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> gpurun_out/r2_ncu1.err; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_search -s 3 -c 1 -o gpurun_out/r2_prof_search python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-locate > /dev/null 2> gpurun_out/r2_ncu2.err; echo "ncu full rc=$?"
timeout 1200 python tools/run_configs.py c2r --out gpurun_out/r2_configs.jsonl > gpurun_out/r2_c2r.log 2>&1; echo "c2r rc=$?"; tail -3 gpurun_out/r2_c2r.log | cut -c1-3000
