#!/bin/bash
# gpurun call 3b of round 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_parity.py -m gpu -x -q -k "packed or uint32 or prepacked or sharded or unencodable or rank_variants" > gpurun_out/r2_t_packed.log 2>&1; echo "packed rc=$?"
tail -5 gpurun_out/r2_t_packed.log
timeout 300 python tools/host_pack_bench.py latency > gpurun_out/r2_host_pack_latency.txt 2>&1; cat gpurun_out/r2_host_pack_latency.txt
timeout 600 python tools/trace_e2e.py > /dev/null 2> gpurun_out/r2_trace_e2e.txt; grep -B1 -A40 "^call 2" gpurun_out/r2_trace_e2e.txt | head -5; awk '/call 1 ms/{f=1} f' gpurun_out/r2_trace_e2e.txt | head -40 | cut -c1-200
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench4.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench4.json')); print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e')})[:3000])"
NCU="ncu --clock-control none --nvtx --nvtx-include timed/"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> gpurun_out/r2_ncu1.err; echo "ncu launches rc=$?"
timeout 900 $NCU --set full --import-source on -k regex:k_search -c 1 -o gpurun_out/r2_prof_search python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-locate > /dev/null 2> gpurun_out/r2_ncu2.err; echo "ncu search rc=$?"
GDX_PACK_HYBRID=0 timeout 900 $NCU --set full --import-source on -k regex:"k_search.*1, 0, 1" -c 1 -o gpurun_out/r2_prof_search_packed python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-locate > /dev/null 2> gpurun_out/r2_ncu3.err; echo "ncu packed rc=$?"
