#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/host_pack_bench.py knobs > gpurun_out/r2_host_pack_knobs2.txt 2>&1; cat gpurun_out/r2_host_pack_knobs2.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2_bench.json')); print(json.dumps({'value':d['value'],'e2e':d['e2e']['value'],'e2e_ms':d['e2e']['ms_per_step'],'packed':d['e2e']['packed_queries_per_step'],'pageable':d['e2e']['pageable']['value'],'prepacked':d['e2e']['prepacked']['value'],'locate':d['locate']['value'],'compact':d['locate']['compact']['value']}))"
