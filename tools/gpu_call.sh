#!/bin/bash
# final code: whole GPU suite + full bench line
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "all rc=$?"; tail -3 gpurun_out/t_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2_bench.json')); print(json.dumps({'value':d['value'],'ms':d['ms_per_step'],'e2e':d['e2e']['value'],'pageable':d['e2e']['pageable']['value'],'prepacked':d['e2e']['prepacked']['value'],'locate':d['locate']['value'],'compact':d['locate']['compact']['value'],'noacc':d['no_accelerators']['value'],'noacc_e2e':d['no_accelerators']['e2e']['value'],'frac':d['roofline']['frac'],'cpu':d['cpu_baseline']['value'],'parity':d.get('oracle_parity'),'cores':d['e2e']['host_cores_per_rank']}))"
