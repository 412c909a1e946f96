#!/bin/bash
# final evidence for the row-context kernel: whole GPU suite, full bench line, launch list
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "all rc=$?"; tail -3 gpurun_out/t_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2_bench.json')); print(json.dumps({'value':d['value'],'ms':d['ms_per_step'],'e2e':d['e2e']['value'],'pageable':d['e2e']['pageable']['value'],'prepacked':d['e2e']['prepacked']['value'],'locate':d['locate']['value'],'compact':d['locate']['compact']['value'],'noacc':d['no_accelerators']['value'],'frac':d['roofline']['frac'],'traffic':d['roofline']['traffic'],'cpu':d['cpu_baseline']['value'],'parity':d.get('oracle_parity'),'cores':d['e2e']['host_cores_per_rank']}))"
NCU="ncu --clock-control none --nvtx --nvtx-include timed/"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> gpurun_out/r2_ncu1.err; echo "ncu launches rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
