#!/bin/bash
# final bench line at N = $GDX_BENCH_N on the final code
mkdir -p gpurun_out
N=${GDX_BENCH_N:-8}
nproc
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; F=gpurun_out/r2_bench.json
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"; F=gpurun_out/r2_bench_n$N.json
fi
python -c "
import json,sys; d=json.load(open('$F')); print(json.dumps({'n':d['n_gpus'],'value':d['value'],'ms':d['ms_per_step'],'e2e':d['e2e']['value'],'e2e_ms':d['e2e']['ms_per_step'],'packed':d['e2e']['packed_queries_per_step'],'pageable':d['e2e']['pageable']['value'],'prepacked':d['e2e']['prepacked']['value'],'locate':d['locate']['value'],'compact':d['locate']['compact']['value'],'noacc':d['no_accelerators']['value'],'noacc_e2e':d['no_accelerators']['e2e']['value'],'single':(d.get('single_process') or {}).get('value'),'parity':d['oracle_parity'],'cores':d['e2e']['host_cores_per_rank']}))"
