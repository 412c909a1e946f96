#!/bin/bash
# N = 2 on the final code: replicas with all three accelerators, sharded batch, single-process arm
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n2.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2.json')); print(json.dumps({'n':d['n_gpus'],'value':d['value'],'ms':d['ms_per_step'],'e2e':d['e2e']['value'],'pageable':d['e2e']['pageable']['value'],'prepacked':d['e2e']['prepacked']['value'],'locate':d['locate']['value'],'compact':d['locate']['compact']['value'],'noacc':d['no_accelerators']['value'],'single':(d.get('single_process') or {}).get('value'),'parity':d['oracle_parity'],'cores':d['e2e']['host_cores_per_rank'],'setup':d['config']['setup_s'],'rowctx':d['config'].get('row_context_table_bytes')}))"
nvidia-smi --query-gpu=memory.used --format=csv | head -3
