#!/bin/bash
# gpurun call 8 of round 2 (2 GPUs): multi-GPU tests, bench at N = 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader; nproc
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --durations=4 > gpurun_out/r2_t_multi.log 2>&1; echo "multi rc=$?"; tail -14 gpurun_out/r2_t_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r2_bench_n2.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2.json')); print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','single_process','oracle_parity')})[:3500]); print(json.dumps(d['locate'])[:600]); print(d['config']['setup_s'])"
