#!/bin/bash
# gpurun call 9 of round 2 (8 GPUs): bench at N = 8, multi-GPU tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l; nproc; free -g | head -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 rc=$?"; tail -5 gpurun_out/r2_bench_n8.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n8.json')); print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','single_process','oracle_parity')})[:4500]); print(json.dumps(d['locate'])[:500]); print(d['config']['setup_s'])"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --durations=4 > gpurun_out/r2_t_multi8.log 2>&1; echo "multi rc=$?"; tail -8 gpurun_out/r2_t_multi8.log
