#!/bin/bash
# one gpurun call of round 2: new tests first, then the whole GPU suite, then a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_box.txt; nproc >> gpurun_out/r2_box.txt; free -g >> gpurun_out/r2_box.txt
timeout 900 python -m pytest tests/test_gpu_packed.py -m gpu -x -q > gpurun_out/r2_t_packed.log 2>&1; echo "packed rc=$?"
tail -25 gpurun_out/r2_t_packed.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_packed.py > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?"
tail -25 gpurun_out/r2_t_all.log
GDX_TRACE=0 timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench1.err; cat gpurun_out/r2_bench1.json | head -c 6000
