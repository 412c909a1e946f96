#!/bin/bash
# offsets guard: the new test first (short timeout), then the whole GPU suite, then the headline bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_packed.py -m gpu -q -x > gpurun_out/t_packed.log 2>&1; rc=$?; echo "packed rc=$rc"; tail -5 gpurun_out/t_packed.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "all rc=$?"; tail -3 gpurun_out/t_all.log
timeout 600 python bench.py > gpurun_out/bench_n1_guard.json 2> gpurun_out/bench_n1_guard.err; echo "bench rc=$?"; cat gpurun_out/bench_n1_guard.json | cut -c1-1500
