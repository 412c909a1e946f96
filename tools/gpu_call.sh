#!/bin/bash
# memcheck over all row context tests (incl. the packed kernel on a large batch), racecheck over the edge-case test
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_row_context.py -m gpu -q -x > gpurun_out/r2_sanitizer_memcheck_row_context.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_row_context.log
timeout 40 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_row_context.py -m gpu -q -x -k "policy or 0-8" > gpurun_out/r2_sanitizer_racecheck_row_context.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck_row_context.log
