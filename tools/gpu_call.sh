#!/bin/bash
# randomised parity with the row context table drawn at random (on in 70 % of the cases where it applies)
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/t_fuzz.log 2>&1; echo "fuzz test rc=$?"; tail -2 gpurun_out/t_fuzz.log
timeout 150 python tools/fuzz_parity.py --seconds 100 --seed 31 > gpurun_out/r2_fuzz_row_context.txt 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/r2_fuzz_row_context.txt
