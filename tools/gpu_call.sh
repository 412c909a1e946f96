#!/bin/bash
# the library after moving the row context arithmetic into row_context.h: row context tests + packed tests + smoke
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_row_context.py tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/t_refactor.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_refactor.log
