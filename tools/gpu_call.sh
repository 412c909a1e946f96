#!/bin/bash
# gpurun call 6 of round 2: packed tests first (fast fail), then the rest of the suite, then protein
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_packed.py -m gpu -x -q --durations=5 > gpurun_out/r2_t_packed.log 2>&1; rc=$?; echo "packed rc=$rc"; tail -12 gpurun_out/r2_t_packed.log
if [ $rc -eq 0 ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 --deselect tests/test_gpu_packed.py > gpurun_out/r2_t_all.log 2>&1; rc=$?; echo "all rc=$rc"; tail -16 gpurun_out/r2_t_all.log
fi
if [ $rc -eq 0 ]; then
  timeout 420 python tools/run_configs.py c4d0 --out gpurun_out/r2_configs_kg5.jsonl > gpurun_out/r2_c4_kg5.log 2>&1; echo "c4 kg5 rc=$?"
  python - <<'PY'
import json
for f in ('gpurun_out/r2_configs_kg5.jsonl',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, {k:d[k] for k in ('count_kernel_ms','count_e2e_ms','cursors_kernel_ms','locate_e2e_ms','lf_steps','verified_queries','seed_table_depth')})
    except Exception as e: print(f,'ERR',e)
PY
fi
