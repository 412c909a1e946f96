#!/bin/bash
# gpurun call 5 of round 2: full GPU suite, then protein config (tight timeouts)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2_t_all.log 2>&1; rc=$?; echo "all rc=$rc"
tail -22 gpurun_out/r2_t_all.log
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
