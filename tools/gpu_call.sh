#!/bin/bash
# gpurun call 7 of round 2: rest of the suite, protein config, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_toggles.py -m gpu -x -q --durations=8 -k "seed_table or toggles" > gpurun_out/r2_t_rest.log 2>&1; rc=$?; echo "rest rc=$rc"; tail -16 gpurun_out/r2_t_rest.log
timeout 420 python tools/run_configs.py c4d0 --out gpurun_out/r2_configs_kg5.jsonl > gpurun_out/r2_c4_kg5.log 2>&1; echo "c4 kg5 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2_configs_kg5.jsonl',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, {k:d[k] for k in ('count_kernel_ms','count_e2e_ms','cursors_kernel_ms','locate_e2e_ms','lf_steps','verified_queries','seed_table_depth')})
    except Exception as e: print(f,'ERR',e)
PY
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference rc=$?"; tail -3 gpurun_out/r2_bench_reference.err; cut -c1-2500 gpurun_out/r2_bench_reference.json
