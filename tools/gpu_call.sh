#!/bin/bash
# final verification: the whole GPU suite on the final code, smoke, C3 with the sampled suffix array
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2_t_all.log 2>&1; rc=$?; echo "all rc=$rc"; tail -14 gpurun_out/r2_t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 420 python tools/run_configs.py c3s --out gpurun_out/r2_configs_c3s.jsonl > gpurun_out/r2_cfg_c3s.log 2>&1; echo "c3s rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_configs_c3s.jsonl').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('config','count_kernel_ms','count_e2e_ms','locate_e2e_ms','locate_kernels_ms','hits','walk_steps','dense_suffix_array_bytes')})
PY
