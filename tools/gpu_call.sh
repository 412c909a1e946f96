#!/bin/bash
# gpurun call 11 of round 2: full suite on the final code, final bench line, C4 capture, C3 trace
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2_t_all.log 2>&1; rc=$?; echo "all rc=$rc"; tail -14 gpurun_out/r2_t_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2_bench.json')); print(json.dumps({'value':d['value'],'ms':d['ms_per_step'],'e2e':d['e2e']['value'],'pageable':d['e2e']['pageable']['value'],'prepacked':d['e2e']['prepacked']['value'],'locate':d['locate']['value'],'noacc':d['no_accelerators']['value'],'frac':d['roofline']['frac'],'dram_frac':d['roofline']['dram_frac']}))"
timeout 400 ncu --clock-control none --nvtx --nvtx-include timed/ --set full --import-source on -k regex:k_search --launch-skip 3 -c 1 -o gpurun_out/r2_prof_search_protein python tools/c4_kernel.py > /dev/null 2> gpurun_out/r2_ncu6.err; echo "ncu c4 rc=$?"; tail -2 gpurun_out/r2_ncu6.err
GDX_TRACE=1 timeout 420 python tools/run_configs.py c3 --out gpurun_out/r2_configs_c3b.jsonl > gpurun_out/r2_cfg_c3b.log 2> gpurun_out/r2_c3_trace.txt; echo "c3 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_configs_c3b.jsonl').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('count_kernel_ms','count_e2e_ms','cursors_e2e_ms','locate_e2e_ms')})
PY
grep "host: all chunks" gpurun_out/r2_c3_trace.txt | head -12
