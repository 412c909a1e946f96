#!/bin/bash
# gpurun call 4 of round 2: full GPU suite + bench + A/B variants
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?"
tail -6 gpurun_out/r2_t_all.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench5.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench5.json')); print(json.dumps({'value':d['value'],'e2e':d['e2e']['value'],'pageable':d['e2e']['pageable']['value'],'prepacked':d['e2e']['prepacked']['value'],'locate':d['locate']['value'],'loc_ms':d['locate']['ms_per_step']}))"
AB="--steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-locate"
GDX_SEED_TABLE=16 timeout 600 python bench.py $AB > gpurun_out/r2_ab_seed16.json 2> gpurun_out/r2_ab_seed16.err; echo "seed16 rc=$?"
for rows in 2 3 4; do GENEDEX_B200_LIB=$PWD/genedex_b200/csrc/variants/libmultirow.so GDX_VERIFY_ROWS=$rows timeout 600 python bench.py $AB > gpurun_out/r2_ab_multirow$rows.json 2> gpurun_out/r2_ab_multirow$rows.err; echo "multirow$rows rc=$?"; done
GENEDEX_B200_LIB=$PWD/genedex_b200/csrc/variants/libmultirow.so GDX_VERIFY_ROWS=2 GDX_SEED_TABLE=16 timeout 600 python bench.py $AB > gpurun_out/r2_ab_seed16_multirow2.json 2> gpurun_out/r2_ab_seed16_multirow2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_ab_*.json')):
    try:
        d=json.load(open(f)); print(f, 'value %.3f G q/s  ms %.3f  lf_steps %d  e2e %.3f G' % (d['value']/1e9, d['ms_per_step'], d['config']['lf_steps_per_step'], d['e2e']['value']/1e9))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 900 python tools/run_configs.py c4d0 --out gpurun_out/r2_configs_kg5.jsonl > gpurun_out/r2_c4_kg5.log 2>&1; echo "c4 kg5 rc=$?"
GENEDEX_B200_LIB=$PWD/genedex_b200/csrc/variants/libkg4.so timeout 900 python tools/run_configs.py c4d0 --out gpurun_out/r2_configs_kg4.jsonl > gpurun_out/r2_c4_kg4.log 2>&1; echo "c4 kg4 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2_configs_kg5.jsonl','gpurun_out/r2_configs_kg4.jsonl'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, {k:d[k] for k in ('count_kernel_ms','count_e2e_ms','cursors_kernel_ms','locate_e2e_ms','lf_steps','verified_queries')})
    except Exception as e: print(f,'ERR',e)
PY
