#!/bin/bash
# gpurun call 10 of round 2: final N=1 bench, ncu evidence, other configs
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2_bench.json')); print(json.dumps({'value':d['value'],'ms':d['ms_per_step'],'e2e':d['e2e']['value'],'pageable':d['e2e']['pageable']['value'],'prepacked':d['e2e']['prepacked']['value'],'locate':d['locate']['value'],'noacc':d['no_accelerators']['value'],'frac':d['roofline']['frac']}))"
NCU="ncu --clock-control none --nvtx --nvtx-include timed/"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> gpurun_out/r2_ncu1.err; echo "ncu launches rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:k_search -c 1 -o gpurun_out/r2_prof_search python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-locate > /dev/null 2> gpurun_out/r2_ncu2.err; echo "ncu search rc=$?"
GDX_PACK_HYBRID=0 timeout 600 $NCU --set full --import-source on -k regex:k_search --launch-skip 9 -c 1 -o gpurun_out/r2_prof_search_packed python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-locate > /dev/null 2> gpurun_out/r2_ncu3.err; echo "ncu packed rc=$?"
timeout 400 $NCU --set full --import-source on -k regex:k_locate_walk --launch-skip 2 -c 1 -o gpurun_out/r2_prof_walk_sampled python tools/locate_multi.py sampled > /dev/null 2> gpurun_out/r2_ncu4.err; echo "ncu walk sampled rc=$?"; grep "call" gpurun_out/r2_ncu4.err | tail -3
timeout 400 $NCU --set full --import-source on -k regex:k_locate_walk --launch-skip 2 -c 1 -o gpurun_out/r2_prof_walk_dense python tools/locate_multi.py dense > /dev/null 2> gpurun_out/r2_ncu5.err; echo "ncu walk dense rc=$?"; grep "call" gpurun_out/r2_ncu5.err | tail -3
rm -f gpurun_out/r2_configs.jsonl
for c in c2r c2mix c3 c4d0 c4d5; do timeout 420 python tools/run_configs.py $c --out gpurun_out/r2_configs.jsonl > gpurun_out/r2_cfg_$c.log 2>&1; echo "$c rc=$?"; done
python - <<'PY'
import json
for line in open('gpurun_out/r2_configs.jsonl'):
    d=json.loads(line); print(d['config'], {k:d.get(k) for k in ('count_kernel_ms','count_e2e_ms','cursors_kernel_ms','cursors_e2e_ms','locate_e2e_ms','hits','lf_steps','verified_queries','seed_table_depth')}, round(d['roofline']['frac'],3))
PY
