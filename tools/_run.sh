python -m pytest tests/test_gpu_parity.py tests/test_gpu_toggles.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
python tools/fuzz_parity.py --seconds 60 --seed 31 2>&1 | tail -2
for sd in 15 14 16 0; do echo "seed=$sd"; GDX_SEED_TABLE=$sd python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-locate 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1fM q/s  %.3f ms | e2e %.1fM %.2f ms | seed %s %s B | lf %s | %s'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['e2e']['ms_per_step'],d['config']['seed_table_depth'],d['config']['seed_table_bytes'],d['config']['lf_steps_per_step'],d['config']['setup_s']))"; done
