for d in 12 13 14; do
for srt in 1 0; do
echo "D=$d sort=$srt"; GDX_SORT_QUERIES=$srt python bench.py --lookup-depth $d --steps 10 --warmup 3 --no-cpu-baseline --no-locate 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1fM q/s  %.3f ms | e2e %.1fM %.2f ms | build %s'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['e2e']['ms_per_step'],d['config']['setup_s']))"
done; done
