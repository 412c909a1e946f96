for v in 8 5 4 3 2; do echo "VERIFY_MIN=$v"; GDX_VERIFY_MIN=$v python tools/run_configs.py c4d0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('count %.3f ms  cursors %.3f ms  locate e2e %.2f  lf_steps %d'%(d['count_kernel_ms'], d['cursors_kernel_ms'], d['locate_e2e_ms'], d['lf_steps']))"; done
echo DNA; for v in 8 4 3; do GDX_VERIFY_MIN=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-locate 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1fM q/s  %.3f ms'%(d['value']/1e6,d['ms_per_step']))"; done
