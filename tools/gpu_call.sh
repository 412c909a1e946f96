#!/bin/bash
# row context table: targeted tests first, then the parity suites (the table is built automatically for DNA indexes),
# then the headline kernel with and without it
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_row_context.py -m gpu -q -x > gpurun_out/t_rowctx.log 2>&1; rc=$?; echo "rowctx rc=$rc"; tail -15 gpurun_out/t_rowctx.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_packed.py -m gpu -q -x > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/t_parity.log
for rc_on in 1 0; do
  GDX_ROW_CONTEXT=$rc_on timeout 400 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-locate > gpurun_out/bench_rowctx_$rc_on.json 2> gpurun_out/bench_rowctx_$rc_on.err; echo "bench ctx=$rc_on rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_rowctx_$rc_on.json')); print(json.dumps({'ctx':$rc_on,'value':d['value'],'ms':d['ms_per_step'],'e2e':d['e2e']['value'],'frac':d['roofline']['frac'],'rowctx_bytes':d['config'].get('row_context_table_bytes'),'setup':d['config']['setup_s']}))"
done
