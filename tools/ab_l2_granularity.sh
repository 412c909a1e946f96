#!/bin/bash
# A/B of cudaLimitMaxL2FetchGranularity for the random-gather ceiling and the headline kernel
for g in 0 32 64; do
  echo "== GDX_L2_FETCH_GRANULARITY=$g"
  GDX_L2_FETCH_GRANULARITY=$g python tools/gather_bench.py gpurun_out/gather_g$g.json | grep -E '"table_gb": 1.5'
  GDX_L2_FETCH_GRANULARITY=$g python bench.py --steps 5 --warmup 3 --no-locate --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
done
