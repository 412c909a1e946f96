#!/bin/bash
summ='import json,sys
d=json.loads(sys.stdin.readlines()[-1]); l=d.get("locate") or {}
print("value %.1fM q/s  %.2f ms | e2e %.1fM q/s %.2f ms | frac %.3f | locate e2e %.1fM q/s %.2f ms (walk kernels %.2f ms)" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["roofline"]["frac"], l.get("value",0)/1e6, l.get("ms_per_step",0), l.get("kernel_ms_locate",0)))'
for cfg in "8 96 16" "4 64 8" "16 128 24" "8 192 16"; do
  set -- $cfg
  echo "== first=$1 max=$2 tail=$3"
  GDX_CHUNK_FIRST_MB=$1 GDX_CHUNK_MAX_MB=$2 GDX_CHUNK_TAIL_MB=$3 python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "$summ"
done
