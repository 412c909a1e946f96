#!/bin/bash
summ='import json,sys
d=json.loads(sys.stdin.readlines()[-1]); l=d.get("locate") or {}
print("value %.1fM q/s %s ms | e2e %.1fM q/s %.2f ms | frac %.3f | locate e2e %.1fM q/s %.2f ms" % (d["value"]/1e6, d["config"]["step_ms_min_median_max"], d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["roofline"]["frac"], l.get("value",0)/1e6, l.get("ms_per_step",0)))'
for cfg in "8 64 8" "4 32 4" "2 16 2" "16 48 4"; do
  set -- $cfg
  echo "== first=$1 max=$2 tail=$3"
  GDX_CHUNK_FIRST_MB=$1 GDX_CHUNK_MAX_MB=$2 GDX_CHUNK_TAIL_MB=$3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "$summ"
done
