#!/bin/bash
summ='import json,sys
d=json.loads(sys.stdin.readlines()[-1]); l=d.get("locate") or {}
print("value %.1fM q/s | e2e %.1fM q/s %.2f ms | locate e2e %.1fM q/s %.2f ms" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], l.get("value",0)/1e6, l.get("ms_per_step",0)))'
for cfg in "8 64 8" "4 32 4" "4 64 4" "2 32 2" "4 48 2"; do
  set -- $cfg
  echo -n "first=$1 max=$2 tail=$3: "
  GDX_CHUNK_FIRST_MB=$1 GDX_CHUNK_MAX_MB=$2 GDX_CHUNK_TAIL_MB=$3 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$summ"
done
