#!/bin/bash
mkdir -p gpurun_out
timeout 60 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-locate > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_last.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_last.json')); print(json.dumps({'value':d['value'],'ms':d['ms_per_step'],'l2':d['config']['l2'],'traffic':d['roofline']['traffic'],'frac':d['roofline']['frac']}))"
