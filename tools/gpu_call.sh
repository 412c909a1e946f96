#!/bin/bash
# gpurun call 12 of round 2: new tests (compact locate, repeat-rich at 2 %, packed fuzz), bench, long packed fuzz
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_full_size.py tests/test_gpu_fuzz.py -m gpu -x -q --durations=6 -k "compact or repeat_rich or every_batch_packed" > gpurun_out/r2_t_new.log 2>&1; rc=$?; echo "new tests rc=$rc"; tail -12 gpurun_out/r2_t_new.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2_bench.json')); print(json.dumps({'value':d['value'],'e2e':d['e2e']['value'],'prepacked':d['e2e']['prepacked']['value'],'locate':d['locate']['value'],'loc_ms':d['locate']['ms_per_step'],'compact':d['locate']['compact']}))"
GDX_PACK_MIN_BYTES=0 GDX_STAGE_MIN_BYTES=0 timeout 400 python tools/fuzz_parity.py --seconds 240 --seed 21 > gpurun_out/r2_fuzz_packed.txt 2>&1; echo "fuzz packed rc=$?"; tail -2 gpurun_out/r2_fuzz_packed.txt
timeout 300 python tools/fuzz_parity.py --seconds 180 --seed 22 > gpurun_out/r2_fuzz.txt 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/r2_fuzz.txt
