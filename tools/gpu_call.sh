#!/bin/bash
# gpurun call 3a of round 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_parity.py -m gpu -x -q -k "packed or uint32 or prepacked or sharded or unencodable or rank_variants" > gpurun_out/r2_t_packed.log 2>&1; echo "packed rc=$?"
tail -5 gpurun_out/r2_t_packed.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench3.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench3.json')); print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e')})[:3000])"
GDX_TRACE=1 timeout 1200 python tools/run_configs.py c2r --out gpurun_out/r2_configs.jsonl > gpurun_out/r2_c2r.log 2> gpurun_out/r2_c2r_trace.txt; echo "c2r rc=$?"; grep -n "gdx trace" gpurun_out/r2_c2r_trace.txt | sed -n 30,75p | cut -c1-220
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kat or edge or invalid or unsearchable" > gpurun_out/r2_sanitizer_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck_smoke.log
