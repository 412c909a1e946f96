#!/bin/bash
# sanitizer passes with every batch forced through the host packer / packed kernel / exception path
mkdir -p gpurun_out
GDX_PACK_MIN_BYTES=0 GDX_STAGE_MIN_BYTES=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kat or edge or invalid or unsearchable or verification_shortcut" > gpurun_out/r2_sanitizer_memcheck_packed_tests.log 2>&1; echo "memcheck packed rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_packed_tests.log
GDX_PACK_MIN_BYTES=0 GDX_STAGE_MIN_BYTES=0 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kat or edge or invalid" > gpurun_out/r2_sanitizer_racecheck_packed_tests.log 2>&1; echo "racecheck packed rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck_packed_tests.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rank_variants or reference_parts or many_hits" > gpurun_out/r2_sanitizer_memcheck_ingest.log 2>&1; echo "memcheck ingest rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_ingest.log
