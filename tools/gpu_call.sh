#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/host_pack_bench.py knobs > gpurun_out/r2_host_pack_knobs.txt 2>&1; cat gpurun_out/r2_host_pack_knobs.txt
