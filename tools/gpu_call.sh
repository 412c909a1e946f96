#!/bin/bash
mkdir -p gpurun_out
lscpu | grep -i "model name\|L3\|L2\|^CPU(s)" > gpurun_out/chunk_ab.txt
timeout 1000 python tools/chunk_ab.py >> gpurun_out/chunk_ab.txt 2> gpurun_out/chunk_ab.err; echo "rc=$?"; cat gpurun_out/chunk_ab.txt
