#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/pack_tuning_ab.py > gpurun_out/r2_pack_tuning_ab.txt 2>&1; cat gpurun_out/r2_pack_tuning_ab.txt | tail -20
