#!/bin/bash
mkdir -p gpurun_out
GDX_TRACE=1 timeout 300 python tools/debug_locate_invalid.py 2>&1 | grep -v "gdx trace\] chunk" | tail -40
