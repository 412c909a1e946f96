python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "extend or cursor" 2>&1 | tail -4
python tools/run_configs.py c3 c4d0 --out gpurun_out/configs_r1_x.jsonl > gpurun_out/configs_x.log 2>&1; tail -2 gpurun_out/configs_x.log | cut -c1-1200
