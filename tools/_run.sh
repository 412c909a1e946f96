python -m pytest tests -m gpu -q --timeout=1500 -p no:cacheprovider 2>&1 | tail -3
python tools/fuzz_parity.py --seconds 40 --seed 51 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final3.json 2> gpurun_out/bench_r1_final3.err; tail -c 300 gpurun_out/bench_r1_final3.err
python tools/run_configs.py --out gpurun_out/configs_r1_final3.jsonl > gpurun_out/configs_final3.log 2>&1; tail -1 gpurun_out/configs_final3.log | cut -c1-200
