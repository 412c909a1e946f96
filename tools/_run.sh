python -m pytest tests -m gpu -q --timeout=1500 -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err; tail -c 400 gpurun_out/bench_r1_final2.err
GDX_SEED_TABLE=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_dense_only.json 2>/dev/null
python tools/run_configs.py --out gpurun_out/configs_r1_final2.jsonl > gpurun_out/configs_final2.log 2>&1; tail -1 gpurun_out/configs_final2.log | cut -c1-300
