python -m pytest tests -m gpu -q --timeout=1500 -p no:cacheprovider 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -c 600 gpurun_out/bench_r1_final.err
GDX_DENSE_SA=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_sampled_sa.json 2>/dev/null
python tools/run_configs.py --out gpurun_out/configs_r1_final.jsonl > gpurun_out/configs_final.log 2>&1; tail -3 gpurun_out/configs_final.log
