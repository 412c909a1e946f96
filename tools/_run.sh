python -m pytest tests/test_gpu_parity.py tests/test_gpu_toggles.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
python tools/fuzz_parity.py --seconds 45 --seed 21 2>&1 | tail -2
export LENS=16,24,50 REPS=3
echo dense; python tools/dram_by_length.py
echo nodense; GDX_DENSE_SA=0 python tools/dram_by_length.py
