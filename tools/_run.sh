python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
export LENS=16,24,50 REPS=3
python tools/dram_by_length.py
python tools/dram_by_length.py
