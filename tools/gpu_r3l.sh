timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "ddc" 2>&1 | tail -2
