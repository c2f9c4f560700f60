timeout 600 python -m pytest tests/test_gpu_matching.py -m gpu -x -q -k "general_float" 2>&1 | tail -3
