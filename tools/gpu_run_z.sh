timeout 900 python -m pytest tests/test_gpu_tail_forms.py tests/test_gpu_matching.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/pair_latency_probe.py 2>&1 | head -8
timeout 300 python tools/tail_ab_probe.py 2>&1 | tail -2
