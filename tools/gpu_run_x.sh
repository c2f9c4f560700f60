timeout 900 python -m pytest tests/test_gpu_matching.py tests/test_gpu_full_parity.py tests/test_gpu_api_contract.py -m gpu -x -q 2>&1 | tail -3
for g in 0 4 2 8; do SLAMB200_GEN_SUB=$g timeout 300 python tools/gen_sub_probe.py 2>&1 | tail -6; done
