python -m pytest tests/test_gpu_matching.py tests/test_gpu_full_parity.py tests/test_gpu_api_contract.py tests/test_gpu_host_cpp.py -m gpu -x -q > gpurun_out/r02_gputest_o.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_o.log; tail -5 gpurun_out/r02_gputest_o.log
python tools/tail_ab_probe.py 2>&1 | tail -4
SLAMB200_NO_CARVEOUT=1 python tools/tail_ab_probe.py 2>&1 | tail -2
python tools/pair_latency_probe.py 2>&1 | head -8
