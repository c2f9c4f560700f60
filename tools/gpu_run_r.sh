timeout 900 python -m pytest tests/test_gpu_tail_forms.py tests/test_gpu_matching.py tests/test_gpu_full_parity.py tests/test_gpu_api_contract.py tests/test_gpu_host_cpp.py tests/test_gpu_device_set.py -m gpu -x -q > gpurun_out/r02_gputest_r.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_r.log; tail -5 gpurun_out/r02_gputest_r.log
timeout 200 python tools/pair_latency_probe.py 2>&1 | head -12
timeout 300 python tools/tail_ab_probe.py 2>&1 | tail -2
