timeout 600 python -m pytest tests/test_gpu_tail_forms.py -m gpu -x -q > gpurun_out/r02_gputest_p.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_p.log; tail -15 gpurun_out/r02_gputest_p.log
timeout 300 python tools/tail_ab_probe.py 2>&1 | tail -4
timeout 200 python tools/pair_latency_probe.py 2>&1 | head -8
