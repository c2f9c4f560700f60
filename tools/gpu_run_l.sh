python -m pytest tests/test_gpu_matching.py tests/test_gpu_full_parity.py tests/test_gpu_api_contract.py tests/test_gpu_device_set.py tests/test_gpu_window_sharding.py tests/test_gpu_host_cpp.py -m gpu -x -q > gpurun_out/r02_gputest_l.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_l.log; tail -5 gpurun_out/r02_gputest_l.log
python tools/device_set_probe.py 2>&1 | tail -4
SLAMB200_SET_THREADS=0 python tools/device_set_probe.py 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=1 value', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['other_kernels_ms_per_step'], d['gpu_launches'])"
python tools/pair_latency_probe.py 2>&1 | head -7
