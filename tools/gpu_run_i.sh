python -m pytest tests/test_gpu_matching.py tests/test_gpu_api_contract.py -m gpu -x -q > gpurun_out/r02_gputest_i.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_i.log; tail -4 gpurun_out/r02_gputest_i.log
for n in -1 8; do python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --e2e-uploaders $n 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('narrowers', '$n', 'e2e', round(d['e2e']['value']), 'floor', round(d['e2e']['host_floor']['pairs_per_s_floor']), 'value', round(d['value']))"; done
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --e2e-steps 2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_e2e.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launch_e2e.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r02_launches_e2e.csv | head -12
