timeout 600 python -m pytest tests/test_gpu_host_cpp.py tests/test_gpu_api_contract.py tests/test_gpu_matching.py -m gpu -x -q 2>&1 | tail -3
for pt in 15 16 12; do
SLAMB200_HOST_TRACE=1 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 8 --e2e-pack-threads $pt 2>gpurun_out/trace_p$pt.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('pack=$pt value', round(d['value']), 'e2e', round(e['value']), 'equal', e['results_equal_device_resident_run'], 'floor', round(e['host_floor']['pairs_per_s_floor']), round(e['host_floor']['ms_per_step_narrowing_alone'],2))"
grep match_batch_host gpurun_out/trace_p$pt.err | tail -3
done
