timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_t.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_t.log; tail -4 gpurun_out/r02_gputest_t.log
for nt in 0 1 0 1; do
SLAMB200_PACK_NT=$nt python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('NT=$nt value', round(d['value']), 'e2e', round(e['value']), 'floor', round(e['host_floor']['pairs_per_s_floor']), round(e['host_floor']['ms_per_step_narrowing_alone'],2))"
done
