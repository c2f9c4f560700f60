for nt in 0 1; do
SLAMB200_HOST_TRACE=1 SLAMB200_PACK_NT=$nt python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 6 2>gpurun_out/trace_$nt.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('NT=$nt value', round(d['value']), 'e2e', round(e['value']), 'floor', round(e['host_floor']['pairs_per_s_floor']), round(e['host_floor']['ms_per_step_narrowing_alone'],2))"
grep match_batch_host gpurun_out/trace_$nt.err | tail -4
done
