# the round's N=1 verification: GPU tests, the default bench line, the ncu launch list of the
# device-resident steps and a --set full capture of the single-pair kernels
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final3.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_final3.log; tail -4 gpurun_out/r02_gputest_final3.log
( time python bench.py > gpurun_out/r02_bench_n1_final3.json 2> gpurun_out/r02_bench_n1_final3.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_final3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['other_kernels_ms_per_step'], d['e2e']['value'], d['e2e']['host_floor']['pairs_per_s_floor'], d['gpu_launches'])
print({k: (v.get('us_per_pair') or v.get('kernel_us_per_pair') or v.get('ms_per_host_call')) for k, v in d.get('extras', {}).items() if isinstance(v, dict)})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_final3.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_launch.log 2>&1; echo launchlist_rc=$?; tail -2 gpurun_out/ncu_launch.log
