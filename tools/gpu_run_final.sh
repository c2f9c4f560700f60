timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final2.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_final2.log; tail -4 gpurun_out/r02_gputest_final2.log
( time python bench.py > gpurun_out/r02_bench_n1_final2.json 2> gpurun_out/r02_bench_n1_final2.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_final2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['other_kernels_ms_per_step'], d['e2e']['value'], d['gpu_launches'])
print({k: (v.get('us_per_pair') or v.get('kernel_us_per_pair')) for k, v in d.get('extras', {}).items() if isinstance(v, dict)})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_final2.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo launchlist_rc=$?
ncu --set full --clock-control none --import-source on -k regex:"sift_tc_kernel|tc_tail_compact" -c 4 -o gpurun_out/r02_single_pair -f python tools/prof_run.py single > gpurun_out/ncu_single.log 2>&1; echo ncu_rc=$?
ls -la gpurun_out/r02_single_pair.ncu-rep
