python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_b.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_b.log; tail -5 gpurun_out/r02_gputest_b.log
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -2 gpurun_out/r02_bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_b.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac_of_burst'], d['roofline']['other_kernels_ms_per_step'], d['e2e']['value'])
PY
python tools/tc_modes.py 32 2>&1 | tail -12
