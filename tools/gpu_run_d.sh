python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_d.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_d.log; tail -25 gpurun_out/r02_gputest_d.log
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -2 gpurun_out/r02_bench_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_d.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['other_kernels_ms_per_step'], d['e2e']['value'])
PY
