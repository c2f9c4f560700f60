python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_h.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_h.log; tail -6 gpurun_out/r02_gputest_h.log
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err; tail -2 gpurun_out/r02_bench_h.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_h.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'])
print(json.dumps(d['e2e'], indent=1))
PY
for n in 8 12 16 22; do python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --e2e-uploaders $n 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('narrowers', '$n', 'e2e', round(d['e2e']['value']), 'floor', round(d['e2e']['host_floor']['pairs_per_s_floor']))"; done
python tools/graph_probe.py 2>&1 | tail -8
