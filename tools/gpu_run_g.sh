nvidia-smi -L
python -m pytest tests/test_gpu_device_set.py tests/test_gpu_host_cpp.py -m gpu -x -q > gpurun_out/r02_gputest_g.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_g.log; tail -6 gpurun_out/r02_gputest_g.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo bench_n2_rc=$?; tail -2 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d['window_extras'].get('inprocess_device_set'), indent=1))
PY
for cfg in "4 -1" "12 0" "8 4" "2 14"; do set -- $cfg; python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --e2e-uploaders $1 --e2e-pack-threads $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('uploaders/pack', '$1', '$2', 'e2e', round(d['e2e']['value']), 'value', round(d['value']))"; done
nproc; python tools/pack_bench.py | tail -6
