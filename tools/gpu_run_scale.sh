N=$1
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo bench_rc=$?; tail -3 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('host_floor'))
print(json.dumps(d.get('window_extras'), indent=1)[:3500])
PY
python -m pytest tests/test_gpu_device_set.py tests/test_gpu_window_sharding.py -m gpu -q 2>&1 | tail -3
