# 2-GPU pass: device-set tests, sharded-window test with 2 ranks, N=2 and N=1 bench lines
nvidia-smi -L
python -m pytest tests/test_gpu_device_set.py tests/test_gpu_window_sharding.py tests/test_gpu_ransac.py tests/test_gpu_full_parity.py -m gpu -x -q -k "device_set or window or ransac or essential or cfg5 or chain" > gpurun_out/r02_gputest_e.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_e.log; tail -8 gpurun_out/r02_gputest_e.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo bench_n2_rc=$?; tail -3 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('window_extras'), indent=1)[:3000])
PY
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_full.json 2> gpurun_out/r02_bench_n1_full.err; echo bench_n1_rc=$?; tail -3 gpurun_out/r02_bench_n1_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1_full.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('window_extras'), indent=1)[:2500])
print(json.dumps(d.get('extras'), indent=1)[:3000])
PY
