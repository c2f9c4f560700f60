nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_f.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_f.log; tail -8 gpurun_out/r02_gputest_f.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo bench_n2_rc=$?; tail -3 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], d['e2e'])
print(json.dumps(d.get('window_extras'), indent=1)[:3000])
PY
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_full.json 2> gpurun_out/r02_bench_n1_full.err; echo bench_n1_rc=$?; tail -3 gpurun_out/r02_bench_n1_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1_full.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], d['e2e'])
print(json.dumps(d.get('window_extras'), indent=1)[:2500])
print({k: v for k, v in d['extras'].items() if k in ('cfg1_sift_single_pair','cfg5_ransac_2048x5000','cfg2_orb_single_pair_tcgen05')})
PY
python tools/pair_latency_probe.py 2>&1 | tail -14
