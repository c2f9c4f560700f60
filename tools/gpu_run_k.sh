python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_k.log 2>&1; echo rc=$? >> gpurun_out/r02_gputest_k.log; tail -6 gpurun_out/r02_gputest_k.log
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_k.json 2> gpurun_out/r02_bench_k.err; tail -2 gpurun_out/r02_bench_k.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_k.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['other_kernels_ms_per_step'], d['e2e']['value'])
PY
python tools/pair_latency_probe.py 2>&1 | head -8
ncu --set full --clock-control none --import-source on -k regex:"sift_tc_kernel|tc_tail_fused|compact_kernel" -c 3 -o gpurun_out/r02_final -f python tools/prof_run.py all > gpurun_out/ncu_final.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_final.ncu-rep 2>/dev/null | grep -E "kernel:|time_duration|dram__bytes|inst_executed.sum|tensor" 
