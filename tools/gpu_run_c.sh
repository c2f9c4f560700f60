# profiling pass: launch list of the bench command, full capture of the 210-pair kernels, tensor metrics
set -x
ncu --query-metrics 2>/dev/null | grep -i "tensor" | head -60 > gpurun_out/r02_ncu_tensor_metrics_available.txt
python tools/prof_run.py all > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"sift_tc_kernel|tc_tail_fused|compact_kernel" -c 6 -o gpurun_out/r02_all -f python tools/prof_run.py all > gpurun_out/ncu_all.log 2>&1
ncu -i gpurun_out/r02_all.ncu-rep --page raw --csv > gpurun_out/r02_all_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_all.ncu-rep --page source --csv --kernel-name regex:tc_tail_fused > gpurun_out/r02_tail_source.csv 2>/dev/null
ncu -i gpurun_out/r02_all.ncu-rep --page source --csv --kernel-name regex:sift_tc_kernel > gpurun_out/r02_tc_source.csv 2>/dev/null
ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed_pipe_tmem.sum,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"sift_tc_kernel" -c 2 --csv --log-file gpurun_out/r02_tc_tensor_pipe.csv python tools/prof_run.py all > /dev/null 2>&1
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out | tail -12
