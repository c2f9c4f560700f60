lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|^CPU\(s\)|L3|Flags" | cut -c1-400 > gpurun_out/r02_host_info.txt 2>&1
numactl -H >> gpurun_out/r02_host_info.txt 2>&1
cat /sys/kernel/mm/transparent_hugepage/enabled >> gpurun_out/r02_host_info.txt 2>&1
nproc >> gpurun_out/r02_host_info.txt
free -g >> gpurun_out/r02_host_info.txt
nvidia-smi topo -m >> gpurun_out/r02_host_info.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2>&1 | tail -4
( time python bench.py --impl reference > gpurun_out/r02_bench_ref_default.json 2> gpurun_out/r02_bench_ref_default.err ) 2>&1 | tail -4
python tools/pack_bench.py 2>&1 | tail -20
