SLAMB200_UPLOAD_DMA=1 timeout 600 python -m pytest tests/test_gpu_host_cpp.py tests/test_gpu_matching.py -m gpu -x -q -k "packed or host or pinned or cfg3" 2>&1 | tail -2
for dma in 0 1 0 1; do
SLAMB200_UPLOAD_DMA=$dma SLAMB200_HOST_TRACE=1 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 8 2>gpurun_out/trace_d$dma.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('dma=$dma value', round(d['value']), 'e2e', round(e['value']), 'equal', e['results_equal_device_resident_run'], 'floor', round(e['host_floor']['pairs_per_s_floor']), round(e['host_floor']['ms_per_step_narrowing_alone'],2), e['host_threads'])"
grep match_batch_host gpurun_out/trace_d$dma.err | tail -2
done
