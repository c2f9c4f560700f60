# e2e A/B on one box: alternates the settings so that the shared host's drift hits both alike
for rep in 1 2 3; do
for cfg in "SLAMB200_UPLOAD_DMA=0 SLAMB200_HOST_CHUNK=14" "SLAMB200_UPLOAD_DMA=1 SLAMB200_HOST_CHUNK=14" "SLAMB200_UPLOAD_DMA=0 SLAMB200_HOST_CHUNK=7" "SLAMB200_UPLOAD_DMA=1 SLAMB200_HOST_CHUNK=28"; do
env $cfg python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('$cfg', 'e2e', round(e['value']), 'floor', round(e['host_floor']['pairs_per_s_floor']))"
done; done
