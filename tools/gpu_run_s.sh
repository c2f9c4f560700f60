for f in 2 1 2 1; do
SLAMB200_TAIL_FORM=$f python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('form $f value', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['other_kernels_ms_per_step'], d['gpu_launches'], d['clocks']['sm_mhz'])"
done
