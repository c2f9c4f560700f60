for wc in "3 24" "4 18" "5 14" "6 12" "6 8" "7 10"; do set -- $wc; timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 8 --e2e-workers $1 --e2e-chunk $2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('workers $1 chunk $2 ->', round(d['e2e']['value']), 'pairs/s')"; done
