for k in 0 1 2 3 4; do echo skip=$k; SLAMB200_TAIL_SKIP=$k timeout 200 python tools/pair_latency_probe.py 2>&1 | head -4; done
