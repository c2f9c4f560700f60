python tools/tc_trace.py 2>&1 | tail -40
python tools/graph_probe.py 2>&1 | tail -8
