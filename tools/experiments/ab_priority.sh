#!/bin/bash
# the driver's run shape (20 timed steps after 5 warm-up steps) with and without high-priority range-decode streams
for rep in 1 2 3; do
for pr in 0 1; do
OPN_RD_PRIORITY=$pr python bench.py --steps ${STEPS:-20} --warmup ${WARMUP:-5} --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('prio=$pr', round(1e3*d['ms_per_step'],1), round(d['value']))"
done; done
