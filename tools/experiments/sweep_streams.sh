#!/bin/bash
# usage (GPU box): bash tools/experiments/sweep_streams.sh
for n in 148 592 1776 4096 8192 16384; do
  python bench.py --steps 60 --warmup 5 --streams $n --no-cpu-baseline > gpurun_out/sw.json 2>gpurun_out/sw.err || { echo "n=$n failed"; tail -3 gpurun_out/sw.err; continue; }
  python -c "
import json;d=json.load(open('gpurun_out/sw.json'));k=d['config']['per_kernel_ms']
print('streams=$n', {a:round(1e3*b,1) for a,b in k.items() if a!='note'}, 'step_us=%.1f value=%.0f e2e=%.0f frac=%.3f'%(1e3*d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']))"
done
