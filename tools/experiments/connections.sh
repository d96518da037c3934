#!/bin/bash
# Step time of the resident loop against CUDA_DEVICE_MAX_CONNECTIONS (hardware work queues; default 8).  The
# pipeline uses 8 entropy streams + 3 more: with 8 queues some streams share one and pick up false dependencies.
for c in ${CONNS:-8 32}; do
  for rep in 1 2 3 4; do
    CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/conn.json 2>gpurun_out/conn.err || { echo "failed"; tail -3 gpurun_out/conn.err; continue; }
    python -c "
import json;d=json.load(open('gpurun_out/conn.json'));print('connections=$c step_us=%.1f value=%.0f e2e=%.0f'%(1e3*d['ms_per_step'],d['value'],d['e2e']['value']))"
  done
done
