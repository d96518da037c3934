#!/bin/bash
# Compare prebuilt library variants (build_variants/*.so, see DESIGN.md "what was tried") on one box:
#   tools/experiments/variants.sh w5 w6 w8
cp opus-native_b200/libopusb200.so /tmp/orig.so
for v in "$@"; do
  cp build_variants/$v.so opus-native_b200/libopusb200.so
  for rep in 1 2; do
    timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/var.json 2>gpurun_out/var.err || { echo "$v failed"; tail -5 gpurun_out/var.err; }
    python -c "
import json;d=json.load(open('gpurun_out/var.json'));k=d['config']['per_kernel_ms'];print('$v', {a:round(1e3*b,1) for a,b in k.items() if a!='note'}, 'step_us=%.1f value=%.0f'%(1e3*d['ms_per_step'],d['value']))"
  done
done
cp /tmp/orig.so opus-native_b200/libopusb200.so
