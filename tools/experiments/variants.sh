#!/bin/bash
# Compare prebuilt library variants (build_variants/*.so, see build_variant.sh) on one box:
#   tools/experiments/variants.sh w5c4 w10c2 ...          (STEPS / WARMUP from the environment, default 200 / 10)
cp opus-native_b200/libopusb200.so /tmp/orig.so
for v in "$@"; do
  cp build_variants/$v.so opus-native_b200/libopusb200.so
  for rep in 1 2; do
    timeout 180 python bench.py --steps ${STEPS:-200} --warmup ${WARMUP:-10} --no-cpu-baseline > gpurun_out/var.json 2>gpurun_out/var.err || { echo "$v failed"; tail -5 gpurun_out/var.err; }
    python -c "
import json;d=json.load(open('gpurun_out/var.json'));k=d['detail']['per_kernel_ms'];print('$v', {a.split(' ')[0]:round(1e3*b,1) for a,b in k.items() if a!='note'}, 'step_us=%.1f value=%.0f e2e=%.0f frac=%.3f'%(1e3*d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']))"
  done
done
cp /tmp/orig.so opus-native_b200/libopusb200.so
