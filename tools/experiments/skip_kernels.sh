#!/bin/bash
# Diagnostic (numbers are NOT bench values): step time of the resident loop with individual kernels left out
# after the first 8 calls.  Needs build_variants/skip.so, a build of host_runtime.cpp with an OPN_SKIP mask
# (1 range decode, 2 expand, 4 kernel 1, 8 kernel 2) around the four launches of run_bucket.
cp opus-native_b200/libopusb200.so /tmp/orig.so
cp build_variants/skip.so opus-native_b200/libopusb200.so
for m in ${MASKS:-0 1 4 8 12}; do
  OPN_SKIP=$m timeout 120 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/skip.json 2>gpurun_out/skip.err || { echo "mask $m failed"; tail -3 gpurun_out/skip.err; continue; }
  python -c "
import json;d=json.load(open('gpurun_out/skip.json'));print('skip mask $m: step_us=%.1f'%(1e3*d['ms_per_step']))"
done
cp /tmp/orig.so opus-native_b200/libopusb200.so
