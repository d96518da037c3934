#!/bin/bash
# round 2: SYNTH-CELT/2 range decode, packets per warp (divergence against idle lanes)
O=gpurun_out; mkdir -p $O
cp opus-native_b200/libopusb200.so /tmp/orig.so
for v in c2L32 c2L16 c2L8 c2L4 c2L2 c2L1; do
    cp build_variants/$v.so opus-native_b200/libopusb200.so
    timeout 300 python bench.py --bitstream 2 --steps 100 --warmup 10 --no-cpu-baseline > $O/var.json 2>$O/var.err || { echo "$v failed"; tail -5 $O/var.err; }
    python -c "
import json;d=json.load(open('$O/var.json'));k=d['detail']['per_kernel_ms'];print('$v', {a.split(' ')[0]:round(1e3*b,1) for a,b in k.items() if a!='note'}, 'step_us=%.1f value=%.0f'%(1e3*d['ms_per_step'],d['value']))"
done
cp /tmp/orig.so opus-native_b200/libopusb200.so
