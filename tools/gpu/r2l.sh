#!/bin/bash
# round 2: SYNTH-CELT/2 range decode: register cap (co-residency with the frame kernel) and stream priority
O=gpurun_out; mkdir -p $O
cp opus-native_b200/libopusb200.so /tmp/orig.so
for v in c2rd_1 c2rd_16 c2rd_21; do
  for pr in 0 1; do
    cp build_variants/$v.so opus-native_b200/libopusb200.so
    OPN_RD_PRIORITY=$pr timeout 300 python bench.py --bitstream 2 --steps 100 --warmup 10 --no-cpu-baseline > $O/var.json 2>$O/var.err || { echo "$v failed"; tail -5 $O/var.err; }
    python -c "
import json;d=json.load(open('$O/var.json'));k=d['detail']['per_kernel_ms'];print('$v prio=$pr', {a.split(' ')[0]:round(1e3*b,1) for a,b in k.items() if a!='note'}, 'step_us=%.1f value=%.0f'%(1e3*d['ms_per_step'],d['value']))"
  done
done
cp /tmp/orig.so opus-native_b200/libopusb200.so
OPN_RD_PRIORITY=1 timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/var.json 2>$O/var.err
python -c "
import json;d=json.load(open('$O/var.json'));print('synth1 prio=1 step_us=%.1f'%(1e3*d['ms_per_step']))"
timeout 300 python bench.py --mix --steps 200 --warmup 10 > $O/r2l_mix.json 2>$O/var.err
python -c "
import json;d=json.load(open('$O/r2l_mix.json'));print('mix (long frames first) step_us=%.1f value=%.0f'%(1e3*d['ms_per_step'],d['value']), d['detail'])"
timeout 300 python -m pytest tests -m gpu -x -q -k "mixed or celt2" 2>&1 | tail -3
