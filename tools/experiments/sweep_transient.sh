for t in 0 100 1000; do
python bench.py --steps 60 --warmup 5 --no-cpu-baseline --transient-permille $t > gpurun_out/sw.json 2>gpurun_out/sw.err || { echo "failed"; tail -5 gpurun_out/sw.err; }
python -c "
import json;d=json.load(open('gpurun_out/sw.json'));k=d['config']['per_kernel_ms'];print('transient=$t', {a:round(1e3*b,1) for a,b in k.items() if a!='note'}, 'step_us=%.1f value=%.0f frac=%.3f'%(1e3*d['ms_per_step'],d['value'],d['roofline']['frac']))"
done
