python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for n in 4096; do
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --streams $n > gpurun_out/sw.json 2>gpurun_out/sw.err || { echo "failed"; tail -5 gpurun_out/sw.err; }
python -c "
import json;d=json.load(open('gpurun_out/sw.json'));k=d['config']['per_kernel_ms'];print('n=$n', {a:round(1e3*b,1) for a,b in k.items() if a!='note'}, 'step_us=%.1f value=%.0f e2e=%.0f e2e_i16=%.0f (%.3f ms) frac=%.3f'%(1e3*d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e_i16']['value'],d['e2e_i16']['ms_per_step'],d['roofline']['frac']))"
done
