python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/sw.json 2>gpurun_out/sw.err || { echo "failed"; tail -3 gpurun_out/sw.err; }
python -c "
import json;d=json.load(open('gpurun_out/sw.json'));k=d['config']['per_kernel_ms'];print('sym_us=%.1f imdct_us=%.1f step_us=%.1f value=%.0f e2e=%.0f e2e_ms=%.3f'%(1e3*k['k_synth_symbols'],1e3*k['k_imdct_post'],1e3*d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))"
