python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for sk in 0 7; do
  OPN_IMDCT_SKIP=$sk python bench.py --steps 60 --warmup 5 --streams 4096 --no-cpu-baseline > gpurun_out/sw.json 2>gpurun_out/sw.err || { echo "skip=$sk failed"; tail -3 gpurun_out/sw.err; continue; }
  python -c "
import json;d=json.load(open('gpurun_out/sw.json'));k=d['config']['per_kernel_ms'];print('skip=$sk sym_us=%.1f imdct_us=%.1f step_us=%.1f'%(1e3*k['k_synth_symbols'],1e3*k['k_imdct_post'],1e3*d['ms_per_step']))"
done
