python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in 1 2 3 4 6 8; do
  OPN_IMDCT_WPC=$w python bench.py --steps 100 --warmup 5 > gpurun_out/sw.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/sw.json'));print('wpc=$w', d['config']['per_kernel_ms'])"
done
