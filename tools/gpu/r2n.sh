#!/bin/bash
# round 2: band energies applied in SYNTH-CELT/2 (denormalise_bands), mixed-frame steps for both layouts, smooth_fade
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "celt2 or mixed or smooth" 2>&1 | tail -15 > $O/r2n_pytest_new.txt; cat $O/r2n_pytest_new.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r2n_pytest.txt; cat $O/r2n_pytest.txt
timeout 300 python bench.py --bitstream 2 --steps 200 --warmup 10 --no-cpu-baseline > $O/r2n_celt2.json 2> $O/r2n.err || tail -5 $O/r2n.err
python -c "
import json;d=json.load(open('$O/r2n_celt2.json'));print('celt2', d['ms_per_step'], d['value'], d['detail']['per_kernel_ms'], d['e2e']['value'])"
