#!/bin/bash
# round 2: device-resident mixed-frame steps (OPN_FLAG_MIXED_FRAMES): parity, then the kept bench lines
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "mixed" 2>&1 | tail -15 > $O/r2j_pytest_mixed.txt; cat $O/r2j_pytest_mixed.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r2j_pytest.txt; cat $O/r2j_pytest.txt
timeout 300 python bench.py --mix --steps 200 --warmup 10 > $O/r2j_mix.json 2> $O/r2j_mix.err || tail -5 $O/r2j_mix.err
cat $O/r2j_mix.json | cut -c1-1500
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --transient-permille 1000 > $O/r2j_alltransient.json 2> $O/r2j_at.err || tail -5 $O/r2j_at.err
python -c "
import json;d=json.load(open('$O/r2j_alltransient.json'));print('all transient', d['ms_per_step'], d['value'], d['detail']['per_kernel_ms'])"
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/r2j_std.json 2> $O/r2j_std.err || tail -5 $O/r2j_std.err
python -c "
import json;d=json.load(open('$O/r2j_std.json'));print('standard', d['ms_per_step'], d['value'], d['detail']['per_kernel_ms'], d['roofline']['frac'])"
