#!/bin/bash
# round 2, call E: SYNTH-CELT/2 on the GPU, three launch groups, w12c2 variant
set -x
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2e_pytest.log
tail -25 $O/r2e_pytest.log
python bench.py --steps 20 --warmup 5 > $O/r2e_bench_20.json 2> $O/r2e_bench_20.err; tail -3 $O/r2e_bench_20.err
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/r2e_bench_200.json 2> $O/r2e_bench_200.err
python bench.py --steps 200 --warmup 10 --bitstream 2 > $O/r2e_bench_200_celt2.json 2> $O/r2e_bench_200_celt2.err; tail -3 $O/r2e_bench_200_celt2.err
python - <<'PY'
import json
for f in ("r2e_bench_20","r2e_bench_200","r2e_bench_200_celt2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e_pageable"]["value"]), "frac", round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["detail"]["per_kernel_ms"].items() if k!="note"})
    except Exception as e:
        print(f, "FAILED", e)
PY
STEPS=200 WARMUP=10 bash tools/experiments/variants.sh $VARIANTS 2>&1 | tee $O/r2e_variants.log
ls -la $O | tail -5
