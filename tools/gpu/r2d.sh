#!/bin/bash
# round 2, call D: packed adds, range-decode refills off the chain, variants
set -x
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2d_pytest.log
tail -15 $O/r2d_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2d_bench_20.json 2> $O/r2d_bench_20.err; tail -3 $O/r2d_bench_20.err
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/r2d_bench_200.json 2> $O/r2d_bench_200.err
python - <<'PY'
import json
for f in ("r2d_bench_20","r2d_bench_200"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["detail"]["per_kernel_ms"].items() if k!="note"})
    except Exception as e:
        print(f, "FAILED", e)
PY
STEPS=200 WARMUP=10 bash tools/experiments/variants.sh $VARIANTS 2>&1 | tee $O/r2d_variants.log
cp opus-native_b200/libopusb200.so $O/r2d_lib.so
ncu --set full --clock-control none --import-source on -k regex:'k_frame_w|k_synth_rangedec' -s 6 -c 2 -o $O/r2d_full -f python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/r2d_ncu_f.log 2>&1
ls -la $O | tail -5
