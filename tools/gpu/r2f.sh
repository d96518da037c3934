#!/bin/bash
# round 2, call F: everything so far -- parity, headline bench, reference arm, SYNTH-CELT/2 line, launch list, ncu captures
set -x
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2f_pytest.log
tail -8 $O/r2f_pytest.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2f_bench_ref.json 2> $O/r2f_bench_ref.err; tail -2 $O/r2f_bench_ref.err
python bench.py --steps 20 --warmup 5 > $O/r2f_bench_20.json 2> $O/r2f_bench_20.err; tail -3 $O/r2f_bench_20.err
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/r2f_bench_200.json 2> $O/r2f_bench_200.err
python bench.py --steps 200 --warmup 10 --bitstream 2 > $O/r2f_bench_200_celt2.json 2> $O/r2f_bench_200_celt2.err; tail -3 $O/r2f_bench_200_celt2.err
python - <<'PY'
import json
for f in ("r2f_bench_ref","r2f_bench_20","r2f_bench_200","r2f_bench_200_celt2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "frac", d.get("roofline",{}).get("frac"), {k:round(v,4) for k,v in d.get("detail",{}).get("per_kernel_ms",{}).items() if k!="note"})
    except Exception as e:
        print(f, "FAILED", e)
PY
STEPS=200 WARMUP=10 bash tools/experiments/variants.sh $VARIANTS 2>&1 | tee $O/r2f_variants.log
cp opus-native_b200/libopusb200.so $O/r2f_lib.so
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2f_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2f_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_frame_w|k_synth_rangedec' -s 8 -c 4 -o $O/r2f_full -f python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/r2f_ncu_f.log 2>&1
ls -la $O | tail -5
