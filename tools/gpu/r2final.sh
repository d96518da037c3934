#!/bin/bash
# round 2, final pass: parity, smoke, every kept bench line, launch list, ncu captures
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q > $O/fin_pytest.log 2>&1; echo "pytest rc=$?" >> $O/fin_pytest.log; tail -4 $O/fin_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 5 > $O/fin_bench_ref.json 2> $O/fin_bench_ref.err; tail -2 $O/fin_bench_ref.err
python bench.py --steps 20 --warmup 5 > $O/fin_bench_20.json 2> $O/fin_bench_20.err; tail -3 $O/fin_bench_20.err
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/fin_bench_200.json 2> $O/fin_bench_200.err
python bench.py --steps 200 --warmup 10 --bitstream 2 --no-cpu-baseline > $O/fin_bench_200_celt2.json 2> $O/fin_bench_200_celt2.err; tail -3 $O/fin_bench_200_celt2.err
python bench.py --steps 200 --warmup 10 --mix > $O/fin_bench_200_mix.json 2> $O/fin_bench_200_mix.err; tail -3 $O/fin_bench_200_mix.err
python bench.py --steps 200 --warmup 10 --transient-permille 1000 --no-cpu-baseline > $O/fin_bench_200_transient.json 2> $O/fin_bench_200_transient.err
python - <<'PY'
import json
for f in ("fin_bench_ref","fin_bench_20","fin_bench_200","fin_bench_200_celt2","fin_bench_200_mix","fin_bench_200_transient"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d.get("e2e",{}).get("value",0)), "frac", d.get("roofline",{}).get("frac"), {k.split(" ")[0]:round(v,4) for k,v in d.get("detail",{}).get("per_kernel_ms",{}).items() if k!="note"})
    except Exception as e:
        print(f, "FAILED", e)
PY
cp opus-native_b200/libopusb200.so $O/fin_lib.so
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/fin_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/fin_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_frame_w|k_synth_rangedec' -s 8 -c 4 -o $O/fin_full -f python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/fin_ncu_f.log 2>&1
ls -la $O | tail -4
