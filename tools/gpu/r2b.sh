#!/bin/bash
# round 2, call B: fused frame kernel -- parity, bench (fused and unfused variants), launch list, ncu captures
set -x
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
tail -15 $O/r2b_pytest.log
tools/experiments/packed_math > $O/r2b_packed_math.log 2>&1; cat $O/r2b_packed_math.log
python bench.py --steps 20 --warmup 5 > $O/r2b_bench_20.json 2> $O/r2b_bench_20.err; tail -3 $O/r2b_bench_20.err
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/r2b_bench_200.json 2> $O/r2b_bench_200.err
OPN_UNFUSED_EXPAND=1 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/r2b_bench_200_unfused.json 2> $O/r2b_bench_200_unfused.err
python - <<'PY'
import json
for f in ("r2b_bench_20","r2b_bench_200","r2b_bench_200_unfused"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["detail"]["per_kernel_ms"])
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "comb or batch_decode_chain or synth_symbols or mixed_frame" > $O/r2b_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2b_memcheck.log; tail -4 $O/r2b_memcheck.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2b_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/r2b_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_frame_w|k_synth_rangedec' -s 6 -c 2 -o $O/r2b_full -f python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/r2b_ncu_f.log 2>&1
ls -la $O | tail -12
