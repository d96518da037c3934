#!/bin/bash
# Final measurements of round 2 (second session): full GPU test suite, headline bench lines, SILK line, launch lists, ncu captures.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/f3_pytest.log 2>&1; tail -3 gpurun_out/f3_pytest.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/f3_bench_ref.json 2> gpurun_out/f3_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/f3_bench_20.json 2> gpurun_out/f3_bench_20.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f3_bench_20_2.json 2>> gpurun_out/f3_bench_20.err
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/f3_bench_200.json 2> gpurun_out/f3_bench_200.err
python bench.py --silk --steps 200 --warmup 10 > gpurun_out/f3_bench_silk.json 2> gpurun_out/f3_bench_silk.err
python bench.py --silk --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f3_bench_silk_20.json 2>> gpurun_out/f3_bench_silk.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/f3_silk_launches.csv python bench.py --silk --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f3_ncu_l.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_silk_frame -s 6 -c 3 -o gpurun_out/f3_silk_frame python bench.py --silk --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f3_ncu_f.log 2>&1
ncu --set full --clock-control none -k regex:k_silk_rangedec -s 3 -c 1 -o gpurun_out/f3_silk_rd python bench.py --silk --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f3_ncu_r.log 2>&1
ls -la gpurun_out/f3_*
