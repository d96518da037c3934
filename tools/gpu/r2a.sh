#!/bin/bash
# round 2, call A: baseline of the round-1 pipeline on today's box, new parity tests, packed-math probe, sanitizers
set -x
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
tools/experiments/packed_math > $O/r2a_packed_math.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/r2a_bench_20.json 2> $O/r2a_bench_20.err
python bench.py --steps 400 --warmup 20 --no-cpu-baseline > $O/r2a_bench_400.json 2> $O/r2a_bench_400.err
SEL="imdct or comb or batch_decode_chain or synth_symbols or event_walk"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > $O/r2a_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "imdct or comb or batch_decode_chain" > $O/r2a_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/r2a_racecheck.log
tail -3 $O/r2a_pytest.log; cat $O/r2a_packed_math.log; tail -2 $O/r2a_memcheck.log; tail -2 $O/r2a_racecheck.log
