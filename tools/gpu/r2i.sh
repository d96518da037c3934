#!/bin/bash
# round 2: parity of the one-channel-per-pass frame kernel, then variants of its CTA shape against the previous kernel
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r2i_pytest.txt; cat $O/r2i_pytest.txt
STEPS=200 WARMUP=10 tools/experiments/variants.sh base_w5c4 p3_w7c4 p3_w8c4 p3_w6c4 p3_w9c3 2>&1 | tee $O/r2i_variants.txt
