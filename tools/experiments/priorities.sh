#!/bin/bash
# Stream priorities of the four pipeline stages (needs build_variants/prio.so: host_runtime.cpp built with an
# OPN_PRIO="rd,ex,k1,k2" override, 0 = lowest).  Three separate processes per setting: the spread between
# processes is as large as most effects.
cp opus-native_b200/libopusb200.so /tmp/orig.so
cp build_variants/prio.so opus-native_b200/libopusb200.so
for pr in ${PRIOS:-9,0,0,0 0,0,0,0 0,1,2,3 0,0,1,1 3,0,1,2}; do
  for rep in 1 2 3; do
    OPN_PRIO=$pr PYTHONPATH=. timeout 100 python tools/experiments/step_jitter.py 6 200 | sed "s/^/prio $pr: /"
  done
done
cp /tmp/orig.so opus-native_b200/libopusb200.so
