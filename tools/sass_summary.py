#!/usr/bin/env python3
"""Per-kernel SASS mnemonic counts of libopusb200.so: the instructions that show what the kernels are made of --
UBLKCP / SYNCS (1-D TMA bulk copies and mbarriers), FADD2 (packed sums), FMUL / FADD (scalar), FFMA / FFMA2 / FMUL2
(must be absent from the float path: -fmad=false and no packed products), LDS / STS / LDG / STG, SHFL.

usage: tools/sass_summary.py [library] > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "opus-native_b200", "libopusb200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
counts, cur = collections.OrderedDict(), None
for ln in sass.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        counts[cur][m.group(1)] += 1
cols = ["UBLKCP", "SYNCS", "FADD2", "FADD", "FMUL", "FMUL2", "FFMA", "FFMA2", "LDS", "STS", "LDG", "STG", "SHFL", "MUFU"]
arch = re.findall(r"arch = (\S+)", subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout + sass)
print(f"# {os.path.basename(so)}: cubins for {sorted(set(arch))}; static SASS instruction counts per kernel")
print(f"{'kernel':70s} {'total':>6s} " + " ".join(f"{c:>6s}" for c in cols))
for k, c in counts.items():
    name = re.sub(r"^(void )?opn::", "", k)
    print(f"{name[:70]:70s} {sum(c.values()):6d} " + " ".join(f"{c[x]:6d}" for x in cols))
