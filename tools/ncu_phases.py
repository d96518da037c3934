#!/usr/bin/env python3
"""Per-phase share of executed warp-instructions and stall samples for one kernel.

usage: tools/ncu_phases.py <report.ncu-rep> <mangled-substr> <file> name:lo-hi [name:lo-hi ...]
A SASS instruction belongs to the first phase whose line range contains ANY frame of its inline
chain inside <file> (so a butterfly inlined into pass A counts as pass A)."""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "opus-native_b200", "libopusb200.so")


def main():
    rep, mangled, fname = sys.argv[1:4]
    phases = []
    for a in sys.argv[4:]:
        n, r = a.split(":")
        lo, hi = r.split("-")
        phases.append((n, int(lo), int(hi)))
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
    if os.environ.get("KREGEX"):  # report with several kernels: pick one
        cmd += ["--kernel-name", "regex:" + os.environ["KREGEX"]]
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][-1]
    hdr = rows[hi]
    ia, ie, ism = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if f.startswith("opn_kernels")][0]
        sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
    chain_of, cur, on, last = {}, None, False, None
    for ln in sass.split("\n"):
        if ln.startswith(".text."):
            on = mangled in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:  # -gi prints one line per inline level, innermost first, before the instruction
            if cur is None:
                cur = []
            cur.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            if cur:
                last = cur
            chain_of[int(m.group(1), 16)] = last
            cur = None

    def phase(chain):
        if chain:
            for n, lo, hi_ in phases:
                if any(f == fname and lo <= l <= hi_ for f, l in chain):
                    return n
        return "other"

    inst, smp = defaultdict(int), defaultdict(int)
    stalls = defaultdict(lambda: defaultdict(int))
    base = None
    for r in rows[hi + 1:]:
        if len(r) <= ism or not r[ia] or r[0] == "Address":
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        p = phase(chain_of.get(a - base))
        inst[p] += int(r[ie] or 0)
        smp[p] += int(r[ism] or 0)
        for h, i in stall_cols.items():
            stalls[p][h] += int(r[i] or 0)
    ti, ts = sum(inst.values()), sum(smp.values())
    print(f"{ti} warp-instructions, {ts} samples")
    for n in [p[0] for p in phases] + ["other"]:
        top = sorted(stalls[n].items(), key=lambda kv: -kv[1])[:4]
        tops = ", ".join(f"{h[6:]} {100 * v / max(smp[n], 1):.0f}%" for h, v in top if v)
        print(f"{n:12s} inst {100 * inst[n] / ti:5.1f}%  samples {100 * smp[n] / max(ts, 1):5.1f}%   [{tops}]")


if __name__ == "__main__":
    main()
