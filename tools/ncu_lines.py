#!/usr/bin/env python3
"""Rank source lines of one kernel by executed warp-instructions and stall samples.

Joins `ncu --page source --csv` (per-SASS-address metrics) with `nvdisasm -g` (address -> file:line,
including inlined callees) for the kernel's cubin extracted from libopusb200.so.

usage: tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top_n]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.environ.get("OPN_SO", os.path.join(ROOT, "opus-native_b200", "libopusb200.so"))  # the library the report was taken with


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # the export holds one block per profiled launch: ["Kernel Name", name] / header / instructions
    blocks, cur_blk = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur_blk = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur_blk)
        elif cur_blk is not None and r and r[0] == "Address":
            cur_blk["hdr"] = r
        elif cur_blk is not None and cur_blk["hdr"] and r:
            cur_blk["rows"].append(r)
    blk = [b for b in blocks if re.search(kre, b["name"])][0]
    kname, hdr = blk["name"], blk["hdr"]
    ia, ie, ism = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    isrc = hdr.index("Source")
    per_addr = []
    base = None
    for r in blk["rows"]:
        if len(r) <= ism or not r[ia]:
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        per_addr.append((a - base, int(r[ie] or 0), int(r[ism] or 0), r[isrc]))
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if f.startswith("opn_kernels")][0]
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
    # "void k<3, 2>(Args)" -> base name "k" plus the mangled integer template arguments "ILi3ELi2EE"
    m = re.search(r"([\w:]+)\s*(?:<([^>]*)>)?\s*\(", kname)
    fn = (m.group(1) if m else kname).split("::")[-1]
    def mangle(a):  # "(int)3" -> Li3E, "(bool)1" -> Lb1E
        t = re.match(r"\s*\((\w+)\)\s*(\w+)", a)
        code = {"int": "i", "bool": "b", "unsigned": "j"}.get(t.group(1), "i") if t else "i"
        return "L%s%sE" % (code, t.group(2) if t else a.strip())
    targs = "I" + "".join(mangle(a) for a in m.group(2).split(",")) + "E" if m and m.group(2) else ""
    line_of = {}
    cur, on = None, False
    for ln in sass.split("\n"):
        if ln.startswith(".text."):
            on = fn in ln and targs in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    agg = defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for off, inst, smp, _ in per_addr:
        k = line_of.get(off, ("?", 0))
        agg[k][0] += inst
        agg[k][1] += smp
        tot_i += inst
        tot_s += smp
    print(f"{kname}: {tot_i} warp-instructions, {tot_s} samples, {len(per_addr)} SASS instructions")
    src_cache = {}

    def src(f, l):
        if f not in src_cache:
            p = os.path.join(ROOT, "opus-native_b200", "csrc", f)
            src_cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
        s = src_cache[f]
        return s[l - 1].strip()[:100] if 0 < l <= len(s) else ""

    for (f, l), (inst, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * inst / max(tot_i, 1):5.1f}% inst {100 * smp / max(tot_s, 1):5.1f}% stall  {f}:{l:<4d} {src(f, l)}")


if __name__ == "__main__":
    main()
