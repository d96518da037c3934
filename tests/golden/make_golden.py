#!/usr/bin/env python3
"""Extract the known-answer vectors of the reference crate's own unit tests into
tests/golden/reference_kats.json.  Run in the build container (needs /root/reference);
the JSON travels to the GPU box, the reference does not.

Sources (all under /root/reference/src):
  range_coder/mod.rs:155-188   test_tell / test_tell_frac / test_tell_frac_limits literals
  celt/pvc.rs:439-451          test_pvq_v literals
  celt/comb_filter/mod.rs:207-224  TEST_VECTOR1 / TEST_VECTOR2 (+ the test's parameters :197-204)
  lib.rs:641-651               TEST_PACKET_* byte strings
"""
import json
import os
import re

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")


def read(p):
    return open(os.path.join(REF, p)).read()


def main():
    kats = {}
    rc = read("range_coder/mod.rs")
    tell = re.findall(r"TellImpl \{ bits_total: (\w+|u32::MAX), range: (\w+|u32::MAX) \}\.(tell|tell_frac)\(\), (\w+)\)", rc)

    def num(s):
        return 0xFFFFFFFF if s == "u32::MAX" else int(s, 0)

    kats["tell"] = [[num(a), num(b), num(d)] for a, b, f, d in tell if f == "tell"]
    kats["tell_frac"] = [[num(a), num(b), num(d)] for a, b, f, d in tell if f == "tell_frac"]
    m = re.search(r"entropy - ([\d.]+)\)", rc)
    kats["simple_uint_bits"] = {
        "entropy": float(m.group(1)),
        "tell_frac_over_8": float(re.search(r"-3\.0\) - ([\d.]+)\)", rc).group(1)),
        "range_bytes": int(re.search(r"range_bytes\(\), (\d+)\);\n\n        drop", rc).group(1)),
    }
    pvc = read("celt/pvc.rs")
    kats["pvq_v"] = [[int(a), int(b), int(c)] for a, b, c in re.findall(r"assert_eq!\(pvq_v\((\d+), (\d+)\), (\d+)\)", pvc)]
    pn = re.search(r"let pn: \[u32; 22\] = \[([^\]]*)\]", pvc).group(1)
    pk = re.search(r"let pk_max: \[u32; 22\] = \[([^\]]*)\]", pvc).group(1)
    kats["pvc_pn"] = [int(x) for x in pn.replace("\n", " ").split(",") if x.strip()]
    kats["pvc_pk_max"] = [int(x) for x in pk.replace("\n", " ").split(",") if x.strip()]
    comb = read("celt/comb_filter/mod.rs")
    for name in ("TEST_VECTOR1", "TEST_VECTOR2"):
        blk = re.search(name + r": &\[f32; N\] = &\[([^\]]*)\]", comb).group(1)
        kats["comb_" + name.lower()] = [float(x) for x in blk.replace("\n", " ").split(",") if x.strip()]
    consts = dict(re.findall(r"const (T0|T1|SIZE|N|OVERLAP): usize = (\d+);", comb))
    kats["comb_params"] = {k: int(v) for k, v in consts.items()}
    kats["comb_params"]["G0"] = float(re.search(r"const G0: f32 = ([\d.]+);", comb).group(1))
    kats["comb_params"]["G1"] = float(re.search(r"const G1: f32 = ([\d.]+);", comb).group(1))
    lib = read("lib.rs")
    for name in ("SINGLE", "CBR", "VBR", "INVALID"):
        blk = re.search(r"TEST_PACKET_" + name + r": &\[u8\] = &\[([^\]]*)\]", lib).group(1)
        kats["packet_" + name.lower()] = [int(x, 0) for x in blk.replace("\n", " ").split(",") if x.strip()]
    with open(OUT, "w") as f:
        json.dump(kats, f, indent=1)
    print("wrote", OUT, {k: (len(v) if hasattr(v, "__len__") else v) for k, v in kats.items()})


if __name__ == "__main__":
    main()
