"""Container-only check: the tables tools/gen_tables.py recomputes from formulas are
bit-identical to the literals in the reference crate.  Skipped where /root/reference
is absent (the GPU box)."""
import importlib.util
import os
import re
import struct

import pytest

REF = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _gen():
    spec = importlib.util.spec_from_file_location("gen_tables", os.path.join(ROOT, "tools", "gen_tables.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _block(path, start_pat):
    out, on = [], False
    for line in open(os.path.join(REF, path)).read().split("\n"):
        if re.search(start_pat, line):
            on = True
            continue
        if on:
            if re.match(r"^\s*\];", line):
                break
            out.append(line.split("//")[0])
    return "\n".join(out)


NUM = r"-?\d+\.\d+(?:e-?\d+)?"


def _bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def test_trig_bit_identical():
    g = _gen()
    ref = [float(s) for s in re.findall(NUM, _block("celt/mdct.rs", r"const TRIG"))]
    mine = g.trig_table()
    assert len(ref) == len(mine) == 1800
    assert [_bits(a) for a in ref] == [_bits(b) for b in mine]


def test_window_bit_identical():
    g = _gen()
    ref = [float(s) for s in re.findall(NUM, _block("celt/mode.rs", r"const WINDOW"))]
    mine = g.window_table()
    assert len(ref) == len(mine) == 120
    assert [_bits(a) for a in ref] == [_bits(b) for b in mine]


def test_twiddles_bit_identical():
    g = _gen()
    pairs = re.findall(r"r:\s*(" + NUM + r"),\s*i:\s*(" + NUM + ")", _block("celt/kiss_fft.rs", r"const TWIDDLES_480000_960"))
    mine = g.twiddle_table()
    assert len(pairs) == len(mine) == 480
    for (r, i), (mr, mi) in zip(pairs, mine):
        assert _bits(float(r)) == _bits(mr) or (float(r) == 0.0 and mr == 0.0)
        assert _bits(float(i)) == _bits(mi) or (float(i) == 0.0 and mi == 0.0)


@pytest.mark.parametrize("n", [480, 240, 120, 60])
def test_bitrev_identical(n):
    g = _gen()
    ref = [int(s) for s in re.findall(r"\d+", _block("celt/kiss_fft.rs", rf"const BITREV_{n}:"))]
    assert ref == g.bitrev_table(n)


def test_fft_factors_identical():
    g = _gen()
    src = open(os.path.join(REF, "celt/kiss_fft.rs")).read()
    found = re.findall(r"nfft: (\d+),.*?factors: \[([^\]]*)\]", src, flags=re.S)
    assert len(found) == 4
    for nfft, fac in found:
        fac = [int(x) for x in fac.split(",")]
        mine = g.FACTORS[int(nfft)]
        assert fac[: len(mine)] == mine and not any(fac[len(mine):])


def test_pvq_u_identical():
    g = _gen()
    rows_ref = [int(s) for s in re.findall(r"\d+", _block("celt/pvc.rs", r"const CELT_PVQ_U_ROW"))]
    data_ref = [int(s) for s in re.findall(r"\d+", _block("celt/pvc.rs", r"const CELT_PVQ_U_DATA"))]
    rows, data = g.pvq_tables()
    assert rows == rows_ref
    assert data == data_ref and len(data) == 1272


def test_allocation_tables_identical():
    g = _gen()
    for name, mine in (("ALLOC_VECTORS", g.ALLOC_VECTORS), ("CACHE_INDEX", g.CACHE_INDEX), ("CACHE_BITS", g.CACHE_BITS), ("CACHE_CAPS", g.CACHE_CAPS)):
        ref = [int(x) for x in re.findall(r"-?\d+", _block("celt/mode.rs", rf"const {name}:"))]  # _block starts after the declaration line
        assert ref == mine, name
    assert len(g.ALLOC_VECTORS) == 231 and len(g.CACHE_INDEX) == 105 and len(g.CACHE_BITS) == 392 and len(g.CACHE_CAPS) == 168


def test_mode_constants_identical():
    g = _gen()
    eb = [int(s) for s in re.findall(r"\d+", _block("celt/mode.rs", r"const E_BANDS"))]
    ln = [int(s) for s in re.findall(r"\d+", _block("celt/mode.rs", r"const LOG_N"))]
    assert eb == g.E_BANDS and ln == g.LOG_N
    gains = [float(s) for s in re.findall(NUM, _block("celt/comb_filter/mod.rs", r"const GAINS"))]
    assert [_bits(x) for x in gains] == [_bits(g.f32(q / 32768.0)) for q in g.COMB_GAINS_Q15]
