"""SYNTH-CELT/2 on the CPU: the product's frame logic (csrc/celt2.cuh, compiled for the host inside the packet generator)
against the oracle's independent C restatement (oracle/celt2.c).  Both generators draw the same splitmix64 symbol values and
must write the same bytes: every budget decision of the allocation (band boosts, trim, compute_allocation, band skipping,
intensity / dual-stereo, fine bits, theta splits, leaf sizes) is driven by tell_frac, so a single diverging decision
changes the stream.  The side records (allocation vector, fine bits, priorities, coded bands, part and pulse counts,
final range, tell_frac) must agree field by field, and the oracle must decode what either generator wrote back to it.
PARITY UNPINNED: the reference holds the tables and primitives, not the frame logic (celt/decoder.rs:47-56 is todo!())."""
import ctypes as C

import numpy as np
import pytest

import opus_native_b200 as opn
import oracle_lib as O

CONFIGS = [(3, 2, 160), (3, 1, 100), (2, 2, 130), (1, 2, 100), (0, 2, 80), (0, 1, 48), (3, 2, 300), (3, 1, 48), (2, 1, 64), (1, 1, 60),
           (0, 2, 200), (3, 2, 600)]


def _fields_equal(rec, side, where):
    for f in opn.CELT2_SIDE_DTYPE.names:
        y = getattr(side, f)
        y = np.ctypeslib.as_array(y) if hasattr(y, "__len__") else y
        assert np.array_equal(rec[f], y), (where, f, rec[f], y)


@pytest.mark.parametrize("lm,channels,pkt_bytes", CONFIGS)
def test_generators_agree_byte_for_byte(lm, channels, pkt_bytes):
    for s in range(40):
        a, ta = np.zeros(pkt_bytes, np.uint8), np.zeros(1, opn.CELT2_SIDE_DTYPE)
        ra = opn.lib().opn_celt2_packet(s, 7, lm, channels, pkt_bytes, 300, a.ctypes.data, ta.ctypes.data)
        b, tb = np.zeros(pkt_bytes, np.uint8), O.Celt2Side()
        rb = O.lib().orc_celt2_packet(s, 7, lm, channels, pkt_bytes, 300, O.ptr(b), C.byref(tb))
        assert ra == rb, (s, ra, rb)
        if ra < 0:
            continue
        assert np.array_equal(a, b), (s, int(np.nonzero(a != b)[0][0]))
        _fields_equal(ta[0], tb, s)
        # and the oracle decodes the packet back to the same record
        side, _, y, coef = O.celt2_decode_symbols(a[1:], lm, channels)
        _fields_equal(ta[0], side, ("decode", s))
        assert int(np.abs(y).sum()) == side.n_pulses
        assert side.tell_frac // 8 <= 8 * (pkt_bytes - 1)
        if lm >= 1 and pkt_bytes <= 160:
            assert side.tell_frac // 8 >= 8 * (pkt_bytes - 1) - 8  # where the band caps do not bind, the allocation spends the whole budget


def test_allocation_follows_the_budget():
    """More bytes, more pulses; a transient frame with LM >= 2 reserves the anti-collapse bit."""
    pulses = []
    for pkt_bytes in (40, 80, 160, 320):
        n = [O.celt2_packet(s, 0, 3, 2, pkt_bytes, 0)[1].n_pulses for s in range(20)]
        pulses.append(np.mean(n))
    assert pulses == sorted(pulses) and pulses[-1] > 4 * pulses[0]
    seen = 0
    for s in range(40):
        t = O.celt2_packet(s, 1, 3, 1, 120, 1000)[1]
        assert t.transient == 1
        seen += t.anti_collapse
    assert 0 < seen < 40


def test_fill_is_packet_by_packet():
    blk = opn.celt2_fill(3, 5, 2, 3, 2, 2, 130, 100, n_threads=3)
    for f in range(3):
        for s in range(5):
            assert np.array_equal(blk[f, s], opn.celt2_packet(3 + s, 2 + f, 2, 2, 130, 100)[0])
