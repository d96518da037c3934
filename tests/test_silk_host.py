"""CPU tests of the SYNTH-SILK/1 pieces that need no GPU: the product's packet generator against the oracle's independent one,
the tables, and the properties of the oracle restatement (oracle/silk.c) that the GPU parity tests then rely on."""
import re
import os

import numpy as np
import pytest

import opus_native_b200 as opn
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table(name, ctype=float):
    txt = open(os.path.join(ROOT, "oracle", "oracle_tables.h")).read()
    body = re.search(r"ORC_" + name + r"\[\d+\] = \{(.*?)\};", txt, re.S).group(1)
    toks = [t.strip() for t in body.replace("\n", " ").split(",") if t.strip()]
    return np.array([float.fromhex(t.rstrip("f")) if ctype is float else int(t) for t in toks])


@pytest.mark.parametrize("bw,ms,ch,nb", [(2, 20, 1, 80), (2, 20, 2, 160), (2, 10, 1, 48), (1, 20, 1, 70), (1, 10, 2, 90), (0, 20, 2, 120), (0, 10, 1, 40)])
def test_product_and_oracle_generators_write_the_same_packets(bw, ms, ch, nb):
    a = opn.silk_fill(5, 24, 3, 6, bw, ms, ch, nb, n_threads=2)
    b = O.silk_fill(5, 24, 3, 6, bw, ms, ch, nb, n_threads=2)
    assert np.array_equal(a, b)
    toc = a[0, 0, 0]
    assert opn.query_packet_codec_mode(a[0, 0]) is not None
    assert (toc >> 3) == bw * 4 + (1 if ms == 20 else 0) and bool(toc & 4) == (ch == 2) and (toc & 3) == 0


def test_tables():
    for up in (3, 4, 6):
        h = _table(f"SILK_UP{up}").reshape(up, 8)
        assert np.allclose(h.sum(axis=1), 1.0, atol=1e-6)          # every phase has unity DC gain
        assert np.allclose(h, h[::-1, ::-1], atol=1e-7)              # linear phase: the prototype is symmetric
    ltp = _table("SILK_LTP_Q14", int).reshape(8, 5)
    assert np.all(ltp.sum(axis=1) < 0.8 * 16384) and np.all(ltp == ltp[:, ::-1])
    gains = _table("SILK_GAIN_Q10", int)
    assert gains[0] == 2048 and gains[-1] == 4096 * 1024 and np.all(np.diff(gains) > 0)
    for name, n in (("SILK_TYPE_ICDF", 3), ("SILK_DELTA_GAIN_ICDF", 9), ("SILK_CONTOUR_ICDF", 4), ("SILK_LTP_ICDF", 8)):
        t = _table(name, int)
        assert len(t) == n and t[-1] == 0 and np.all(np.diff(t) < 0) and t[0] < 256
    p = _table("SILK_PULSES_ICDF", int).reshape(2, 9)
    assert np.all(p[:, -1] == 0) and np.all(np.diff(p, axis=1) < 0)


def test_oracle_symbols_round_trip_through_the_generator():
    """The generator draws every symbol and range-encodes it; the decoder must read the same values back (the draws are
    recomputed here from the same splitmix64 stream)."""
    for s in range(12):
        pk = O.silk_fill(s, 1, 0, 1, 2, 20, 1, 96)[0, 0]
        side, exc, out16, pcm = O.SilkStream(1).decode(pk[1:], 2, 20, 1)
        ch = side.ch[0]
        assert 0 <= ch.type <= 2 and 16 <= ch.gidx[0] < 52 and all(0 <= g < 64 for g in ch.gidx)
        assert all(abs(ch.gidx[f] - ch.gidx[f - 1]) <= 1 for f in range(1, 4))
        assert all(0 <= r < (32 if k < 2 else 16) for k, r in enumerate(ch.rc_idx))
        if ch.type == 2:
            assert all(32 <= l <= 288 for l in ch.lag) and all(0 <= i < 8 for i in ch.ltp_idx)
        assert all(0 <= k <= 8 for k in ch.pulses)
        assert all(ch.index[b] < O.lib().orc_pvq_v(16, ch.pulses[b]) for b in range(20) if ch.pulses[b])
        assert side.tell_frac <= 8 * 8 * 95 and np.abs(pcm).max() > 0


def test_oracle_state_is_carried_and_resets():
    pk = O.silk_fill(3, 1, 0, 4, 2, 20, 1, 96)[:, 0]
    a = O.SilkStream(1)
    first = a.decode(pk[0, 1:], 2, 20, 1)[3]
    second = a.decode(pk[1, 1:], 2, 20, 1)[3]
    fresh = O.SilkStream(1).decode(pk[1, 1:], 2, 20, 1)[3]
    assert not np.array_equal(second, fresh)                          # filter / excitation / resampler history matter
    assert np.abs(first).max() < 1.0
    # a change of the internal rate restarts every filter: WB after NB equals WB from rest
    b = O.SilkStream(1)
    nbp = O.silk_fill(3, 1, 0, 1, 0, 20, 1, 60)[0, 0]
    b.decode(nbp[1:], 0, 20, 1)
    assert np.array_equal(b.decode(pk[1, 1:], 2, 20, 1)[3], fresh)


def test_oracle_lost_frames():
    st = O.SilkStream(2)
    assert np.all(st.decode(b"", 2, 20, 2, lost=True)[3] == 0)        # nothing decoded yet: silence, state untouched
    pk = O.silk_fill(8, 1, 0, 2, 2, 20, 2, 170)[:, 0]
    good = st.decode(pk[0, 1:], 2, 20, 2)[3]
    lost1 = st.decode(b"", 2, 20, 2, lost=True)[3]
    lost2 = st.decode(b"", 2, 20, 2, lost=True)[3]
    assert np.abs(good).max() > 0 and np.isfinite(lost1).all()
    # no excitation: the synthesis filter rings out
    assert np.abs(lost2[-200:]).max() <= np.abs(lost1[:200]).max() + 1e-9


def test_oracle_channel_mapping():
    """stream_channels -> channels (decoder.rs:332): a mono packet fills both outputs, a mono decoder takes the mid channel of a
    stereo packet, mid/side -> left/right is sat16(m +- s) on the internal-rate samples."""
    mono = O.silk_fill(2, 1, 0, 1, 2, 20, 1, 96)[0, 0]
    p1 = O.SilkStream(1).decode(mono[1:], 2, 20, 1)[3]
    p2 = O.SilkStream(2).decode(mono[1:], 2, 20, 1)[3].reshape(-1, 2)
    assert np.array_equal(p2[:, 0], p1) and np.array_equal(p2[:, 1], p1)
    st = O.silk_fill(2, 1, 0, 1, 2, 20, 2, 170)[0, 0]
    _, _, mid16, _ = O.SilkStream(1).decode(st[1:], 2, 20, 2)
    _, exc, lr16, _ = O.SilkStream(2).decode(st[1:], 2, 20, 2)
    m = mid16[0].astype(np.int64)
    l, r = lr16[0].astype(np.int64), lr16[1].astype(np.int64)
    ok = (np.abs(l) < 32767) & (np.abs(r) < 32767)
    assert ok.sum() > 200 and np.array_equal((l + r)[ok], 2 * m[ok])


def test_oracle_fec_decodes_the_redundant_copy():
    """decode_fec (LostFlag::DecodeFec): the packet's leading redundant block is decoded instead of its regular frame; a packet
    without one conceals; the regular decode of a packet with a copy walks through it and decodes what follows."""
    with_copy = O.silk_fill(4, 1, 0, 3, 2, 20, 1, 170, lbrr_permille=1000)[:, 0]
    without = O.silk_fill(4, 1, 0, 3, 2, 20, 1, 170, lbrr_permille=0)[:, 0]
    a = O.SilkStream(1)
    a.decode(with_copy[0, 1:], 2, 20, 1)
    fec_side, _, _, fec_pcm = a.decode(with_copy[1, 1:], 2, 20, 1, fec=True)
    reg_side, _, _, reg_pcm = a.decode(with_copy[1, 1:], 2, 20, 1)
    assert fec_side.lbrr == 1 and reg_side.lbrr == 1
    assert fec_side.tell_frac < reg_side.tell_frac          # the copy comes first, the regular frame after it
    assert not np.array_equal(fec_pcm, reg_pcm) and np.abs(fec_pcm).max() > 0
    b = O.SilkStream(1)
    b.decode(without[0, 1:], 2, 20, 1)
    side, _, _, pcm = b.decode(without[1, 1:], 2, 20, 1, fec=True)
    c = O.SilkStream(1)
    c.decode(without[0, 1:], 2, 20, 1)
    assert side.lbrr == 0 and np.array_equal(pcm, c.decode(b"", 2, 20, 1, lost=True)[3])   # no copy: concealment
