"""Pins the CPU oracle to the reference crate's own known-answer tests.
Every test names the reference test it restates.  CPU only."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

import oracle_lib as O

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
L = O.lib()


# ---- range_coder/mod.rs:155-188 test_tell / test_tell_frac / test_tell_frac_limits
def test_tell_kats():
    assert len(KATS["tell"]) == 13 and len(KATS["tell_frac"]) == 9
    for bits_total, rng, want in KATS["tell"]:
        assert L.orc_tell(bits_total, rng) == want
    for bits_total, rng, want in KATS["tell_frac"]:
        assert L.orc_tell_frac(bits_total, rng) == want


# ---- range_coder/mod.rs:191-263 test_simple_uint_bits
def test_simple_uint_bits():
    ops, vals, entropy = [], [], 0.0
    for ft in range(2, 1024):
        for i in range(ft):
            entropy += math.log(ft) * math.log2(math.e)
            ops.append((O.OP_UINT, ft, 0))
            vals.append(i)
    for ftb in range(1, 16):
        for i in range(1 << ftb):
            entropy += ftb
            ops.append((O.OP_BITS, ftb, 0))
            vals.append(i)
    ops = np.array(ops, dtype=O.OP_DTYPE)
    vals = np.array(vals, dtype=np.uint32)
    buf, tf, range_bytes, final_tf, err = O.enc_run_script(10 * 1024 * 1024, ops, vals)
    assert err == 0
    k = KATS["simple_uint_bits"]
    assert abs(entropy - k["entropy"]) < 2.3e-10  # f64::EPSILON in the reference; sum order differs
    assert final_tf / 8.0 == k["tell_frac_over_8"]
    assert range_bytes == k["range_bytes"] == 497192
    # raw bits cost exactly ftb bits each (mod.rs:209-219): tell_frac rises by 8*ftb
    nb = (1 << 16) - 2
    raw_tf = tf[-nb:].astype(np.int64)
    raw_bits = ops["a"][-nb:].astype(np.int64)
    assert np.all(np.diff(np.concatenate([[tf[-nb - 1]], raw_tf])) == 8 * raw_bits)
    out, _ = O.dec_run_script(buf, ops)
    assert np.array_equal(out["value"], vals)
    assert out["tell_frac"][-1] == final_tf
    assert np.array_equal(out["tell_frac"], tf)  # per-symbol enc/dec agreement (mod.rs:367-374)


# ---- range_coder/mod.rs:271-298 test_encoder_prefers_range_coder_data
def test_encoder_prefers_range_coder_data():
    buf = np.zeros(2, np.uint8)
    e = O.Enc()
    L.orc_enc_init(C.byref(e), O.ptr(buf), 2)
    L.orc_enc_bits(C.byref(e), 0x55, 7)
    for v, ft in [(1, 2), (1, 3), (1, 4), (1, 5), (2, 6), (6, 7)]:
        L.orc_enc_uint(C.byref(e), v, ft)
    L.orc_enc_done(C.byref(e))  # busts: the reference unwraps Ok here too
    d = O.Dec()
    L.orc_dec_init(C.byref(d), O.ptr(buf), 2)
    assert L.orc_dec_bits(C.byref(d), 7) == 0x05
    assert [L.orc_dec_uint(C.byref(d), ft) for ft in (2, 3, 4, 5, 6, 7)] == [1, 1, 1, 1, 2, 6]


# ---- range_coder/mod.rs:498-516 test_patch_initial_bits
def test_patch_initial_bits():
    buf = np.zeros(10000, np.uint8)
    e = O.Enc()
    L.orc_enc_init(C.byref(e), O.ptr(buf), len(buf))
    for v, lp in [(0, 1), (0, 1), (1, 6), (0, 2)]:
        assert L.orc_enc_bit_logp(C.byref(e), v, lp) == 0
    assert L.orc_enc_patch_initial_bits(C.byref(e), 0, 2) == 0
    assert L.orc_enc_done(C.byref(e)) == 0
    assert L.orc_enc_range_bytes(C.byref(e)) == 2
    assert buf[0] == 63


# ---- range_coder/mod.rs:519-528 test_shrink
def test_shrink():
    buf = np.zeros(10000, np.uint8)
    e = O.Enc()
    L.orc_enc_init(C.byref(e), O.ptr(buf), len(buf))
    for v in (1, 2, 3, 4):
        L.orc_enc_uint(C.byref(e), v, 255)
    L.orc_enc_done(C.byref(e))
    L.orc_enc_shrink(C.byref(e), 5)
    d = O.Dec()
    L.orc_dec_init(C.byref(d), O.ptr(buf), 5)
    assert [L.orc_dec_uint(C.byref(d), 255) for _ in range(4)] == [1, 2, 3, 4]


# ---- range_coder/mod.rs:301-377 test_random_data (property; numpy RNG replaces nanorand)
def test_random_data_roundtrip_and_tell():
    rnd = np.random.default_rng(42)
    for _ in range(256):
        ft = int(rnd.integers(2, 1024))
        sz = int(rnd.integers(128, 512))
        zeros = rnd.integers(0, 14) == 0
        data = np.zeros(sz, np.uint32) if zeros else rnd.integers(0, ft, sz).astype(np.uint32)
        ops = np.array([(O.OP_UINT, ft, 0)] * sz, dtype=O.OP_DTYPE)
        buf, tf, rb, ftf, err = O.enc_run_script(10000, ops, data)
        assert err == 0
        out, _ = O.dec_run_script(buf, ops)
        assert np.array_equal(out["value"], data)
        assert np.array_equal(out["tell_frac"], tf)
        assert (ftf // 8 + 7) // 8 + 1 >= rb


# ---- range_coder/mod.rs:381-495 test_compatibility: 4 encode x 4 decode methods for binary symbols
def test_compatibility():
    rnd = np.random.default_rng(42)
    meth = [O.OP_BIT_VIA_DECODE, O.OP_BIT_VIA_DECODE_BIN, O.OP_BIT_LOGP, O.OP_ICDF]
    pool = np.array([1, 0], np.uint8)
    for _ in range(256):
        sz = int(rnd.integers(128, 512))
        data = rnd.integers(0, 2, sz).astype(np.uint32)
        logp = rnd.integers(1, 17, sz)

        def mk(methods):
            return np.array([(meth[m], 0, lp) if meth[m] == O.OP_ICDF else (meth[m], lp, 0) for m, lp in zip(methods, logp)], dtype=O.OP_DTYPE)

        eops, dops = mk(rnd.integers(0, 4, sz)), mk(rnd.integers(0, 4, sz))
        buf, tf, rb, ftf, err = O.enc_run_script(10000, eops, data, icdf_pool=pool)
        assert err == 0
        out, _ = O.dec_run_script(buf, dops, icdf_pool=pool)
        assert np.array_equal(out["value"], data)
        assert np.array_equal(out["tell_frac"], tf)


# ---- range_coder/mod.rs:537-570 test_laplace
def test_laplace():
    rnd = np.random.default_rng(42)
    val = (rnd.integers(0, 16, 10000) - 7).astype(np.int32)
    decay = rnd.integers(5000, 16000, 10000).astype(np.uint32)
    val[:3] = [3, 0, -1]
    decay[:3] = [6000, 5800, 5600]
    ops = np.array([(O.OP_LAPLACE, L.orc_laplace_start_freq(int(d)), int(d)) for d in decay], dtype=O.OP_DTYPE)
    buf, tf, rb, ftf, err = O.enc_run_script(40000, ops, val.view(np.uint32))
    assert err == 0
    out, _ = O.dec_run_script(buf, ops)
    assert np.array_equal(out["value"].view(np.int32), val)
    assert np.array_equal(out["tell_frac"], tf)


# ---- celt/pvc.rs:439-451 test_pvq_v
def test_pvq_v_kats():
    assert len(KATS["pvq_v"]) == 11
    for n, k, want in KATS["pvq_v"]:
        assert L.orc_pvq_v(n, k) == want


# ---- celt/pvc.rs:453-503 test_pvc
def test_pvc_roundtrip():
    def get_pulses(i):
        return i if i < 8 else (8 + (i & 7)) << ((i >> 3) - 1)

    for n, kmax in zip(KATS["pvc_pn"], KATS["pvc_pk_max"]):
        for pseudo in range(1, 41):
            k = get_pulses(pseudo)
            if k > kmax:
                break
            nc = L.orc_pvq_v(n, k)
            inc = max(nc // 2000, 1)  # reference uses nc/20000; thinned to keep the CPU suite short
            y = np.zeros(n, np.int32)
            for i in range(0, nc, inc):
                yy = L.orc_cwrsi(O.ptr(y), n, k, i)
                assert int(np.abs(y).sum()) == k
                assert yy == float((y.astype(np.int64) ** 2).sum())
                assert L.orc_icwrs(O.ptr(y), n) == i


# ---- celt/kiss_fft.rs:594-703 test_dft: forward and inverse vs O(n^2) f64 DFT, SNR > 130 dB
@pytest.mark.parametrize("shift,nfft", [(3, 60), (2, 120), (1, 240), (0, 480)])
@pytest.mark.parametrize("inverse", [False, True])
def test_dft(shift, nfft, inverse):
    rnd = np.random.default_rng(42)
    x = ((rnd.integers(0, 32767, nfft) - 16384) + 1j * (rnd.integers(0, 32767, nfft) - 16384)) * 32768.0
    if inverse:
        x = x / nfft
    x = x.astype(np.complex64)
    bitrev = np.ctypeslib.as_array(L.orc_fft_bitrev(shift), shape=(nfft,))
    buf = np.zeros(nfft, np.complex64)
    if inverse:
        buf[bitrev] = np.conj(x)
    else:
        buf[bitrev] = (x * np.float32(L.orc_fft_scale(shift))).astype(np.complex64)
    L.orc_fft_process(shift, O.ptr(buf))
    got = np.conj(buf) if inverse else buf
    k = np.arange(nfft)
    W = np.exp((2j if inverse else -2j) * np.pi * np.outer(k, k) / nfft)
    want = W @ x.astype(np.complex128) / (1 if inverse else nfft)
    snr = 10 * np.log10((np.abs(want) ** 2).sum() / (np.abs(want - got) ** 2).sum())
    assert snr > 130.0


# ---- celt/mdct.rs:639-757 test_mdct: forward SNR > 130 dB, inverse SNR > 60 dB vs f64 definition
@pytest.mark.parametrize("shift,nfft", [(3, 240), (2, 480), (1, 960), (0, 1920)])
def test_mdct_forward(shift, nfft):
    rnd = np.random.default_rng(42)
    x = ((rnd.integers(0, 32768, nfft) - 16384) * 32768.0).astype(np.float32)
    out = np.zeros(nfft, np.float32)
    win = np.ones(nfft // 2, np.float32)
    L.orc_mdct_forward(O.ptr(x.copy()), O.ptr(out), O.ptr(win), nfft // 2, shift, 1)
    i = np.arange(nfft // 2)[:, None]
    k = np.arange(nfft)[None, :]
    want = (np.cos(2 * np.pi * (k + 0.5 + 0.25 * nfft) * (i + 0.5) / nfft) / (nfft // 4)) @ x.astype(np.float64)
    got = out[: nfft // 2].astype(np.float64)
    assert 10 * np.log10((want ** 2).sum() / ((want - got) ** 2).sum()) > 130.0


@pytest.mark.parametrize("shift,nfft", [(3, 240), (2, 480), (1, 960), (0, 1920)])
def test_mdct_backward(shift, nfft):
    rnd = np.random.default_rng(42)
    x = ((rnd.integers(0, 32768, nfft) - 16384) * 32768.0 / nfft).astype(np.float32)
    out = np.zeros(nfft, np.float32)
    win = np.ones(nfft // 2, np.float32)
    L.orc_mdct_backward(O.ptr(x), O.ptr(out), O.ptr(win), nfft // 2, shift, 1)
    out[nfft - 1 - np.arange(nfft // 4)] = out[nfft // 2 + np.arange(nfft // 4)]  # mdct.rs:735-738
    i = np.arange(nfft)[:, None]
    k = np.arange(nfft // 2)[None, :]
    want = np.cos(2 * np.pi * (i + 0.5 + 0.25 * nfft) * (k + 0.5) / nfft) @ x[: nfft // 2].astype(np.float64)
    assert 10 * np.log10((want ** 2).sum() / ((want - out) ** 2).sum()) > 60.0


# ---- celt/comb_filter/mod.rs:227-270 golden vectors
def test_comb_filter_golden():
    p = KATS["comb_params"]
    size, n = p["SIZE"], p["N"]
    x = np.arange(size, dtype=np.float32)
    y = np.zeros(size, np.float32)
    O.comb_filter(y, size - n, x, size - n, p["T0"], p["T1"], n, p["G0"], p["G1"], 0, 0, p["OVERLAP"])
    want = np.array(KATS["comb_test_vector1"], np.float32)
    assert np.all(np.abs(1.0 - y[size - n:] / want) < 1e-5)
    assert np.array_equal(y[size - n:], want)  # in fact bit-identical to the printed literals


def test_comb_filter_inplace_golden():
    p = KATS["comb_params"]
    size, n = p["SIZE"], p["N"]
    y = np.arange(size, dtype=np.float32)
    O.comb_filter_inplace(y, size - n, p["T0"], p["T1"], n, p["G0"], p["G1"], 0, 0, p["OVERLAP"])
    want = np.array(KATS["comb_test_vector2"], np.float32)
    assert np.all(np.abs(1.0 - y[size - n:] / want) < 1e-5)


# ---- math.rs:237-298 bitexact trig checksums
def test_bitexact_cos():
    chk, max_d, min_d, last = 0, 0, 32767, 32767
    for i in range(64, 16321):
        q = L.orc_bitexact_cos(i)
        chk ^= q * i
        d = last - q
        max_d, min_d, last = max(max_d, d), min(min_d, d), q
    assert (L.orc_bitexact_cos(64), L.orc_bitexact_cos(16320), L.orc_bitexact_cos(8192)) == (32767, 200, 23171)
    assert (chk, max_d, min_d) == (89408644, 5, 0)


def test_bitexact_log2tan():
    chk, max_d, min_d, last = 0, 0, 15059, 15059
    for i in range(64, 8193):
        mid, side = L.orc_bitexact_cos(i), L.orc_bitexact_cos(16384 - i)
        q = L.orc_bitexact_log2tan(mid, side)
        assert q == -L.orc_bitexact_log2tan(side, mid)
        chk ^= q * i
        d = last - q
        max_d, min_d, last = max(max_d, d), min(min_d, d), q
    assert (chk, max_d, min_d) == (15821257, 61, -2)
    assert L.orc_bitexact_log2tan(32767, 200) == 15059
    assert L.orc_bitexact_log2tan(30274, 12540) == 2611
    assert L.orc_bitexact_log2tan(23171, 23171) == 0


# ---- lib.rs:653-768 query_packet_* tables
def test_query_packet_tables():
    bw = [L.orc_packet_bandwidth(O.ptr(np.array([c << 3], np.uint8))) for c in range(32)]
    assert bw == [0] * 4 + [1] * 4 + [2] * 4 + [3, 3, 4, 4] + [0] * 4 + [2] * 4 + [3] * 4 + [4] * 4
    fs = [L.orc_packet_samples_per_frame(O.ptr(np.array([c << 3], np.uint8)), 48000) for c in range(32)]
    assert fs == [480, 960, 1920, 2880] * 3 + [480, 960, 480, 960] + [120, 240, 480, 960] * 4
    b = lambda *x: O.ptr(np.array(x, np.uint8))
    assert L.orc_packet_channels(b(0)) == 1 and L.orc_packet_channels(b(4)) == 2
    assert [L.orc_packet_frame_count(b(0), 1), L.orc_packet_frame_count(b(1), 1), L.orc_packet_frame_count(b(2), 1)] == [1, 2, 2]
    assert L.orc_packet_frame_count(b(3), 1) < 0 and L.orc_packet_frame_count(b(3, 5), 2) == 5
    assert L.orc_packet_sample_count(b(70), 1, 48000) == 960
    assert L.orc_packet_sample_count(b(3), 1, 48000) < 0
    assert L.orc_packet_sample_count(b(255, 5), 2, 48000) == 4800


# ---- lib.rs:771-860 parse_packet KATs
@pytest.mark.parametrize("name,count,frames,sizes,payload_off,packet_off", [
    ("single", 1, [1], [11], 1, 12), ("cbr", 2, [1, 6], [5, 5], 1, 11), ("vbr", 2, [2, 6], [4, 6], 2, 12)])
def test_parse_packet(name, count, frames, sizes, payload_off, packet_off):
    pkt = np.array(KATS["packet_" + name], np.uint8)
    fr, sz = np.zeros(48, np.uint32), np.zeros(48, np.uint32)
    po, ko = C.c_uint32(0), C.c_uint32(0)
    assert L.orc_parse_packet(O.ptr(pkt), len(pkt), 0, O.ptr(fr), O.ptr(sz), C.byref(po), C.byref(ko)) == count
    assert list(fr[:count]) == frames and list(sz[:count]) == sizes
    assert (po.value, ko.value) == (payload_off, packet_off)


def test_parse_packet_invalid():
    pkt = np.array(KATS["packet_invalid"], np.uint8)
    fr, sz = np.zeros(48, np.uint32), np.zeros(48, np.uint32)
    assert L.orc_parse_packet(O.ptr(pkt), len(pkt), 0, O.ptr(fr), O.ptr(sz), None, None) < 0


# ---- lib.rs:863-890 test_pcm_soft_clip
def test_pcm_soft_clip():
    s = np.zeros(8, np.float32)
    base = ((np.arange(1024) & 255) * (1.0 / 32.0) - 4.0).astype(np.float32)
    for i in range(0, 1024, 37):
        x = base.copy()
        tail = np.ascontiguousarray(x[i:])
        L.orc_pcm_soft_clip(O.ptr(tail), len(tail), 1, O.ptr(s), 8)
        assert tail.max() <= 1.0 and tail.min() >= -1.0
    for ch in range(1, 9):
        x = base.copy()
        L.orc_pcm_soft_clip(O.ptr(x), 1024, ch, O.ptr(s), 8)
        n = (1024 // ch) * ch
        assert x[:n].max() <= 1.0 and x[:n].min() >= -1.0


def test_cwrsi_single_pulse_closed_form():
    """k_synth_expand decodes single-pulse parts without walking the dimensions: for K = 1 cwrsi
    (pvc.rs:182-284) gives y[i] = +1 for i < n and y[2n-1-i] = -1 otherwise.  Checked for every band size
    the reference's test_pvc uses (pvc.rs:463-469) and every index."""
    for n in [2, 3, 4, 6, 8, 9, 11, 12, 16, 18, 22, 24, 32, 36, 44, 48, 64, 72, 88, 96, 144, 176]:
        for i in range(2 * n):
            y = np.zeros(n, np.int32)
            yy = L.orc_cwrsi(O.ptr(y), n, 1, i)
            w = np.zeros(n, np.int32)
            if i < n:
                w[i] = 1
            else:
                w[2 * n - 1 - i] = -1
            assert np.array_equal(y, w) and yy == 1.0, (n, i)


def _pvq_tables():
    """U(n,k) table and row offsets as the library holds them (opn_tables.h, generated by tools/gen_tables.py)."""
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "opus-native_b200", "csrc", "opn_tables.h")).read()

    def tab(name):
        m = re.search(r"%s\[\d+\] = \{(.*?)\};" % name, src, re.S)
        return [int(x.strip().rstrip("u")) for x in m.group(1).replace("\n", " ").split(",") if x.strip()]

    return tab("OPN_PVQ_U_DATA"), tab("OPN_PVQ_U_ROW")


def test_cwrsi_event_walk_matches_oracle():
    """k_synth_expand (symbols.cuh) walks a part in events: while k < n the run of empty dimensions is found by
    bisection on running row sums of U, with the single unsigned test i - C(k,n) + C(k,n-t-1) < V(n-t-1,k).
    This is the same algorithm in Python on the same tables, checked against the oracle's sequential cwrsi
    (pvc.rs:182-284) for the part shapes of SYNTH-CELT/1, shapes that enter the k >= n regime, and the largest
    (n,k) the reference's own test uses with k < n."""
    U, ROW = _pvq_tables()
    M = 0xFFFFFFFF
    CW = [(0, 0)] * len(U)  # (running row sum up to the previous column, U(k+1,m) - U(k,m)), as upload_tables builds it
    for k in range(15):
        first = ROW[k] + k
        end = ROW[k + 1] + k + 1 if k < 14 else len(U)
        next_end = ROW[k + 2] + k + 2 if k < 13 else len(U)
        acc = 0
        for i in range(first, end):
            m = i - ROW[k]
            up = ROW[k + 1] + m if k < 14 else len(U)
            w = (U[up] - U[i]) & M if (m >= k + 1 and up < next_end) else 0
            CW[i] = (acc, w)
            acc = (acc + U[i]) & M
    # ev_nmax[k]: largest n whose sums C(k,n) + U(k+1,n) stay below 2^32 (upload_tables, opn_kernels.cu)
    last_col = lambda r: (ROW[r + 1] + r if r < 14 else len(U) - 1) - ROW[r]
    NMAX = [0] * 16
    for k in range(14):
        rowsum = 0
        for n in range(k, min(last_col(k), last_col(k + 1)) + 1):
            rowsum += U[ROW[k] + n]
            if n > k:
                if rowsum + U[ROW[k + 1] + n] < (1 << 32):
                    NMAX[k] = n
                else:
                    break
    assert NMAX[4] == 176 and NMAX[9] == 23 and NMAX[11] == 17 and NMAX[12] == 15, NMAX  # (24,9), (18,11), (16,12) of the ladder take the guard

    def walk(n, k, i):
        y = [0] * n
        pos = 0
        events = 0
        while n > 2:
            events += 1
            if k >= n:
                rn = ROW[n]
                p = U[rn + k + 1]
                sg = -1 if i >= p else 0
                i -= p if sg else 0
                k0 = k
                if U[rn + n] > i:
                    k = n
                    while True:
                        k -= 1
                        p = U[ROW[k] + n]
                        if p <= i:
                            break
                else:
                    p = U[rn + k]
                    while p > i:
                        k -= 1
                        p = U[rn + k]
                i -= p
                y[pos] = (k0 - k + sg) ^ sg
                pos += 1
                n -= 1
                continue
            rk, rk1 = ROW[min(k, 14)], ROW[min(k + 1, 14)]
            if n <= NMAX[k]:
                T = n - max(k, 2)
                ic = (i - (CW[rk + n][0] + U[rk + n])) & M
                lo, hi = 0, T
                while lo < hi:
                    mid = (lo + hi) >> 1
                    c, w = CW[rk + n - mid]
                    if ((ic + c) & M) < w:
                        lo = mid + 1
                    else:
                        hi = mid
                if lo:
                    i = (ic + CW[rk + n + 1 - lo][0]) & M
            else:  # the sums could wrap: one dimension, as the reference tests it
                T, lo = 1, 0
                p = U[rk + n]
                if p <= i < U[rk1 + n]:
                    i -= p
                    lo = 1
            pos += lo
            n -= lo
            if lo < T:
                q = U[rk1 + n]
                sg = -1 if i >= q else 0
                i -= q if sg else 0
                k0 = k
                while True:
                    k -= 1
                    p = U[ROW[k] + n]
                    if p <= i:
                        break
                i -= p
                y[pos] = (k0 - k + sg) ^ sg
                pos += 1
                n -= 1
        if n == 2:
            p = 2 * k + 1
            sg = -1 if i >= p else 0
            i -= p if sg else 0
            k0 = k
            k = (i + 1) >> 1
            if k:
                i -= 2 * k - 1
            y[pos] = (k0 - k + sg) ^ sg
            sg = -i
            y[pos + 1] = (k + sg) ^ sg
        return y, events

    def V(n, k):
        u = lambda a, b: U[ROW[min(a, b)] + max(a, b)]
        return u(n, k) + u(n, k + 1)

    rnd = np.random.default_rng(5)
    shapes = [(8, 1), (16, 2), (24, 3), (32, 4), (36, 4), (44, 5), (8, 3), (6, 4), (4, 6), (3, 10), (12, 9),
              (11, 12), (22, 9), (176, 4), (96, 5), (48, 6), (9, 20), (2, 5)]
    for n, k in shapes:
        v = V(n, k)
        idxs = {0, 1, v - 1, v - 2, v // 2} | {int(x) for x in rnd.integers(0, v, 300)}
        for i in idxs:
            y = np.zeros(n, np.int32)
            L.orc_cwrsi(O.ptr(y), n, k, int(i))
            w, events = walk(n, k, int(i))
            assert list(y) == w, (n, k, i)
            if k < n and n <= NMAX[k]:
                assert events <= k + 1, (n, k, i, events)  # one event per pulse-bearing dimension plus the last run
    # the whole ladder of the reference's test_pvc (pvc.rs:462-503), a few indices per shape
    get_pulses = lambda i: i if i < 8 else (8 + (i & 7)) << ((i >> 3) - 1)
    for n, kmax in zip(KATS["pvc_pn"], KATS["pvc_pk_max"]):
        for pseudo in range(1, 41):
            k = get_pulses(pseudo)
            if k > kmax:
                break
            v = V(n, k)
            for i in {0, v - 1, v // 3} | {int(x) for x in rnd.integers(0, v, 5)}:
                y = np.zeros(n, np.int32)
                L.orc_cwrsi(O.ptr(y), n, k, int(i))
                assert list(y) == walk(n, k, int(i))[0], (n, k, i)


def test_sample_from_f32_follows_the_crates_conversions():
    """`Sample::from_f32` (lib.rs:63-107) has no test in the reference; the oracle restates it with Rust's cast
    rules (truncate toward zero, saturate, NaN -> 0) and the crate's own clamp bounds, including the two that look
    odd: the i32 upper bound 2_147_483_647.0 is 2^31 as an f32 (the cast then saturates to i32::MAX) and the
    unsigned types clamp to midpoint + full scale, so full-scale positive input lands on the midpoint."""
    x = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 2.0, -2.0, np.nan, 1.0 / 32768.0, 0.99999, -0.99999, 3.0e-10], np.float32)

    def conv(fmt, dtype):
        out = np.zeros(x.size, dtype)
        assert L.orc_sample_from_f32(fmt, O.ptr(x), O.ptr(out), x.size) == 0
        return out

    assert conv(1, np.int16).tolist() == [0, 32767, -32768, 16384, -16384, 32767, -32768, 0, 1, 32767, -32767, 0]
    assert conv(2, np.int32).tolist() == [0, 2147483647, -2147483648, 1073741824, -1073741824, 2147483647, -2147483648, 0,
                                           65536, 2147462144, -2147462144, 0]
    assert conv(3, np.uint16).tolist() == [32768, 32768, 0, 32768, 16384, 32768, 0, 0, 32768, 32768, 0, 32768]
    assert conv(4, np.uint32).tolist() == [2147483648, 2147483648, 0, 2147483648, 1073741824, 2147483648, 0, 0, 2147483648,
                                            2147483648, 21504, 2147483648]
    assert np.array_equal(conv(0, np.float32)[[0, 1, 2, 3]], x[[0, 1, 2, 3]])
    f64 = conv(5, np.float64)
    assert np.array_equal(f64[~np.isnan(f64)], x[~np.isnan(x)].astype(np.float64)) and np.isnan(f64[7])
    assert L.orc_sample_from_f32(9, O.ptr(x), O.ptr(np.zeros(x.size, np.float64)), x.size) == -1
