"""GPU parity tests: every CUDA entry point of libopusb200 (called through the C ABI) against the
CPU oracle on the same seeded inputs.  Integer/byte/index results must be bit-exact; float PCM must
be within max-abs 1e-5 and SNR > 100 dB (north_star), and is in fact expected to be bit-identical
because the kernels keep the reference's operation order and are built with -fmad=false."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

import opus_native_b200 as opn
import oracle_lib as O

pytestmark = pytest.mark.gpu
KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
PCM_TOL = 1e-5  # north_star: max-abs PCM error
SYNTH = dict(bitstream=opn.BITSTREAM_SYNTH_CELT_1)  # CELT frames only decode after this explicit opt-in (opusb200.h)


def snr_db(want, got):
    want = want.astype(np.float64)
    err = ((want - got.astype(np.float64)) ** 2).sum()
    return 200.0 if err == 0 else 10 * np.log10((want ** 2).sum() / err)


def assert_pcm(want, got, what=""):
    assert np.all(np.isfinite(got)), what
    assert np.abs(want.astype(np.float64) - got).max() <= PCM_TOL, what
    assert snr_db(want, got) > 100.0, what


# ------------------------------------------------------------------ range decoder (a1-a9)
def test_rangedec_simple_uint_bits_kat():
    """range_coder/mod.rs:191-263 replayed by one warp on the GPU: every symbol, the per-symbol
    tell_frac and the final coder state match; the stream is the reference's 497192-byte KAT."""
    ops, vals = [], []
    for ft in range(2, 1024):
        ops += [(opn.OP_UINT, ft, 0)] * ft
        vals += list(range(ft))
    for ftb in range(1, 16):
        ops += [(opn.OP_BITS, ftb, 0)] * (1 << ftb)
        vals += list(range(1 << ftb))
    ops = np.array(ops, opn.OP_DTYPE)
    vals = np.array(vals, np.uint32)
    buf, tf, range_bytes, final_tf, err = opn.enc_run_script(1 << 20, ops, vals)
    assert err == 0 and range_bytes == KATS["simple_uint_bits"]["range_bytes"]
    assert final_tf / 8.0 == KATS["simple_uint_bits"]["tell_frac_over_8"]
    out, _ = opn.op_rangedec_script(buf, [0], [len(buf)], ops)
    assert np.array_equal(out[0]["value"], vals)
    assert np.array_equal(out[0]["tell_frac"], tf)
    want, _ = O.dec_run_script(buf, ops)
    assert np.array_equal(out[0], want)


def test_rangedec_encoder_prefers_range_coder_data_kat():
    """range_coder/mod.rs:271-298"""
    ops = np.array([(opn.OP_BITS, 7, 0)] + [(opn.OP_UINT, ft, 0) for ft in (2, 3, 4, 5, 6, 7)], opn.OP_DTYPE)
    buf, _, _, _, _ = opn.enc_run_script(2, ops, [0x55, 1, 1, 1, 1, 2, 6])
    out, _ = opn.op_rangedec_script(buf, [0], [2], ops)
    assert out[0]["value"].tolist() == [0x05, 1, 1, 1, 1, 2, 6]


def _random_script(rnd, n_ops, with_pulses=True):
    ops, vals, ys = [], [], []
    for _ in range(n_ops):
        kind = int(rnd.integers(0, 8 if with_pulses else 7))
        if kind == 0:
            ft = int(rnd.integers(2, 2 ** int(rnd.integers(2, 33)) - 1))
            ops.append((opn.OP_UINT, ft, 0)); vals.append(int(rnd.integers(0, ft)))
        elif kind == 1:
            nb = int(rnd.integers(1, 26))
            ops.append((opn.OP_BITS, nb, 0)); vals.append(int(rnd.integers(0, 1 << nb)))
        elif kind == 2:
            ops.append((opn.OP_BIT_LOGP, int(rnd.integers(1, 16)), 0)); vals.append(int(rnd.integers(0, 2)))
        elif kind == 3:
            ops.append((opn.OP_ICDF, 0, 5)); vals.append(int(rnd.integers(0, 6)))
        elif kind == 4:
            decay = int(rnd.integers(5000, 16000))
            ops.append((opn.OP_LAPLACE, O.lib().orc_laplace_start_freq(decay), decay))
            vals.append(int(rnd.integers(-12, 13)) & 0xFFFFFFFF)
        elif kind == 5:
            ops.append((opn.OP_BIT_VIA_DECODE, int(rnd.integers(1, 16)), 0)); vals.append(int(rnd.integers(0, 2)))
        elif kind == 6:
            ops.append((opn.OP_BIT_VIA_DECODE_BIN, int(rnd.integers(1, 16)), 0)); vals.append(int(rnd.integers(0, 2)))
        else:
            i = int(rnd.integers(0, 22))
            n, kmax = KATS["pvc_pn"][i], KATS["pvc_pk_max"][i]
            k = int(rnd.integers(1, kmax + 1))
            y = np.zeros(n, np.int32)
            O.lib().orc_cwrsi(O.ptr(y), n, k, int(rnd.integers(0, O.lib().orc_pvq_v(n, k))))
            ops.append((opn.OP_PULSES, n, k)); vals.append(0); ys.extend(y.tolist())
    return np.array(ops, opn.OP_DTYPE), vals, ys


ICDF_POOL = np.array([30, 22, 15, 8, 3, 0], np.uint8)


def test_rangedec_random_scripts_all_ops():
    """Many packets, one script per launch: values, tell_frac, rng and pulse vectors, bit-exact.
    Also covers test_random_data / test_compatibility / test_laplace (mod.rs:301-570) semantics:
    decoder tell_frac equals the encoder's after every symbol."""
    rnd = np.random.default_rng(11)
    for trial in range(6):
        ops, _, _ = _random_script(rnd, 150)
        # same script shape, different symbol values per packet
        bufs, tfs = [], []
        n_pk = 40
        for p in range(n_pk):
            vals, ys = [], []
            for op, a, b in ops:
                if op == opn.OP_UINT: vals.append(int(rnd.integers(0, a)))
                elif op == opn.OP_BITS: vals.append(int(rnd.integers(0, 1 << a)))
                elif op == opn.OP_ICDF: vals.append(int(rnd.integers(0, 6)))
                elif op == opn.OP_LAPLACE: vals.append(int(rnd.integers(-12, 13)) & 0xFFFFFFFF)
                elif op == opn.OP_PULSES:
                    y = np.zeros(a, np.int32)
                    O.lib().orc_cwrsi(O.ptr(y), int(a), int(b), int(rnd.integers(0, O.lib().orc_pvq_v(int(a), int(b)))))
                    vals.append(0); ys.extend(y.tolist())
                else: vals.append(int(rnd.integers(0, 2)))
            buf, tf, rb, ftf, err = opn.enc_run_script(1275, ops, vals, ICDF_POOL, ys or None)
            assert err == 0
            bufs.append(buf); tfs.append(tf)
        arena = np.concatenate(bufs)
        offsets = np.arange(n_pk, dtype=np.uint32) * 1275
        lens = np.full(n_pk, 1275, np.uint32)
        ystride = int(sum(a for op, a, b in ops if op == opn.OP_PULSES))
        out, y = opn.op_rangedec_script(arena, offsets, lens, ops, ICDF_POOL, y_stride=ystride)
        for p in range(n_pk):
            want, wy = O.dec_run_script(bufs[p], ops, ICDF_POOL, y_cap=ystride)
            assert np.array_equal(out[p], want), (trial, p)
            assert np.array_equal(y[p, :len(wy)], wy)
            assert np.array_equal(out[p]["tell_frac"], tfs[p])


def test_rangedec_garbage_and_truncated_packets():
    """Corrupt input: random bytes, short and empty buffers, storage shrink.  The decoder must do
    exactly what the reference does (zero-extension, decode_uint saturation, decoder.rs:86-104,255-259)."""
    rnd = np.random.default_rng(5)
    base_ops, _, _ = _random_script(rnd, 120)
    shrink_ops = np.concatenate([base_ops[:60], np.array([(opn.OP_SHRINK, 3, 0), (opn.OP_TELL, 0, 0)], opn.OP_DTYPE), base_ops[60:]])
    # SHRINK by 3 needs len >= 3 (a usize underflow panic in the reference otherwise)
    for ops, lens in ((base_ops, [0, 1, 2, 3, 5, 8, 13, 40, 100, 300, 1275, 17]), (shrink_ops, [3, 4, 5, 8, 13, 40, 100, 300, 1275, 17])):
        lens = np.array(lens, np.uint32)
        offsets = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint32)
        arena = rnd.integers(0, 256, int(lens.sum()) + 4).astype(np.uint8)
        ystride = int(sum(a for op, a, b in ops if op == opn.OP_PULSES))
        out, y = opn.op_rangedec_script(arena, offsets, lens, ops, ICDF_POOL, y_stride=ystride)
        for p in range(len(lens)):
            want, wy = O.dec_run_script(arena[offsets[p]:offsets[p] + lens[p]], ops, ICDF_POOL, y_cap=ystride)
            assert np.array_equal(out[p], want), p
            assert np.array_equal(y[p, :len(wy)], wy), p


# ------------------------------------------------------------------ PVQ (a10, a11)
def test_cwrsi_all_band_sizes():
    """celt/pvc.rs:453-503 test_pvc on the GPU: for every (N, K) of the ladder, evenly spaced codeword
    indices -> pulse vector, sum|y| == K, yy exact, and identical to the oracle's cwrsi."""
    def get_pulses(i):
        return i if i < 8 else (8 + (i & 7)) << ((i >> 3) - 1)

    L = O.lib()
    for n, kmax in zip(KATS["pvc_pn"], KATS["pvc_pk_max"]):
        for pseudo in range(1, 41):
            k = get_pulses(pseudo)
            if k > kmax:
                break
            nc = L.orc_pvq_v(n, k)
            idx = np.unique(np.linspace(0, nc - 1, 48).astype(np.uint64)).astype(np.uint32)
            # one packet per index, each holding a single decode_uint(V) == the codeword index
            eop = np.array([(opn.OP_UINT, nc, 0)], opn.OP_DTYPE)
            bufs = [opn.enc_run_script(16, eop, [int(i)])[0] for i in idx]
            arena = np.concatenate(bufs)
            offsets = np.arange(len(idx), dtype=np.uint32) * 16
            out, y = opn.op_rangedec_script(arena, offsets, np.full(len(idx), 16, np.uint32),
                                            np.array([(opn.OP_PULSES, n, k)], opn.OP_DTYPE), y_stride=n)
            want = np.zeros(n, np.int32)
            for j, i in enumerate(idx):
                yy = L.orc_cwrsi(O.ptr(want), n, k, int(i))
                assert np.array_equal(y[j], want), (n, k, int(i))
                assert int(np.abs(y[j]).sum()) == k
                assert out[j, 0]["value"] == np.float32(yy).view(np.uint32)
                assert L.orc_icwrs(O.ptr(y[j].copy()), n) == i


def test_cwrsi_event_walk_all_band_sizes():
    """The product path's cwrsi (cwrsi_events in csrc/symbols.cuh: the function k_synth_expand calls, reached here through
    OPN_OP_PULSES_EVENTS) on the whole (N, K) ladder of celt/pvc.rs:462-503, including the shapes whose 32-bit running
    sums would wrap ((24,9), (18,11), (16,12): the walk falls back to the reference's one-dimension step there)."""
    def get_pulses(i):
        return i if i < 8 else (8 + (i & 7)) << ((i >> 3) - 1)

    L = O.lib()
    rnd = np.random.default_rng(11)
    for n, kmax in zip(KATS["pvc_pn"], KATS["pvc_pk_max"]):
        for pseudo in range(1, 41):
            k = get_pulses(pseudo)
            if k > kmax:
                break
            nc = L.orc_pvq_v(n, k)
            idx = np.unique(np.concatenate([np.linspace(0, nc - 1, 40).astype(np.uint64), rnd.integers(0, nc, 24).astype(np.uint64)])).astype(np.uint32)
            eop = np.array([(opn.OP_UINT, nc, 0)], opn.OP_DTYPE)
            arena = np.concatenate([opn.enc_run_script(16, eop, [int(i)])[0] for i in idx])
            offsets = np.arange(len(idx), dtype=np.uint32) * 16
            lens = np.full(len(idx), 16, np.uint32)
            out, y = opn.op_rangedec_script(arena, offsets, lens, np.array([(opn.OP_PULSES_EVENTS, n, k)], opn.OP_DTYPE), y_stride=n)
            ref, yref = opn.op_rangedec_script(arena, offsets, lens, np.array([(opn.OP_PULSES, n, k)], opn.OP_DTYPE), y_stride=n)
            assert np.array_equal(y, yref) and np.array_equal(out, ref), (n, k)  # both device implementations agree
            want = np.zeros(n, np.int32)
            for j, i in enumerate(idx):
                yy = L.orc_cwrsi(O.ptr(want), n, k, int(i))
                assert np.array_equal(y[j], want), (n, k, int(i))
                assert out[j, 0]["value"] == np.float32(yy).view(np.uint32)


# ------------------------------------------------------------------ IMDCT + TDAC (a12-a14, a17)
@pytest.mark.parametrize("shift", [0, 1, 2, 3])
def test_imdct_tdac_long_blocks(shift):
    rnd = np.random.default_rng(100 + shift)
    n2 = 960 >> shift
    rows = 37
    coefs = (rnd.uniform(-1, 1, (rows, n2)) / 32).astype(np.float32)
    coefs[:, (100 * 8) >> shift:] = 0  # above band 21
    coefs[0] = 0
    coefs[1] = (rnd.uniform(-1, 1, n2) * 32768.0).astype(np.float32)  # mdct.rs:711-714 scale
    out = np.zeros((rows, n2 + 60), np.float32)
    out[:, :60] = rnd.uniform(-1, 1, (rows, 60)).astype(np.float32)
    want = out.copy()
    for r in range(rows):
        O.mdct_backward(coefs[r], want[r], shift)
    got = opn.op_imdct_tdac(coefs, out.copy(), shift)
    assert np.array_equal(got, want)  # bit-exact
    assert_pcm(want[2:], got[2:])


def test_imdct_matches_f64_definition():
    """celt/mdct.rs:672-701 check_inv: SNR vs the O(n^2) f64 IMDCT (> 60 dB in the reference)."""
    for shift, nfft in [(2, 480), (1, 960), (0, 1920)]:  # N=240 has no un-windowed span with overlap 120
        rnd = np.random.default_rng(42)
        x = ((rnd.integers(0, 32768, nfft // 2) - 16384) * 32768.0 / nfft).astype(np.float32)
        out = np.zeros((1, nfft // 2 + 60), np.float32)
        got = opn.op_imdct_tdac(x[None, :].copy(), out, shift)[0]
        i = np.arange(nfft)[:, None]
        k = np.arange(nfft // 2)[None, :]
        full = np.cos(2 * np.pi * (i + 0.5 + 0.25 * nfft) * (k + 0.5) / nfft) @ x.astype(np.float64)
        # out[60 + j] = y[N/4 + j] for the un-windowed part j in [60, N/2 - 60)  
        j = np.arange(60, nfft // 2 - 60)
        want = full[nfft // 4 + j]
        assert 10 * np.log10((want ** 2).sum() / ((want - got[60 + j]) ** 2).sum()) > 100.0


@pytest.mark.parametrize("blocks", [2, 4, 8])
def test_imdct_tdac_short_blocks(blocks):
    """Transient frames: B interleaved 120-bin blocks (stride = B), chained TDAC (mdct.rs:186,197-198)."""
    rnd = np.random.default_rng(blocks)
    rows = 19
    coefs = (rnd.uniform(-1, 1, (rows, 120 * blocks)) / 32).astype(np.float32)
    out = np.zeros((rows, 120 * blocks + 60), np.float32)
    out[:, :60] = rnd.uniform(-1, 1, (rows, 60)).astype(np.float32)
    want = out.copy()
    for r in range(rows):
        for b in range(blocks):
            O.lib().orc_mdct_backward(O.ptr(coefs[r][b:]), O.ptr(want[r][120 * b:]), O.ptr(O.window()), 120, 3, blocks)
    got = opn.op_imdct_tdac(coefs, out.copy(), 3, blocks=blocks)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ comb filter (a15, a16)
def test_comb_filter_golden_vectors():
    """celt/comb_filter/mod.rs:227-270: TEST_VECTOR1 (out of place) and TEST_VECTOR2 (in place)."""
    p = KATS["comb_params"]
    size, n = p["SIZE"], p["N"]
    x = np.arange(size, dtype=np.float32)[None, :].copy()
    params = [[p["T0"], p["T1"], 0, 0]]
    gains = [[p["G0"], p["G1"]]]
    y = opn.op_comb_filter(np.zeros_like(x), x, size - n, n, params, gains, p["OVERLAP"])
    v1 = np.array(KATS["comb_test_vector1"], np.float32)
    assert np.all(np.abs(1.0 - y[0, size - n:] / v1) < 1e-5)
    assert np.array_equal(y[0, size - n:], v1)
    y2 = opn.op_comb_filter_inplace(x.copy(), size - n, n, params, gains, p["OVERLAP"])
    v2 = np.array(KATS["comb_test_vector2"], np.float32)
    assert np.all(np.abs(1.0 - y2[0, size - n:] / v2) < 1e-5)


def _comb_cases(rnd, rows):
    params, gains = [], []
    for r in range(rows):
        t0, t1 = int(rnd.integers(0, 1023)), int(rnd.integers(0, 1023))
        if r % 5 == 0: t1 = t0
        if r % 7 == 0: t0, t1 = 15, 15
        g0, g1 = float(rnd.integers(0, 9)) * 0.09375, float(rnd.integers(0, 9)) * 0.09375
        if r % 11 == 0: g1 = g0
        params.append([t0, t1, int(rnd.integers(0, 3)), int(rnd.integers(0, 3))])
        gains.append([g0, g1])
    params[0], gains[0] = [40, 40, 1, 1], [0.5, 0.5]      # unchanged filter -> overlap = 0
    params[1], gains[1] = [100, 300, 0, 2], [0.0, 0.0]    # both gains zero -> untouched / copy
    params[2], gains[2] = [100, 300, 0, 2], [0.75, 0.0]   # g1 = 0 -> only the cross-fade part
    params[3], gains[3] = [0, 0, 0, 0], [0.0, 0.75]       # periods below COMBFILTER_MINPERIOD
    params[4], gains[4] = [1022, 15, 2, 0], [0.75, 0.75]
    return np.array(params, np.int32), np.array(gains, np.float32)


@pytest.mark.parametrize("n,overlap", [(960, 120), (480, 120), (120, 120), (240, 64), (64, 8), (961, 120), (2, 0)])
def test_comb_filter_inplace_random(n, overlap):
    rnd = np.random.default_rng(n + overlap)
    rows, off = 48, 1024
    params, gains = _comb_cases(rnd, rows)
    y = rnd.uniform(-0.5, 0.5, (rows, off + n)).astype(np.float32)
    want = y.copy()
    for r in range(rows):
        O.comb_filter_inplace(want[r], off, int(params[r, 0]), int(params[r, 1]), n, float(gains[r, 0]), float(gains[r, 1]),
                              int(params[r, 2]), int(params[r, 3]), overlap)
    got = opn.op_comb_filter_inplace(y.copy(), off, n, params, gains, overlap)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,overlap", [(960, 120), (120, 120), (64, 8)])
def test_comb_filter_out_of_place_random(n, overlap):
    rnd = np.random.default_rng(n)
    rows, off = 48, 1024
    params, gains = _comb_cases(rnd, rows)
    x = rnd.uniform(-0.5, 0.5, (rows, off + n)).astype(np.float32)
    y0 = rnd.uniform(-0.5, 0.5, (rows, off + n)).astype(np.float32)
    want = y0.copy()
    for r in range(rows):
        O.comb_filter(want[r], off, x[r], off, int(params[r, 0]), int(params[r, 1]), n, float(gains[r, 0]), float(gains[r, 1]),
                      int(params[r, 2]), int(params[r, 3]), overlap)
    got = opn.op_comb_filter(y0.copy(), x, off, n, params, gains, overlap)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ soft clip (a19)
@pytest.mark.parametrize("channels", [1, 2, 3])
def test_pcm_soft_clip(channels):
    rnd = np.random.default_rng(channels)
    rows, n = 33, 960
    pcm = (rnd.normal(0, 0.9, (rows, n * channels))).astype(np.float32)
    pcm[0] = ((np.arange(n * channels) & 255) * (1.0 / 32.0) - 4.0).astype(np.float32)  # lib.rs:868-870
    mem = rnd.uniform(-0.2, 0.2, (rows, channels)).astype(np.float32)
    mem[1] = 0
    want, wmem = pcm.copy(), mem.copy()
    for r in range(rows):
        O.lib().orc_pcm_soft_clip(O.ptr(want[r]), n * channels, channels, O.ptr(wmem[r]), channels)
    got_mem = mem.copy()
    got = opn.op_pcm_soft_clip(pcm.copy(), n * channels, channels, got_mem)
    assert np.array_equal(got, want) and np.array_equal(got_mem, wmem)
    # lib.rs:874-877 asserts [-1, 1] for the reference's own sawtooth input (row 0).  For general input
    # the reference's search-loop quirk (see oracle/packet.c) can overshoot; parity is what is checked.
    assert got[0].max() <= 1.0 and got[0].min() >= -1.0


# ------------------------------------------------------------------ SYNTH-CELT/1 symbol kernel
@pytest.mark.parametrize("lm,channels,pkt_bytes", [(3, 2, 160), (3, 1, 100), (2, 2, 130), (1, 2, 100), (0, 2, 80), (0, 1, 48)])
def test_synth_symbols(lm, channels, pkt_bytes):
    n = 96
    pk = opn.synth_fill(1000, n, 3, 1, lm, channels, pkt_bytes, transient_permille=200)[0]
    payload = np.ascontiguousarray(pk[:, 1:])
    side, y, coef = opn.op_synth_symbols(payload.reshape(-1), np.arange(n, dtype=np.uint32) * (pkt_bytes - 1),
                                         np.full(n, pkt_bytes - 1, np.uint32), lm, channels)
    for s in range(n):
        w, wy, wc, _ = O.SynthStream(lm, channels).decode(payload[s])
        for f in ("silence", "postfilter", "octave", "period", "gain_idx", "tapset", "transient", "intra", "final_rng", "tell_frac", "n_pulses"):
            assert side[s][f] == getattr(w, f), (s, f)
        assert np.array_equal(side[s]["coarse"], np.ctypeslib.as_array(w.coarse))
        assert np.array_equal(side[s]["fine"], np.ctypeslib.as_array(w.fine))
        assert np.array_equal(y[s].reshape(-1), wy)
        assert np.array_equal(coef[s].reshape(-1).view(np.uint32), wc.view(np.uint32))


@pytest.mark.parametrize("lm,channels", [(3, 2), (3, 1), (2, 2), (0, 2), (0, 1)])
def test_synth_symbols_garbage_and_truncated_payloads(lm, channels):
    """Corrupt input through the PRODUCT kernels (k_synth_rangedec's uint_precomputed / div_small_quotient / div_magic
    and k_synth_expand), not the generic script decoder: random bytes and truncated real payloads of 2..160 bytes must
    give exactly what the reference's arithmetic gives (zero extension past the buffer, decode_uint saturating at ft-1,
    decoder.rs:86-104, 255-259) -- every symbol, the final range, tell_frac, pulses and coefficients."""
    rnd = np.random.default_rng(100 * lm + channels)
    lens = np.concatenate([np.arange(2, 161, 3), rnd.integers(2, 161, 80)]).astype(np.uint32)
    n = len(lens)
    offsets = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint32)
    arena = rnd.integers(0, 256, int(lens.sum()) + 8).astype(np.uint8)
    # every third packet: a real payload cut short instead of noise (valid prefix, then zero extension)
    real = opn.synth_fill(7, n, 0, 1, lm, channels, 161, transient_permille=300)[0][:, 1:]
    for s in range(0, n, 3):
        arena[offsets[s]:offsets[s] + lens[s]] = real[s, :lens[s]]
    arena[offsets[1]] &= 0x7F  # make sure noise packets do not all start with the silence flag pattern
    side, y, coef = opn.op_synth_symbols(arena, offsets, lens, lm, channels)
    for s in range(n):
        w, wy, wc, _ = O.SynthStream(lm, channels).decode(arena[offsets[s]:offsets[s] + lens[s]])
        for f in ("silence", "postfilter", "octave", "period", "gain_idx", "tapset", "transient", "intra", "final_rng", "tell_frac", "n_pulses"):
            assert side[s][f] == getattr(w, f), (s, int(lens[s]), f)
        assert np.array_equal(side[s]["coarse"], np.ctypeslib.as_array(w.coarse)), s
        assert np.array_equal(side[s]["fine"], np.ctypeslib.as_array(w.fine)), s
        assert np.array_equal(y[s].reshape(-1), wy), (s, int(lens[s]))
        assert np.array_equal(coef[s].reshape(-1).view(np.uint32), wc.view(np.uint32)), s


# ------------------------------------------------------------------ SYNTH-CELT/2: allocation-driven frames (f1, first slice)
CELT2 = dict(bitstream=opn.BITSTREAM_SYNTH_CELT_2)


def _celt2_check(side, y, coef, payloads, lm, channels):
    for s, payload in enumerate(payloads):
        w, _, wy, wc = O.celt2_decode_symbols(payload, lm, channels)
        for f in opn.CELT2_SIDE_DTYPE.names:
            v = getattr(w, f)
            v = np.ctypeslib.as_array(v) if hasattr(v, "__len__") else v
            assert np.array_equal(side[s][f], v), (s, len(payload), f, side[s][f], v)
        assert np.array_equal(y[s].reshape(-1), wy), (s, len(payload))
        assert np.array_equal(coef[s].reshape(-1).view(np.uint32), wc.view(np.uint32)), s


@pytest.mark.parametrize("lm,channels,pkt_bytes", [(3, 2, 160), (3, 1, 100), (2, 2, 130), (2, 1, 64), (1, 2, 100), (1, 1, 60), (0, 2, 80),
                                                   (0, 1, 48), (3, 2, 48), (3, 2, 300)])
def test_celt2_symbols(lm, channels, pkt_bytes):
    """The device frame decode of SYNTH-CELT/2 (k_celt2_rangedec: band boosts, trim, compute_allocation on the mode's tables,
    fine bits, theta splits with bitexact_cos / bitexact_log2tan, leaf sizes from the pulse cache; then the part-list
    expansion) against the oracle: the whole side record -- allocation vector, fine bits and priorities, coded bands,
    intensity / dual-stereo, balance, counts of leaves, pulses, splits, theta checksum, final range, tell_frac -- the
    pulses and the coefficients, bit for bit.  Shapes (n, K) are computed per frame on the device."""
    n = 64
    pk = opn.celt2_fill(500, n, 2, 1, lm, channels, pkt_bytes, transient_permille=250)[0]
    payload = np.ascontiguousarray(pk[:, 1:])
    side, y, coef = opn.op_celt2_symbols(payload.reshape(-1), np.arange(n, dtype=np.uint32) * (pkt_bytes - 1),
                                         np.full(n, pkt_bytes - 1, np.uint32), lm, channels)
    _celt2_check(side, y, coef, [payload[s] for s in range(n)], lm, channels)
    assert side["n_parts"].min() > 0 and side["n_splits"].sum() >= 0


@pytest.mark.parametrize("lm,channels", [(3, 2), (2, 1), (0, 2)])
def test_celt2_symbols_garbage_and_truncated_payloads(lm, channels):
    """Random bytes and truncated payloads of 2..200 bytes: the allocation runs on whatever tell_frac the garbage produces
    (budgets that go negative, bands that get nothing, uint saturation) and must still match the oracle exactly."""
    rnd = np.random.default_rng(7 + lm)
    lens = np.concatenate([np.arange(2, 201, 5), rnd.integers(2, 201, 40)]).astype(np.uint32)
    n = len(lens)
    offsets = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint32)
    arena = rnd.integers(0, 256, int(lens.sum()) + 8).astype(np.uint8)
    real = opn.celt2_fill(9, n, 0, 1, lm, channels, 201, transient_permille=300)[0][:, 1:]
    for s in range(0, n, 2):
        arena[offsets[s]:offsets[s] + lens[s]] = real[s, :lens[s]]
    side, y, coef = opn.op_celt2_symbols(arena, offsets, lens, lm, channels)
    _celt2_check(side, y, coef, [arena[offsets[s]:offsets[s] + lens[s]] for s in range(n)], lm, channels)


@pytest.mark.parametrize("lm,channels,pkt_bytes", [(3, 2, 160), (2, 1, 80), (0, 2, 80)])
def test_celt2_batch_decode_chain(lm, channels, pkt_bytes):
    """SYNTH-CELT/2 through the batch entry point: the frame kernel expands each stream's part list, then IMDCT, TDAC and the
    comb post-filter as for SYNTH-CELT/1; PCM and final_range against the oracle's chained decode, with a lost packet."""
    ns, nfr, nf = 40, 8, 120 << lm
    packets = opn.celt2_fill(70, ns, 0, nfr, lm, channels, pkt_bytes, transient_permille=150)
    dec = opn.BatchDecoder(ns, opn.DecoderConfiguration(48000, channels, 0), **CELT2)
    oracle = [O.Celt2Stream(lm, channels) for _ in range(ns)]
    offsets = np.arange(ns, dtype=np.uint32) * pkt_bytes
    for f in range(nfr):
        lens = np.full(ns, pkt_bytes, np.uint32)
        if f == 3:
            lens[5] = 0
        pcm = np.zeros((ns, nf * channels), np.float32)
        res = dec.decode_float(packets[f].reshape(-1), offsets, lens, pcm, nf)
        assert np.all(res == nf)
        rng = dec.final_ranges()
        for s in range(ns):
            side, want = oracle[s].decode(packets[f, s, 1:] if lens[s] else b"")
            assert np.array_equal(pcm[s], want), (f, s)
            assert rng[s] == (side.final_rng if lens[s] else 0)
            assert_pcm(want, pcm[s], (f, s))


@pytest.mark.parametrize("channels,fs", [(1, 48000), (2, 48000), (2, 24000), (1, 16000), (2, 8000)])
def test_smooth_fade(channels, fs):
    """smooth_fade_into_in1 / _in2 (decoder.rs:833-865), the cross-fade of a mode transition: rows of interleaved samples
    against orc_smooth_fade, bit for bit; overlap = 2.5 ms at the decoder's rate (decoder.rs:731-788), samples after it
    untouched."""
    rnd = np.random.default_rng(channels * 7 + fs)
    rows, overlap = 37, fs // 400
    n = overlap * channels + 11
    in1 = (rnd.standard_normal((rows, n)) * 0.7).astype(np.float32)
    in2 = (rnd.standard_normal((rows, n)) * 0.7).astype(np.float32)
    got = opn.op_smooth_fade(in1, in2, overlap, channels, fs)
    want = in1.copy()
    for r in range(rows):
        O.lib().orc_smooth_fade(O.ptr(in1[r]), O.ptr(in2[r]), O.ptr(want[r]), overlap, channels, fs)
    assert np.array_equal(got, want)
    assert np.array_equal(got[:, overlap * channels:], in1[:, overlap * channels:])
    with pytest.raises(opn.OpusError):
        opn.op_smooth_fade(in1, in2, overlap, channels, 44100)


def test_bitexact_trig_checksums():
    """bitexact_cos / bitexact_log2tan on the device against the reference's own checksums
    (src/math.rs:237-298) and, value by value, against the oracle."""
    i = np.arange(64, 16321, dtype=np.int32)
    c, _ = opn.op_bitexact_trig(x=i.astype(np.int16))
    c = c.astype(np.int64)
    assert (int(c[0]), int(c[-1]), int(c[8192 - 64])) == (32767, 200, 23171)
    assert int(np.bitwise_xor.reduce(c * i)) == 89408644
    d = np.diff(np.concatenate([[32767], c])) * -1
    assert (int(d.max()), int(d.min())) == (5, 0)
    L = O.lib()
    assert np.array_equal(c, [L.orc_bitexact_cos(int(v)) for v in i])
    j = np.arange(64, 8193, dtype=np.int32)
    mid, _ = opn.op_bitexact_trig(x=j.astype(np.int16))
    side, _ = opn.op_bitexact_trig(x=(16384 - j).astype(np.int16))
    _, q = opn.op_bitexact_trig(isin=mid.astype(np.int32), icos=side.astype(np.int32))
    _, qr = opn.op_bitexact_trig(isin=side.astype(np.int32), icos=mid.astype(np.int32))
    assert np.array_equal(q, -qr)
    q = q.astype(np.int64)
    assert int(np.bitwise_xor.reduce(q * j)) == 15821257
    d = np.diff(np.concatenate([[15059], q])) * -1
    assert (int(d.max()), int(d.min())) == (61, -2)
    _, k = opn.op_bitexact_trig(isin=[32767, 30274, 23171], icos=[200, 12540, 23171])
    assert list(k) == [15059, 2611, 0]


# ------------------------------------------------------------------ batch pipeline
def _oracle_chain(packets, lm, channels, apply_comb=True):
    """packets [frames, streams, bytes] -> pcm [frames, streams, nf*C], final_rng [frames, streams]"""
    nfr, ns, _ = packets.shape
    nf = 120 << lm
    pcm = np.zeros((nfr, ns, nf * channels), np.float32)
    rng = np.zeros((nfr, ns), np.uint32)
    for s in range(ns):
        st = O.SynthStream(lm, channels, apply_comb)
        for f in range(nfr):
            side, _, _, p = st.decode(packets[f, s, 1:])
            pcm[f, s], rng[f, s] = p, side.final_rng
    return pcm, rng


@pytest.mark.parametrize("lm,channels,pkt_bytes,postfilter", [(3, 2, 160, True), (3, 2, 160, False), (3, 1, 100, True),
                                                               (2, 2, 130, True), (1, 2, 100, True), (0, 2, 80, True)])
def test_batch_decode_chain(lm, channels, pkt_bytes, postfilter):
    ns, nfr, nf = 48, 12 if lm == 3 else 20, 120 << lm
    packets = opn.synth_fill(7, ns, 0, nfr, lm, channels, pkt_bytes, transient_permille=150)
    want, want_rng = _oracle_chain(packets, lm, channels, postfilter)
    dec = opn.BatchDecoder(ns, opn.DecoderConfiguration(48000, channels, 0), postfilter=postfilter, **SYNTH)
    offsets = np.arange(ns, dtype=np.uint32) * pkt_bytes
    lens = np.full(ns, pkt_bytes, np.uint32)
    exact = True
    for f in range(nfr):
        pcm = np.zeros((ns, nf * channels), np.float32)
        res = dec.decode_float(packets[f].reshape(-1), offsets, lens, pcm, nf)
        assert np.all(res == nf)
        assert_pcm(want[f], pcm, f"frame {f}")
        exact &= np.array_equal(pcm, want[f])
        assert np.array_equal(dec.final_ranges(), want_rng[f])
    assert exact, "PCM within tolerance but not bit-identical to the oracle"


@pytest.mark.parametrize("channels", [2, 1])
def test_batch_frame_size_changes_mid_stream(channels):
    """A stream may change its frame size from packet to packet (decoder.rs:329-341 re-reads the TOC every
    call).  The PCM ring then holds frames of mixed sizes: frames straddle the ring end, and the post-filter
    history of a long frame spans several short ones (kernel 1's split store, kernel 2's three-piece load)."""
    ns = 40
    lms = [3, 0, 1, 3, 2, 0, 0, 3, 1, 2, 3, 3, 0, 2, 3, 1, 1, 3, 0, 3, 3, 2]
    pkt_bytes = {0: 80, 1: 100, 2: 130, 3: 160}
    if channels == 1:
        pkt_bytes = {0: 48, 1: 60, 2: 80, 3: 100}
    oracle = [O.SynthStream(3, channels) for _ in range(ns)]
    dec = opn.BatchDecoder(ns, opn.DecoderConfiguration(48000, channels, 0), **SYNTH)
    exact = True
    for f, lm in enumerate(lms):
        nf, pb = 120 << lm, pkt_bytes[lm]
        pk = opn.synth_fill(31, ns, f, 1, lm, channels, pb, transient_permille=200)[0]
        pcm = np.zeros((ns, nf * channels), np.float32)
        res = dec.decode_float(pk.reshape(-1), np.arange(ns, dtype=np.uint32) * pb, np.full(ns, pb, np.uint32), pcm, nf)
        assert np.all(res == nf), (f, lm, res)
        rng = dec.final_ranges()
        for s in range(ns):
            oracle[s].lm = lm
            side, _, _, want = oracle[s].decode(pk[s, 1:])
            assert rng[s] == side.final_rng
            assert_pcm(want, pcm[s], f"frame {f} lm {lm} stream {s}")
            exact &= np.array_equal(want, pcm[s])
    assert exact, "PCM within tolerance but not bit-identical to the oracle"


def test_batch_submit_wait_two_calls_in_flight():
    """OPN_FLAG_SUBMIT_ONLY / opn_batch_wait: the next call is submitted before the previous one's PCM has
    been waited for; results equal the synchronous path (and the oracle) frame by frame."""
    lm, channels, pkt_bytes, ns, nfr, nf = 3, 2, 160, 2304, 7, 960  # > 1024 streams: several chunks per call
    packets = opn.synth_fill(1234, ns, 0, nfr, lm, channels, pkt_bytes, transient_permille=100)
    offsets = np.arange(ns, dtype=np.uint32) * pkt_bytes
    lens = np.full(ns, pkt_bytes, np.uint32)
    sync = opn.BatchDecoder(ns, opn.DecoderConfiguration(48000, channels, 0), **SYNTH)
    want = []
    for f in range(nfr):
        pcm = np.zeros((ns, nf * channels), np.float32)
        sync.decode_float(packets[f].reshape(-1), offsets, lens, pcm, nf)
        want.append(pcm)
    oracle = [O.SynthStream(lm, channels) for _ in range(8)]
    for f in range(nfr):
        for s in range(8):
            assert np.array_equal(oracle[s].decode(packets[f, s, 1:])[3], want[f][s])
    dec = opn.BatchDecoder(ns, opn.DecoderConfiguration(48000, channels, 0), **SYNTH)
    bufs = [np.zeros((ns, nf * channels), np.float32) for _ in range(2)]
    res = [np.zeros(ns, np.int32) for _ in range(2)]
    arenas = [np.ascontiguousarray(packets[f].reshape(-1)) for f in range(nfr)]
    ticket = None
    for f in range(nfr):
        q = f & 1
        t = dec.decode_float_ptrs(arenas[f].ctypes.data, offsets.ctypes.data, lens.ctypes.data, bufs[q].ctypes.data, nf * channels, nf,
                                  res[q].ctypes.data, opn.FLAG_SUBMIT_ONLY)
        assert t in (0, 1)
        if ticket is not None:
            dec.wait(ticket)
            assert np.array_equal(bufs[(f - 1) & 1], want[f - 1]), f - 1
        ticket = t
    dec.wait(ticket)
    assert np.array_equal(bufs[(nfr - 1) & 1], want[nfr - 1])
    assert np.all(res[0] == nf) and np.all(res[1] == nf)


def test_batch_lost_invalid_and_foreign_packets_do_not_poison_neighbours():
    lm, channels, pkt_bytes, ns, nfr, nf = 3, 2, 160, 16, 6, 960
    packets = opn.synth_fill(99, ns, 0, nfr, lm, channels, pkt_bytes)
    dec = opn.BatchDecoder(ns, **SYNTH)
    oracle = [O.SynthStream(lm, channels) for _ in range(ns)]
    offsets = np.arange(ns, dtype=np.uint32) * pkt_bytes
    for f in range(nfr):
        arena = packets[f].copy()
        lens = np.full(ns, pkt_bytes, np.uint32)
        expect = np.full(ns, nf, np.int32)
        want = np.zeros((ns, nf * channels), np.float32)
        kinds = {}
        if f in (2, 3):
            lens[3] = 0; kinds[3] = "lost"            # packet loss -> concealment frame
            arena[5, 0] = 0x48; kinds[5] = "silk"      # SILK-only TOC: unimplemented in the reference too
            arena[7, 0] = 0xFF; arena[7, 1] = 0x00; kinds[7] = "invalid"   # code 3 with 0 frames
            arena[9, 0] = 0xF8; kinds[9] = "mono"      # mono packet for a stereo decoder
        if f == 0:
            lens[11] = 0; kinds[11] = "lost-first"   # nothing decoded yet -> zeros, state untouched
        for s in range(ns):
            k = kinds.get(s)
            if k is None:
                want[s] = oracle[s].decode(arena[s, 1:])[3]
            elif k == "lost":
                want[s] = oracle[s].decode(b"")[3]
            elif k == "mono":  # decoded with the mono layout and copied to both channels (stream_channels, decoder.rs:332)
                ms = O.MappedStream(channels, 1)
                ms.state = oracle[s].state
                want[s] = ms.decode(arena[s, 1:], lm, 1)[1]
            elif k == "silk":
                expect[s] = -6
            elif k == "invalid":
                expect[s] = -4
        pcm = np.full((ns, nf * channels), 7.0, np.float32)
        res = dec.decode_float(arena.reshape(-1), offsets, lens, pcm, nf)
        assert np.array_equal(res, expect), f
        for s in range(ns):
            if expect[s] < 0:
                assert np.all(pcm[s] == 0)
            else:
                assert np.array_equal(pcm[s], want[s]), (f, s, kinds.get(s))


def test_batch_multi_frame_packets_and_larger_frame_size():
    """Code-1 (two CBR frames) and code-3 packets built from SYNTH-CELT/1 frames: the host splits them
    with parse_packet (lib.rs:345-498) and decodes the frames in order (decoder.rs:399-411)."""
    lm, channels, fb, ns, nf = 2, 2, 129, 8, 480
    frames = opn.synth_fill(3, ns, 0, 6, lm, channels, fb + 1)[:, :, 1:]  # payloads without TOC
    toc = 0x80 | 0x60 | (lm << 3) | 0x4
    dec = opn.BatchDecoder(ns, **SYNTH)
    oracle = [O.SynthStream(lm, channels) for _ in range(ns)]
    # packet A: code 1, frames 0,1 ; packet B: code 3 CBR with 3 frames + 5 bytes padding ; packet C: code 0
    pa = [np.concatenate([[toc | 1], frames[0, s], frames[1, s]]).astype(np.uint8) for s in range(ns)]
    pb = [np.concatenate([[toc | 3, 0x40 | 3, 5], frames[2, s], frames[3, s], frames[4, s], np.zeros(5)]).astype(np.uint8) for s in range(ns)]
    pc = [np.concatenate([[toc], frames[5, s]]).astype(np.uint8) for s in range(ns)]
    fidx = 0
    for pk, count in ((pa, 2), (pb, 3), (pc, 1)):
        ln = len(pk[0])
        arena = np.concatenate(pk)
        offsets = np.arange(ns, dtype=np.uint32) * ln
        lens = np.full(ns, ln, np.uint32)
        pcm = np.zeros((ns, 1920 * channels), np.float32)
        res = dec.decode_float(arena, offsets, lens, pcm, 1920)
        assert np.all(res == count * nf)
        for s in range(ns):
            want = np.concatenate([oracle[s].decode(frames[fidx + w, s])[3] for w in range(count)])
            assert np.array_equal(pcm[s, :count * nf * channels], want)
            assert np.all(pcm[s, count * nf * channels:] == 0)
        fidx += count
    # frame_size smaller than the packet -> FrameSizeTooSmall for every stream (decoder.rs:388-390)
    res = dec.decode_float(np.concatenate(pa), np.arange(ns, dtype=np.uint32) * len(pa[0]), np.full(ns, len(pa[0]), np.uint32),
                           np.zeros((ns, 480 * channels), np.float32), 480)
    assert np.all(res == -5)
    with pytest.raises(opn.OpusError):  # frame_size not a multiple of 2.5 ms (decoder.rs:316-320)
        dec.decode_float(np.concatenate(pc), np.arange(ns, dtype=np.uint32), np.ones(ns, np.uint32), None, 100)


def test_batch_mixed_frame_sizes_across_streams_in_one_call():
    """BASELINE configs[4] shape, CELT part: every stream has its own fixed frame size (2.5/5/10/20 ms, weights
    10/10/30/50 %), 10 % of the frames are transient and 3 % of the packets are lost.  One host call decodes
    one packet of every stream: the runtime buckets the items by frame size and chains the buckets through the
    stage pipeline.  Checked bit for bit against one oracle decoder per stream for 12 calls."""
    rnd = np.random.default_rng(5)
    ns, channels, ncalls = 1300, 2, 12  # two chunks on the host path
    lm_of = rnd.choice([0, 1, 2, 3], size=ns, p=[0.1, 0.1, 0.3, 0.5])
    pkt_bytes = {0: 80, 1: 100, 2: 130, 3: 160}
    oracle = [O.SynthStream(int(lm_of[s]), channels) for s in range(ns)]
    by_lm = {lm: np.nonzero(lm_of == lm)[0] for lm in range(4)}
    dec = opn.BatchDecoder(ns, **SYNTH)
    stride = 160
    exact = True
    for f in range(ncalls):
        arena = np.zeros(ns * stride, np.uint8)
        lens = np.zeros(ns, np.uint32)
        for lm, ids in by_lm.items():
            if len(ids) == 0:
                continue
            pk = opn.synth_fill(9000 + lm, len(ids), f, 1, lm, channels, pkt_bytes[lm], transient_permille=100)[0]
            for k, s in enumerate(ids):
                arena[s * stride:s * stride + pkt_bytes[lm]] = pk[k]
                lens[s] = pkt_bytes[lm]
        lost = rnd.random(ns) < 0.03
        if f == 0:
            lost[:] = False
        lens[lost] = 0
        offsets = np.arange(ns, dtype=np.uint32) * stride
        pcm = np.zeros((ns, 960 * channels), np.float32)
        res = dec.decode_float(arena, offsets, lens, pcm, 960)
        for s in range(ns):
            nf = 120 << int(lm_of[s])
            if lost[s]:
                # decode_native(None) conceals frame_size samples in frames of the last size (decoder.rs:427-441)
                assert res[s] == 960
                want = np.concatenate([oracle[s].decode(b"")[3] for _ in range(960 // nf)])
                got = pcm[s]
            else:
                assert res[s] == nf, (f, s, res[s])
                want = oracle[s].decode(arena[s * stride + 1:s * stride + int(lens[s])])[3]
                got = pcm[s, :nf * channels]
                assert np.all(pcm[s, nf * channels:] == 0)
            assert_pcm(want, got, f"call {f} stream {s}")
            exact &= np.array_equal(want, got)
    assert exact, "PCM within tolerance but not bit-identical to the oracle"


def test_batch_device_resident_path_matches_host_path():
    torch = pytest.importorskip("torch")
    lm, channels, pkt_bytes, ns, nfr, nf = 3, 2, 160, 64, 5, 960
    packets = opn.synth_fill(500, ns, 0, nfr, lm, channels, pkt_bytes, transient_permille=100)
    packets[2, 10, 0] = 0x48  # one foreign packet: reported per stream, neighbours unaffected
    want, _ = _oracle_chain(packets[:, [s for s in range(ns) if s != 10]], lm, channels)
    dev = torch.device("cuda:0")
    d_arena = torch.from_numpy(packets.reshape(-1).copy()).to(dev)
    d_off = torch.arange(ns, dtype=torch.int32, device=dev) * pkt_bytes
    d_len = torch.full((ns,), pkt_bytes, dtype=torch.int32, device=dev)
    d_pcm = torch.zeros((ns, nf * channels), dtype=torch.float32, device=dev)
    d_res = torch.zeros(ns, dtype=torch.int32, device=dev)
    dec = opn.BatchDecoder(ns, **SYNTH)
    keep = [s for s in range(ns) if s != 10]
    for f in range(nfr):
        dec.decode_float_ptrs(d_arena.data_ptr() + f * ns * pkt_bytes, d_off.data_ptr(), d_len.data_ptr(), d_pcm.data_ptr(),
                              nf * channels, nf, d_res.data_ptr(), opn.FLAG_DEVICE_PTRS)
        dec.synchronize()
        res = d_res.cpu().numpy()
        if f == 2:
            assert res[10] == -6
        assert np.all(res[keep] == nf)
        assert np.array_equal(d_pcm.cpu().numpy()[keep], want[f])
    # ring view: the last frame of stream 0 sits just before ring_pos
    ring_ptr, ring_n, pos_ptr = dec.ring()
    assert ring_n == 2880 and ring_ptr and pos_ptr  # 3 x 960: a frame plus the 1024 + 2 samples of history before it


@pytest.mark.parametrize("channels,bitstream", [(2, 1), (1, 1), (2, 2), (1, 2)])
def test_batch_packets_of_either_channel_count(channels, bitstream):
    """stream_channels (decoder.rs:332,376,395): a decoder of `channels` channels takes mono and stereo packets, switching
    per packet.  The packet is decoded with its own layout and mapped in the frame kernel -- mono -> stereo transforms the
    same spectrum into both channels (each with its own overlap and post-filter history), stereo -> mono transforms
    0.5 * (l + r) -- against the oracle's mapped decode, bit for bit; frame sizes mixed as well, some packets lost."""
    rnd = np.random.default_rng(31 * channels + bitstream)
    ns, ncalls, stride = 260, 9, 160
    pkt_bytes = {0: 64, 1: 80, 2: 110, 3: 160}
    fill = opn.synth_fill if bitstream == 1 else opn.celt2_fill
    oracle = [O.MappedStream(channels, bitstream) for _ in range(ns)]
    dec = opn.BatchDecoder(ns, opn.DecoderConfiguration(48000, channels, 0), **(SYNTH if bitstream == 1 else CELT2))
    last_lm = np.full(ns, 3)
    exact = True
    for f in range(ncalls):
        lm_of = rnd.choice([0, 1, 2, 3], size=ns, p=[0.1, 0.1, 0.3, 0.5])
        cs_of = rnd.choice([1, 2], size=ns)
        arena = np.zeros((ns, stride), np.uint8)
        lens = np.zeros(ns, np.uint32)
        for lm in range(4):
            for cs in (1, 2):
                ids = np.nonzero((lm_of == lm) & (cs_of == cs))[0]
                if len(ids):
                    arena[ids, :pkt_bytes[lm]] = fill(700 + 10 * lm + cs, len(ids), f, 1, lm, cs, pkt_bytes[lm], transient_permille=100)[0]
                    lens[ids] = pkt_bytes[lm]
        if f > 0:
            lens[rnd.random(ns) < 0.04] = 0
        pcm = np.zeros((ns, 960 * channels), np.float32)
        res = dec.decode_float(arena.reshape(-1), np.arange(ns, dtype=np.uint32) * stride, lens, pcm, 960)
        rng = dec.final_ranges()
        for s in range(ns):
            if lens[s] == 0:
                nf = 120 << int(last_lm[s])
                assert res[s] == 960
                want = np.concatenate([oracle[s].decode(b"", int(last_lm[s]), channels)[1] for _ in range(960 // nf)])
                got = pcm[s]
            else:
                nf = 120 << int(lm_of[s])
                assert res[s] == nf, (f, s, res[s])
                fr, want = oracle[s].decode(arena[s, 1:int(lens[s])], int(lm_of[s]), int(cs_of[s]))
                assert rng[s] == fr
                got = pcm[s, :nf * channels]
                last_lm[s] = lm_of[s]
            assert_pcm(want, got, (f, s))
            exact &= np.array_equal(want, got)
    assert exact, "PCM within tolerance but not bit-identical to the oracle"


def test_decoder_api_mono_packet_into_stereo_decoder():
    """Decoder::decode_float (decoder.rs:216-232) on a stereo decoder fed a mono packet: both output channels carry the
    packet's one channel (their histories were equal), and the call returns the packet's frame size."""
    dec = opn.Decoder(opn.DecoderConfiguration(48000, 2, 0), bitstream=opn.BITSTREAM_SYNTH_CELT_1)
    ora = O.MappedStream(2, 1)
    for f in range(4):
        pk = opn.synth_fill(5, 1, f, 1, 3, 1, 100)[0, 0]
        out = np.zeros(960 * 2, np.float32)
        n = dec.decode_float(pk, out, 960)
        assert n == 960
        _, want = ora.decode(pk[1:], 3, 1)
        assert np.array_equal(out, want)
        assert np.array_equal(out[0::2], out[1::2]) and np.abs(out).max() > 0


@pytest.mark.parametrize("ns,channels,bitstream", [(2300, 2, 1), (300, 1, 1), (500, 2, 2)])
def test_batch_device_resident_mixed_frame_sizes(ns, channels, bitstream):
    """OPN_FLAG_MIXED_FRAMES: a device-resident step whose streams hold frames of different sizes.  Every stream picks a
    new frame size for every packet (2.5/5/10/20 ms, weights 10/10/30/50 %; 10 % transient frames), 3 % of the packets
    are lost (concealed with one frame of the stream's previous size), a few carry a foreign TOC.  The buckets are built
    on the device (k_mix_key / k_mix_place); ten steps are enqueued back to back (2300 streams: three frame-kernel groups
    per step, 300: one) and every step's PCM, sample counts and the final ranges are compared with one oracle decoder per
    stream, bit for bit."""
    torch = pytest.importorskip("torch")
    rnd = np.random.default_rng(11 + ns)
    nsteps, stride, cap = 10, 160, 960
    pkt_bytes = {0: 64, 1: 80, 2: 110, 3: 160}
    fill = opn.synth_fill if bitstream == 1 else opn.celt2_fill
    oracle = [(O.SynthStream if bitstream == 1 else O.Celt2Stream)(3, channels) for _ in range(ns)]

    def oracle_decode(s, payload):  # -> (final_rng, pcm)
        if bitstream == 1:
            side, _, _, pcm = oracle[s].decode(payload)
        else:
            side, pcm = oracle[s].decode(payload)
            assert pcm is not None
        return side.final_rng, pcm

    arena = np.zeros((nsteps, ns, stride), np.uint8)
    lens = np.zeros((nsteps, ns), np.uint32)
    lm_of = np.zeros((nsteps, ns), np.int64)
    foreign = np.zeros((nsteps, ns), bool)
    for f in range(nsteps):
        lm_of[f] = rnd.choice([0, 1, 2, 3], size=ns, p=[0.1, 0.1, 0.3, 0.5])
        for lm in range(4):
            ids = np.nonzero(lm_of[f] == lm)[0]
            if len(ids) == 0:
                continue
            pk = fill(4000 + lm, len(ids), f, 1, lm, channels, pkt_bytes[lm], transient_permille=100)[0]
            arena[f, ids, :pkt_bytes[lm]] = pk
            lens[f, ids] = pkt_bytes[lm]
        if f > 0:
            lens[f, rnd.random(ns) < 0.03] = 0
        for s in rnd.choice(ns, 3, replace=False):  # SILK TOC, multi-frame code, wrong channel count
            kind = int(rnd.integers(3))
            if lens[f, s]:
                foreign[f, s] = True
                arena[f, s, 0] = [0x48, arena[f, s, 0] | 1, arena[f, s, 0] ^ 4][kind]
    dev = torch.device("cuda:0")
    d_arena = torch.from_numpy(arena.reshape(-1).copy()).to(dev)
    d_off = torch.arange(ns, dtype=torch.int32, device=dev) * stride
    d_len = torch.from_numpy(lens.astype(np.int32)).to(dev)
    d_pcm = torch.zeros((nsteps, ns, cap * channels), dtype=torch.float32, device=dev)
    d_res = torch.full((nsteps, ns), -99, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    dec = opn.BatchDecoder(ns, opn.DecoderConfiguration(48000, channels, 0), **(SYNTH if bitstream == 1 else CELT2))
    flags = opn.FLAG_DEVICE_PTRS | opn.FLAG_INPUTS_READY | opn.FLAG_MIXED_FRAMES
    for f in range(nsteps):
        dec.decode_float_ptrs(d_arena.data_ptr() + f * ns * stride, d_off.data_ptr(), d_len[f].data_ptr(), d_pcm[f].data_ptr(),
                              cap * channels, cap, d_res[f].data_ptr(), flags)
    dec.join()
    dec.synchronize()
    res, got = d_res.cpu().numpy(), d_pcm.cpu().numpy()
    last_lm = np.full(ns, -1)
    last_rng = np.zeros(ns, np.uint32)
    for f in range(nsteps):
        for s in range(ns):
            if foreign[f, s]:
                assert res[f, s] == -6, (f, s, res[f, s])  # Unimplemented, state untouched
                assert not got[f, s].any()
                continue
            if lens[f, s] == 0:
                lm = last_lm[s] if last_lm[s] >= 0 else 3  # nothing decoded yet: zeros of the row's capacity (decoder.rs:473-484)
                oracle[s].lm = int(lm)
                _, want = oracle_decode(s, b"")
                last_rng[s] = 0
            else:
                lm = lm_of[f, s]
                oracle[s].lm = int(lm)
                last_rng[s], want = oracle_decode(s, arena[f, s, 1:int(lens[f, s])])
                last_lm[s] = lm
            nf = 120 << int(lm)
            assert res[f, s] == nf, (f, s, res[f, s], nf)
            assert np.array_equal(got[f, s, :nf * channels], want), (f, s, lm)
            assert not got[f, s, nf * channels:].any()
    assert np.array_equal(dec.final_ranges(), last_rng)
    # a row too short for the stream's frame: FrameSizeTooSmall for that stream only (decoder.rs:388-390)
    d_res2 = torch.full((ns,), -99, dtype=torch.int32, device=dev)
    dec.decode_float_ptrs(d_arena.data_ptr(), d_off.data_ptr(), d_len[0].data_ptr(), d_pcm[0].data_ptr(), cap * channels, 240,
                          d_res2.data_ptr(), flags)
    dec.join()
    dec.synchronize()
    r2 = d_res2.cpu().numpy()
    ok = ~foreign[0]
    assert np.all(r2[ok & (lm_of[0] > 1)] == -5) and np.all(r2[ok & (lm_of[0] <= 1)] == (120 << lm_of[0][ok & (lm_of[0] <= 1)]))
    # after a reset nothing has been decoded: a step of losses only gives zeros of the row's capacity (decoder.rs:473-484),
    # and every bucket but one is empty
    dec.reset()
    d_zero = torch.zeros(ns, dtype=torch.int32, device=dev)
    for cap2 in (960, 240):
        d_pcm[0].fill_(7.0)
        dec.decode_float_ptrs(d_arena.data_ptr(), d_off.data_ptr(), d_zero.data_ptr(), d_pcm[0].data_ptr(), cap * channels, cap2,
                              d_res2.data_ptr(), flags)
        dec.join()
        dec.synchronize()
        assert np.all(d_res2.cpu().numpy() == cap2)
        out = d_pcm[0].cpu().numpy()
        assert not out[:, :cap2 * channels].any() and np.all(out[:, cap2 * channels:] == 7.0)


@pytest.mark.parametrize("postfilter", [True, False])
def test_batch_device_resident_steps_enqueued_back_to_back(postfilter):
    """OPN_FLAG_INPUTS_READY: 14 steps are enqueued without a host wait in between, so range decode, PVQ
    expansion, IMDCT and post-filter of up to eight different steps are in flight at once (eight buffer sets,
    kernel 2 one step behind kernel 1 on the four-frame ring).  Every step's PCM row is written to its own
    dense buffer and checked after one final synchronize, bit for bit, against the oracle."""
    torch = pytest.importorskip("torch")
    lm, channels, pkt_bytes, ns, nfr, nf = 3, 2, 160, 96, 14, 960
    packets = opn.synth_fill(77, ns, 0, nfr, lm, channels, pkt_bytes, transient_permille=120)
    want, want_rng = _oracle_chain(packets, lm, channels, postfilter)
    dev = torch.device("cuda:0")
    d_arena = torch.from_numpy(packets.reshape(-1).copy()).to(dev)
    d_off = torch.arange(ns, dtype=torch.int32, device=dev) * pkt_bytes
    d_len = torch.full((ns,), pkt_bytes, dtype=torch.int32, device=dev)
    d_pcm = torch.zeros((nfr, ns, nf * channels), dtype=torch.float32, device=dev)
    d_res = torch.zeros((nfr, ns), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    dec = opn.BatchDecoder(ns, opn.DecoderConfiguration(48000, channels, 0), postfilter=postfilter, **SYNTH)
    flags = opn.FLAG_DEVICE_PTRS | opn.FLAG_INPUTS_READY
    for f in range(nfr):
        dec.decode_float_ptrs(d_arena.data_ptr() + f * ns * pkt_bytes, d_off.data_ptr(), d_len.data_ptr(), d_pcm[f].data_ptr(),
                              nf * channels, nf, d_res[f].data_ptr(), flags)
    dec.join()
    dec.synchronize()
    assert np.all(d_res.cpu().numpy() == nf)
    got = d_pcm.cpu().numpy()
    for f in range(nfr):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(dec.final_ranges(), want_rng[nfr - 1])


def test_batch_reset_in_the_middle_of_a_pipelined_run():
    """Decoder::reset (decoder.rs:74, 286-303) on a batch with several steps in flight: after the reset the
    streams decode exactly like a new batch (overlap carry, PCM ring, post-filter parameters, soft-clip memory,
    buffer-set rotation all start over), and API misuse is rejected with BadArguments rather than run."""
    torch = pytest.importorskip("torch")
    lm, channels, pkt_bytes, ns, nf = 3, 2, 160, 80, 960
    packets = opn.synth_fill(3100, ns, 0, 9, lm, channels, pkt_bytes, transient_permille=100)
    dev = torch.device("cuda:0")
    d_arena = torch.from_numpy(packets.reshape(-1).copy()).to(dev)
    d_off = torch.arange(ns, dtype=torch.int32, device=dev) * pkt_bytes
    d_len = torch.full((ns,), pkt_bytes, dtype=torch.int32, device=dev)
    d_res = torch.zeros(ns, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    flags = opn.FLAG_DEVICE_PTRS | opn.FLAG_INPUTS_READY

    def run(dec, frames):
        out = torch.zeros((len(frames), ns, nf * channels), dtype=torch.float32, device=dev)
        for k, f in enumerate(frames):
            dec.decode_float_ptrs(d_arena.data_ptr() + f * ns * pkt_bytes, d_off.data_ptr(), d_len.data_ptr(), out[k].data_ptr(),
                                  nf * channels, nf, d_res.data_ptr(), flags)
        dec.synchronize()
        return out.cpu().numpy()

    dec = opn.BatchDecoder(ns, **SYNTH)
    run(dec, [0, 1, 2, 3, 4])          # leaves state behind; five steps were in flight
    dec.reset()
    again = run(dec, [5, 6, 7, 8])
    fresh = run(opn.BatchDecoder(ns, **SYNTH), [5, 6, 7, 8])
    assert np.array_equal(again, fresh)
    want, _ = _oracle_chain(packets[5:9], lm, channels)
    assert np.array_equal(again, want)
    # misuse
    with pytest.raises(opn.OpusError) as e:
        dec.decode_float_ptrs(d_arena.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), None, 0, 100, d_res.data_ptr(), flags)
    assert e.value.kind == "BadArguments"                      # frame size not a multiple of 2.5 ms
    with pytest.raises(opn.OpusError) as e:
        dec.wait(5)
    assert e.value.kind == "BadArguments"
    out16 = np.zeros((ns, nf * channels), np.int16)
    with pytest.raises(opn.OpusError) as e:
        dec.decode_i16_ptrs(0, 0, 0, out16.ctypes.data, nf * channels, nf, 0, 0)
    assert e.value.kind == "BadArguments"


def test_batch_reset_without_synchronize_while_steps_are_in_flight():
    """opn_batch_reset straight after enqueueing device-resident steps, with NO synchronize in between: the frame
    kernels of those steps are still running (they write the PCM ring, the carry and the post-filter state), so the
    reset has to drain every pipeline stream before it clears anything.  Afterwards the batch must decode exactly like
    a new one; a frame kernel that outlived the memsets would leave a stale frame in the ring = comb history of the
    frames that follow."""
    torch = pytest.importorskip("torch")
    lm, channels, pkt_bytes, ns, nf = 3, 2, 160, 2048, 960  # large enough for the enqueued steps to outlast the host
    packets = opn.synth_fill(5100, ns, 0, 8, lm, channels, pkt_bytes, transient_permille=100)
    dev = torch.device("cuda:0")
    d_arena = torch.from_numpy(packets.reshape(-1).copy()).to(dev)
    d_off = torch.arange(ns, dtype=torch.int32, device=dev) * pkt_bytes
    d_len = torch.full((ns,), pkt_bytes, dtype=torch.int32, device=dev)
    d_res = torch.zeros(ns, dtype=torch.int32, device=dev)
    out = torch.zeros((4, ns, nf * channels), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    flags = opn.FLAG_DEVICE_PTRS | opn.FLAG_INPUTS_READY

    def step(dec, f, dst):
        dec.decode_float_ptrs(d_arena.data_ptr() + f * ns * pkt_bytes, d_off.data_ptr(), d_len.data_ptr(), dst, nf * channels if dst else 0,
                              nf, d_res.data_ptr(), flags | (0 if dst else opn.FLAG_NO_PCM_COPY))

    dec = opn.BatchDecoder(ns, **SYNTH)
    for rep in range(3):
        for f in range(4):
            step(dec, f, None)
        dec.reset()  # no synchronize: four steps are in flight
        for k, f in enumerate(range(4, 8)):
            step(dec, f, out[k].data_ptr())
        dec.synchronize()
        got = out.cpu().numpy()
        if rep == 0:
            picks = [0, 1, 777, ns - 1]
            want, _ = _oracle_chain(packets[4:8][:, picks], lm, channels)
        assert np.array_equal(got[:, picks], want), rep
        if rep == 0:
            first = got.copy()
        assert np.array_equal(got, first), rep


def test_pageable_pinned_and_registered_caller_buffers_give_the_same_pcm():
    """The host-buffer entry point takes any host memory: pageable numpy arrays (what a Rust slice is: staged,
    synchronous copies), buffers from opn_host_alloc and memory pinned in place with opn_host_register (asynchronous DMA,
    two calls in flight).  Same packets, same PCM, bit for bit; the oracle checks one of them."""
    lm, channels, pkt_bytes, ns, nfr, nf = 3, 2, 160, 1500, 4, 960
    packets = opn.synth_fill(8800, ns, 0, nfr, lm, channels, pkt_bytes, transient_permille=100)
    offsets = np.arange(ns, dtype=np.uint32) * pkt_bytes
    lens = np.full(ns, pkt_bytes, np.uint32)

    def run(make_pcm, arena_of, submit_only):
        dec = opn.BatchDecoder(ns, **SYNTH)
        out, keep, ticket = [], [], None
        for f in range(nfr):
            pcm = make_pcm()
            keep.append(pcm)
            arr = pcm.array if isinstance(pcm, opn.HostBuffer) else pcm
            res = np.zeros(ns, np.int32)
            t = dec.decode_float_ptrs(arena_of(f).ctypes.data, offsets.ctypes.data, lens.ctypes.data, arr.ctypes.data, nf * channels, nf,
                                      res.ctypes.data, opn.FLAG_SUBMIT_ONLY if submit_only else 0)
            if submit_only:
                if ticket is not None:
                    dec.wait(ticket)
                ticket = t
            assert np.all(res == nf)
            out.append(arr)
        if submit_only:
            dec.wait(ticket)
        return np.stack([o.copy() for o in out])

    pageable = run(lambda: np.zeros((ns, nf * channels), np.float32), lambda f: packets[f].reshape(-1), False)
    pinned_arena = opn.HostBuffer(packets.shape, np.uint8)
    pinned_arena.array[...] = packets
    pinned = run(lambda: opn.HostBuffer((ns, nf * channels), np.float32), lambda f: pinned_arena.array[f].reshape(-1), True)
    regs = [np.zeros((ns, nf * channels), np.float32) for _ in range(nfr)]
    for r in regs:
        opn.host_register(r)
    it = iter(regs)
    registered = run(lambda: next(it), lambda f: packets[f].reshape(-1), True)
    for r in regs:
        opn.host_unregister(r)
    assert np.array_equal(pageable, pinned) and np.array_equal(pageable, registered)
    picks = [0, 1023, 1024, ns - 1]
    want, _ = _oracle_chain(packets[:, picks], lm, channels)
    assert np.array_equal(pageable[:, picks], want)


def test_celt_frames_need_the_explicit_bitstream_opt_in():
    """CeltDecoder::decode is todo!() in the crate (celt/decoder.rs:47-56).  A decoder created the way a drop-in caller
    creates it (OPN_BITSTREAM_OPUS) must not turn CELT packets into sound: it reports Unimplemented, per stream on the
    batch path, and leaves the streams' state alone; only OPN_BITSTREAM_SYNTH_CELT_1 decodes the synthetic layout."""
    lm, channels, pkt_bytes, ns, nf = 3, 2, 160, 8, 960
    packets = opn.synth_fill(1, ns, 0, 1, lm, channels, pkt_bytes)[0]
    offsets = np.arange(ns, dtype=np.uint32) * pkt_bytes
    lens = np.full(ns, pkt_bytes, np.uint32)
    lens[3] = 0  # a lost packet before anything was decoded: zeros (decoder.rs:478-487), not an error
    pcm = np.full((ns, nf * channels), 7.0, np.float32)
    res = opn.BatchDecoder(ns).decode_float(packets.reshape(-1), offsets, lens, pcm, nf)
    assert [int(r) for r in res] == [-6, -6, -6, nf, -6, -6, -6, -6]
    assert not pcm[3].any()
    with pytest.raises(opn.OpusError) as e:
        opn.Decoder().decode_float(packets[0], np.zeros(nf * channels, np.float32), nf)
    assert e.value.kind == "Unimplemented"
    torch = pytest.importorskip("torch")
    d = torch.zeros(16, dtype=torch.int32, device="cuda:0")
    with pytest.raises(opn.OpusError) as e:
        opn.BatchDecoder(4).decode_float_ptrs(d.data_ptr(), d.data_ptr(), d.data_ptr(), None, 0, nf, d.data_ptr(),
                                              opn.FLAG_DEVICE_PTRS | opn.FLAG_NO_PCM_COPY)
    assert e.value.kind == "Unimplemented"
    res = opn.BatchDecoder(ns, **SYNTH).decode_float(packets.reshape(-1), offsets, lens, pcm, nf)
    assert [int(r) for r in res] == [nf] * ns


# ------------------------------------------------------------------ Decoder API (decoder.rs:27-232)
def test_baseline_config0_one_mono_stream_1000_chained_frames():
    """BASELINE.json configs[0]: one synthetic 20 ms 48 kHz mono stream, 1000 chained frames, decoded through
    Decoder::decode_float one packet at a time; the CPU oracle's output is the reference output.  Every
    frame must be bit-identical (PCM and final_range): state (overlap carry, post-filter history across
    many ring wraps, post-filter parameters) is carried for 20 s of audio."""
    lm, channels, pkt_bytes, nf, nfr = 3, 1, 100, 960, 1000
    packets = opn.synth_fill(4242, 1, 0, nfr, lm, channels, pkt_bytes, transient_permille=100)[:, 0]
    dec = opn.Decoder(opn.DecoderConfiguration(48000, 1, 0), **SYNTH)
    st = O.SynthStream(lm, channels)
    pcm = np.zeros(nf, np.float32)
    energy = 0.0
    for f in range(nfr):
        assert dec.decode_float(packets[f], pcm, nf) == nf
        side, _, _, want = st.decode(packets[f, 1:])
        assert np.array_equal(pcm, want), f
        assert dec.final_range == side.final_rng, f
        energy += float((pcm.astype(np.float64) ** 2).sum())
    assert 0.01 < np.sqrt(energy / (nf * nfr)) < 1.0  # a real signal, not silence


def test_decoder_api_single_stream():
    lm, channels, pkt_bytes, nf = 3, 2, 160, 960
    dec = opn.Decoder(opn.DecoderConfiguration(48000, 2, 0), **SYNTH)
    assert (dec.sampling_rate, dec.channels, dec.gain) == (48000, 2, 0)
    assert dec.bandwidth is None and dec.last_packet_duration is None and dec.pitch is None
    st = O.SynthStream(lm, channels)
    pcm = np.zeros(nf * channels, np.float32)
    # loss before any packet: zeros (decoder.rs:478-487)
    assert dec.decode_float(None, pcm, nf) == nf and np.all(pcm == 0)
    for f in range(4):
        pkt, truth = opn.synth_packet(1, f, lm, channels, pkt_bytes)
        assert dec.decode_float(pkt, pcm, nf) == nf
        side, _, _, want = st.decode(pkt[1:])
        assert np.array_equal(pcm, want)
        assert dec.final_range == side.final_rng
        assert dec.bandwidth == "Fullband" and dec.last_packet_duration == nf
        assert dec.pitch == (truth["period"] if truth["postfilter"] else 0)
    # lost packet -> concealment, final_range 0 (decoder.rs:799-803)
    assert dec.decode_float(None, pcm, nf) == nf
    assert np.array_equal(pcm, st.decode(b"")[3]) and dec.final_range == 0
    # decode_fec on a CELT stream conceals instead (decoder.rs:343-350)
    pkt, _ = opn.synth_packet(1, 9, lm, channels, pkt_bytes)
    assert dec.decode_float(pkt, pcm, nf, decode_fec=True) == nf
    assert np.array_equal(pcm, st.decode(b"")[3])
    # argument errors (decoder.rs:316-325, 388-390)
    for bad in (100, 0):
        with pytest.raises(opn.OpusError) as e:
            dec.decode_float(pkt, np.zeros(4000, np.float32), bad)
        assert e.value.kind == "BadArguments"
    with pytest.raises(opn.OpusError) as e:
        dec.decode_float(b"", pcm, nf)
    assert e.value.kind == "BadArguments"
    with pytest.raises(opn.OpusError) as e:
        dec.decode_float(pkt, np.zeros(480 * 2, np.float32), 480)
    assert e.value.kind == "FrameSizeTooSmall"
    with pytest.raises(opn.OpusError) as e:
        dec.decode_float(bytes([0x48, 1, 2, 3]), pcm, nf)
    assert e.value.kind == "Unimplemented"
    dec.reset()
    assert dec.bandwidth is None
    pkt0, _ = opn.synth_packet(1, 0, lm, channels, pkt_bytes)
    assert dec.decode_float(pkt0, pcm, nf) == nf
    assert np.array_equal(pcm, O.SynthStream(lm, channels).decode(pkt0[1:])[3])


def test_decoder_gain_and_i16_output():
    lm, channels, pkt_bytes, nf = 3, 2, 160, 960
    gain_q8 = 1536  # +6 dB
    dec = opn.Decoder(opn.DecoderConfiguration(48000, 2, gain_q8), **SYNTH)
    st = O.SynthStream(lm, channels)
    g = np.float32(np.exp(np.float32(np.float32(6.48814081e-4) * np.float32(gain_q8)) * np.float32(0.6931471805599453)))
    pcm = np.zeros(nf * channels, np.float32)
    mem = np.zeros(2, np.float32)
    dec16 = opn.Decoder(opn.DecoderConfiguration(48000, 2, gain_q8), **SYNTH)
    st16 = O.SynthStream(lm, channels)
    for f in range(3):
        pkt, _ = opn.synth_packet(2, f, lm, channels, pkt_bytes)
        assert dec.decode_float(pkt, pcm, nf) == nf
        want = st.decode(pkt[1:])[3] * g
        assert_pcm(want, pcm)
        out16 = np.zeros(nf * channels, np.int16)
        assert dec16.decode(pkt, out16, nf) == nf
        w = st16.decode(pkt[1:])[3] * g
        # decode<S>: soft clip over samples[..sample_count] (reference quirk, decoder.rs:415-419) then from_f32
        O.lib().orc_pcm_soft_clip(O.ptr(w), nf, channels, O.ptr(mem), 2)
        w16 = np.clip(w * np.float32(32768.0), -32768.0, 32767.0).astype(np.int16)
        assert np.abs(out16.astype(np.int32) - w16.astype(np.int32)).max() <= 1


def test_batch_decode_i16_matches_single_stream_decode_and_oracle():
    """opn_batch_decode_i16 = Decoder::decode::<i16> for every stream: device-side pcm_soft_clip (per-stream
    memory carried from frame to frame) + Sample::from_f32.  Checked against the single-stream decode::<i16>
    (host conversion) sample for sample, and against the oracle's float PCM -> orc_pcm_soft_clip -> from_f32
    within one LSB; a +12 dB decoder gain makes the clip non-trivial (peaks beyond full scale)."""
    lm, channels, pkt_bytes, ns, nfr, nf = 3, 2, 160, 1100, 5, 960  # two chunks on the host path
    gain_q8 = 3072  # +12 dB
    g = np.float32(np.exp(np.float32(np.float32(6.48814081e-4) * np.float32(gain_q8)) * np.float32(0.6931471805599453)))
    packets = opn.synth_fill(900, ns, 0, nfr, lm, channels, pkt_bytes, transient_permille=100)
    cfg = opn.DecoderConfiguration(48000, channels, gain_q8)
    batch = opn.BatchDecoder(ns, cfg, **SYNTH)
    singles = {s: opn.Decoder(cfg, **SYNTH) for s in (0, 7, 1099)}
    oracle = {s: (O.SynthStream(lm, channels), np.zeros(2, np.float32)) for s in singles}
    offsets = np.arange(ns, dtype=np.uint32) * pkt_bytes
    lens = np.full(ns, pkt_bytes, np.uint32)
    clipped = 0
    for f in range(nfr):
        out = np.zeros((ns, nf * channels), np.int16)
        res, _ = batch.decode_i16(packets[f].reshape(-1), offsets, lens, out, nf)
        assert np.all(res == nf)
        for s, dec in singles.items():
            one = np.zeros(nf * channels, np.int16)
            assert dec.decode(packets[f, s], one, nf) == nf
            assert np.array_equal(one, out[s]), (f, s)
            st, mem = oracle[s]
            w = st.decode(packets[f, s, 1:])[3] * g
            clipped += int((np.abs(w) > 1.0).sum())
            O.lib().orc_pcm_soft_clip(O.ptr(w), nf, channels, O.ptr(mem), 2)
            w16 = np.clip(w * np.float32(32768.0), -32768.0, 32767.0).astype(np.int16)
            assert np.abs(out[s].astype(np.int32) - w16.astype(np.int32)).max() <= 1, (f, s)
    assert clipped > 0, "the test signal never exceeded full scale"


def test_decode_i16_lost_packet_after_a_clipping_frame_is_not_clipped():
    """decode_native's None branch (decoder.rs:427-441) neither calls pcm_soft_clip nor touches softclip_mem: a
    concealed frame that follows a clipping frame goes through Sample::from_f32 unclipped, and the clip memory the
    next real packet sees is the one the last real packet left.  Batch and single-stream decode::<i16>, against the
    oracle chain float PCM -> (clip only on real packets) -> from_f32."""
    lm, channels, pkt_bytes, ns, nf = 3, 2, 160, 40, 960
    gain_q8 = 3072  # +12 dB: peaks beyond full scale
    g = np.float32(np.exp(np.float32(np.float32(6.48814081e-4) * np.float32(gain_q8)) * np.float32(0.6931471805599453)))
    packets = opn.synth_fill(7700, ns, 0, 6, lm, channels, pkt_bytes)
    cfg = opn.DecoderConfiguration(48000, channels, gain_q8)
    batch = opn.BatchDecoder(ns, cfg, **SYNTH)
    single = opn.Decoder(cfg, **SYNTH)
    lost_at = {1: {3, 4}, 2: {3}, 4: {3, 9}}  # frame -> streams that lose their packet
    offsets = np.arange(ns, dtype=np.uint32) * pkt_bytes
    oracle = {s: (O.SynthStream(lm, channels), np.zeros(2, np.float32)) for s in (3, 4, 9, 20)}
    clipped = 0
    for f in range(6):
        lens = np.full(ns, pkt_bytes, np.uint32)
        for s in lost_at.get(f, ()):
            lens[s] = 0
        out = np.zeros((ns, nf * channels), np.int16)
        res, _ = batch.decode_i16(packets[f].reshape(-1), offsets, lens, out, nf)
        assert np.all(res == nf)
        one = np.zeros(nf * channels, np.int16)
        assert single.decode(packets[f, 3] if lens[3] else None, one, nf) == nf
        assert np.array_equal(one, out[3]), f
        for s, (st, mem) in oracle.items():
            lost = lens[s] == 0
            w = st.decode(b"" if lost else packets[f, s, 1:])[3] * g
            if not lost:
                clipped += int((np.abs(w) > 1.0).sum())
                O.lib().orc_pcm_soft_clip(O.ptr(w), nf, channels, O.ptr(mem), 2)
            w16 = np.zeros(nf * channels, np.int16)
            assert O.lib().orc_sample_from_f32(1, O.ptr(w), O.ptr(w16), w.size) == 0
            assert np.abs(out[s].astype(np.int32) - w16.astype(np.int32)).max() <= 1, (f, s)
    assert clipped > 0


@pytest.mark.parametrize("dtype,fmt,full_scale", [(np.int16, 1, 32768.0), (np.int32, 2, 2147483648.0), (np.uint16, 3, 32768.0),
                                                   (np.uint32, 4, 2147483648.0), (np.float64, 5, 1.0)])
def test_decode_generic_sample_types_match_oracle(dtype, fmt, full_scale):
    """Decoder::decode::<S> for every S the crate implements `Sample` for (lib.rs:63-107), batch and single-stream.
    Three checks per frame: (1) the single-stream decode::<S> equals the batch row; (2) S::from_f32 on the device
    is exact: the S output equals orc_sample_from_f32 of the device's own decode::<f32> output (same packets, a
    second decoder), for every stream and sample; (3) the whole chain against the oracle's float PCM ->
    orc_pcm_soft_clip -> orc_sample_from_f32 within 4e-5 of full scale or one step of S (the +12 dB gain factor is computed by the
    host library and by numpy here, which may differ in the last bit).  The gain drives peaks past full scale so
    that the clip, the clamps and the saturating casts all fire."""
    lm, channels, pkt_bytes, ns, nfr, nf = 3, 2, 160, 96, 4, 960
    gain_q8 = 3072
    g = np.float32(np.exp(np.float32(np.float32(6.48814081e-4) * np.float32(gain_q8)) * np.float32(0.6931471805599453)))
    packets = opn.synth_fill(4000, ns, 0, nfr, lm, channels, pkt_bytes, transient_permille=100)
    cfg = opn.DecoderConfiguration(48000, channels, gain_q8)
    batch, batch_f = opn.BatchDecoder(ns, cfg, **SYNTH), opn.BatchDecoder(ns, cfg, **SYNTH)
    picks = (0, 31, 95)
    singles = {s: opn.Decoder(cfg, **SYNTH) for s in picks}
    oracle = {s: (O.SynthStream(lm, channels), np.zeros(2, np.float32)) for s in picks}
    offsets = np.arange(ns, dtype=np.uint32) * pkt_bytes
    lens = np.full(ns, pkt_bytes, np.uint32)
    beyond = 0
    for f in range(nfr):
        out = np.zeros((ns, nf * channels), dtype)
        res, _ = batch.decode_pcm(packets[f].reshape(-1), offsets, lens, out, nf)
        assert np.all(res == nf)
        out_f = np.zeros((ns, nf * channels), np.float32)
        res, _ = batch_f.decode_pcm(packets[f].reshape(-1), offsets, lens, out_f, nf)
        assert np.all(res == nf)
        exact = np.zeros((ns, nf * channels), dtype)
        assert O.lib().orc_sample_from_f32(fmt, O.ptr(out_f), O.ptr(exact), exact.size) == 0
        assert np.array_equal(out, exact), (f, np.argwhere(out != exact)[:3].tolist())
        for s in picks:
            one = np.zeros(nf * channels, dtype)
            assert singles[s].decode(packets[f, s], one, nf) == nf
            assert np.array_equal(one, out[s]), (f, s)
            st, mem = oracle[s]
            w = st.decode(packets[f, s, 1:])[3] * g
            beyond += int((np.abs(w) > 1.0).sum())
            O.lib().orc_pcm_soft_clip(O.ptr(w), nf, channels, O.ptr(mem), 2)  # the crate's slice quirk: nf, not nf*channels
            want = np.zeros(nf * channels, dtype)
            assert O.lib().orc_sample_from_f32(fmt, O.ptr(w), O.ptr(want), want.size) == 0
            # north_star's 1e-5 PCM bar, times the 4x gain applied after it; at least one step of an integer type
            tol = 4e-5 * full_scale if dtype == np.float64 else max(1.0, 4e-5 * full_scale)
            assert np.abs(out[s].astype(np.float64) - want.astype(np.float64)).max() <= tol, (f, s)
    assert beyond > 0, "the test signal never exceeded full scale"


def test_decode_pcm_rejects_unknown_format_and_short_buffers():
    dec = opn.Decoder(opn.DecoderConfiguration(48000, 2, 0), **SYNTH)
    pkt, _ = opn.synth_packet(5, 0, 3, 2, 160)
    b = np.frombuffer(bytes(pkt), np.uint8)
    out = np.zeros(1920, np.int32)
    rc = opn.lib().opn_decode_pcm(dec._h, b.ctypes.data_as(C.c_void_p), b.size, out.ctypes.data_as(C.c_void_p), out.size, 17, 960, 0)
    assert rc == -1  # BadArguments
    short = np.zeros(1000, np.int32)  # >= 960 (the crate's per-channel check passes) but < 960*2: Rust would panic on the index
    with pytest.raises(opn.OpusError) as e:
        dec.decode(pkt, short, 960)
    assert e.value.code == -2
    assert opn.lib().opn_sample_size(5) == 8 and opn.lib().opn_sample_size(3) == 2 and opn.lib().opn_sample_size(-1) == 0


# ------------------------------------------------------------------ full-size properties (BASELINE config 2)
def test_config2_full_size_properties():
    """4096 CELT FB 20 ms stereo streams @64 kbps, 3 chained frames: checksum of final ranges and the
    last frame's PCM for all streams against the multithreaded oracle; range state is a function of the
    packet only, PCM depends on the whole chain (carry + comb history)."""
    ns, nfr, lm, channels, pkt_bytes, nf = 4096, 3, 3, 2, 160, 960
    packets = opn.synth_fill(0, ns, 0, nfr, lm, channels, pkt_bytes)
    want_pcm = np.zeros((ns, nf * channels), np.float32)
    x = C.c_uint32(0)
    O.lib().orc_synth_bench(O.ptr(packets), ns, nfr, pkt_bytes, lm, channels, 1, os.cpu_count() or 1, O.ptr(want_pcm), C.byref(x))
    dec = opn.BatchDecoder(ns, **SYNTH)
    offsets = np.arange(ns, dtype=np.uint32) * pkt_bytes
    lens = np.full(ns, pkt_bytes, np.uint32)
    pcm = np.zeros((ns, nf * channels), np.float32)
    acc = np.uint32(0)
    for f in range(nfr):
        res = dec.decode_float(packets[f].reshape(-1), offsets, lens, pcm, nf)
        assert np.all(res == nf)
        acc ^= np.bitwise_xor.reduce(dec.final_ranges())
    assert int(acc) == x.value
    assert_pcm(want_pcm, pcm)
    assert np.array_equal(pcm, want_pcm)
    rms = math.sqrt(float((pcm.astype(np.float64) ** 2).mean()))
    assert 0.01 < rms < 1.0  # the 1e-5 bar is applied on +-1-scale PCM


def test_config2_full_size_device_resident_grouped_pipeline():
    """The benchmark's own path at its own size: 4096 stereo streams, device-resident packets, six steps enqueued back to back
    (range decode up to eight steps ahead, the frame kernel of every step as three concurrent launches on three streams).
    Every step's PCM of ALL streams against the multithreaded oracle decoding the same chain, bit for bit."""
    torch = pytest.importorskip("torch")
    ns, nfr, lm, channels, pkt_bytes, nf = 4096, 6, 3, 2, 160, 960
    packets = opn.synth_fill(9, ns, 0, nfr, lm, channels, pkt_bytes, transient_permille=100)
    dev = torch.device("cuda:0")
    d_arena = torch.from_numpy(packets.reshape(-1).copy()).to(dev)
    d_off = torch.arange(ns, dtype=torch.int32, device=dev) * pkt_bytes
    d_len = torch.full((ns,), pkt_bytes, dtype=torch.int32, device=dev)
    d_pcm = torch.zeros((nfr, ns, nf * channels), dtype=torch.float32, device=dev)
    d_res = torch.zeros((nfr, ns), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    dec = opn.BatchDecoder(ns, **SYNTH)
    flags = opn.FLAG_DEVICE_PTRS | opn.FLAG_INPUTS_READY
    for f in range(nfr):
        dec.decode_float_ptrs(d_arena.data_ptr() + f * ns * pkt_bytes, d_off.data_ptr(), d_len.data_ptr(), d_pcm[f].data_ptr(),
                              nf * channels, nf, d_res[f].data_ptr(), flags)
    dec.join()
    dec.synchronize()
    assert np.all(d_res.cpu().numpy() == nf)
    got = d_pcm.cpu().numpy()
    want = np.zeros((ns, nf * channels), np.float32)
    x = C.c_uint32(0)
    for k in range(1, nfr + 1):  # the oracle's chained decode of the first k frames leaves frame k's PCM
        O.lib().orc_synth_bench(O.ptr(packets[:k]), ns, k, pkt_bytes, lm, channels, 1, os.cpu_count() or 1, O.ptr(want), C.byref(x))
        assert np.array_equal(got[k - 1], want), k


def test_cpp_mirror_example_single_equals_batch():
    """examples/decode_batch.cpp through include/opusb200.hpp: BatchDecoder with two calls in flight and a
    single-stream Decoder agree bit for bit on stream 0, from compiled host code (no Python in the path)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "examples")], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(root, "examples", "decode_batch"), "1500", "5"], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "single==batch:yes" in r.stdout
