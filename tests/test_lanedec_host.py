"""CPU test of the DEVICE range decoder source: csrc/rangedec.cuh `LaneDec` (the branch-free decoder k_synth_rangedec
runs, one lane per packet) is compiled for the host by tests/host_shim/lanedec_host.cpp, with the CUDA intrinsics it
uses given host bodies, and replays range-coder scripts against the oracle (src/range_coder/decoder.rs restated):
every symbol, tell_frac and rng after every call, on encoded streams, truncated streams and random bytes, at every
alignment of the packet inside its 32-bit words (the decoder reads whole aligned words and masks the neighbours off)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ICDF_POOL = np.array([30, 22, 15, 8, 3, 0, 2, 1, 0], np.uint8)
OP_UINT, OP_BITS, OP_BIT_LOGP, OP_ICDF, OP_LAPLACE = 0, 1, 2, 3, 4


@pytest.fixture(scope="module")
def shim():
    so = os.path.join(HERE, "host_shim", "liblanedec_host.so")
    src = os.path.join(HERE, "host_shim", "lanedec_host.cpp")
    hdrs = [os.path.join(ROOT, "opus-native_b200", "csrc", h) for h in ("rangedec.cuh", "celt2.cuh", "celt2_lane.cuh", "mathops.cuh", "opn_tables.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in [src] + hdrs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    L = C.CDLL(so)
    L.lanedec_run_script.restype = C.c_int
    L.lanedec_run_script.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.lanedec_celt2_decode.restype = C.c_int
    L.lanedec_celt2_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return L


def _script(rnd, n_ops):
    ops, vals = [], []
    for _ in range(n_ops):
        kind = int(rnd.integers(0, 5))
        if kind == 0:
            ft = int(rnd.choice([2, 3, 6, 255, 256, 257, 4066763520, 0xFFFFFFFF, int(rnd.integers(2, 2 ** int(rnd.integers(2, 33)) - 1))]))
            ops.append((OP_UINT, ft, int(rnd.integers(0, 2)))); vals.append(int(rnd.integers(0, ft)))  # b = 1: LaneDec::uint_any
        elif kind == 1:
            nb = int(rnd.integers(1, 26))
            ops.append((OP_BITS, nb, 0)); vals.append(int(rnd.integers(0, 1 << nb)))
        elif kind == 2:
            ops.append((OP_BIT_LOGP, int(rnd.integers(1, 16)), 0)); vals.append(int(rnd.integers(0, 2)))
        elif kind == 3:
            tab = [(0, 6, 5), (6, 3, 2)][int(rnd.integers(0, 2))]
            ops.append((OP_ICDF, tab[0], tab[2])); vals.append(int(rnd.integers(0, tab[1])))
        else:
            decay = int(rnd.integers(5000, 16000)) if rnd.random() < 0.5 else 6000 + 400 * int(rnd.integers(0, 21))
            v = int(rnd.integers(-12, 13))  # what the encoder can always represent; garbage packets reach beyond the 16 tabulated magnitudes
            ops.append((OP_LAPLACE, O.lib().orc_laplace_start_freq(decay), decay)); vals.append(v & 0xFFFFFFFF)
    return np.array(ops, O.OP_DTYPE), vals


def _run(L, packet, ops):
    """Every alignment; poison bytes around the packet must not leak in."""
    want, _ = O.dec_run_script(packet, ops, ICDF_POOL)
    n = len(packet)
    for align in range(4):
        arena = np.zeros(n + 16, np.uint8)
        off = (-arena.ctypes.data) % 4 + 4 + align
        arena[:off], arena[off + n:] = 0xAA, 0x55
        arena[off:off + n] = packet
        out = np.zeros(len(ops), O.OUT_DTYPE)
        assert L.lanedec_run_script(arena.ctypes.data + off, n, ops.ctypes.data, len(ops), ICDF_POOL.ctypes.data, out.ctypes.data) == 0
        bad = np.nonzero(out != want)[0]
        assert bad.size == 0, (align, n, int(bad[0]), ops[bad[0]], out[bad[0]], want[bad[0]])


def test_lanedec_encoded_streams_every_alignment(shim):
    rnd = np.random.default_rng(21)
    for trial in range(150):
        ops, vals = _script(rnd, int(rnd.integers(4, 160)))
        nbytes = 4 * len(ops) + 8 + int(rnd.integers(0, 40))  # every op fits: at most 32 bits each
        buf, _, _, _, err = O.enc_run_script(nbytes, ops, vals, ICDF_POOL)
        assert err == 0
        _run(shim, buf, ops)


def test_lanedec_truncated_and_garbage_packets(shim):
    """Zero extension past `storage` on both ends, decode_uint saturation (decoder.rs:86-104, 255-259), magnitudes beyond
    the Laplace table."""
    rnd = np.random.default_rng(22)
    for trial in range(150):
        ops, vals = _script(rnd, int(rnd.integers(20, 200)))
        nbytes = int(rnd.integers(1, 200))
        if trial % 3 == 0:
            full, _, _, _, err = O.enc_run_script(4 * len(ops) + 8, ops, vals, ICDF_POOL)
            assert err == 0
            pkt = full[:nbytes].copy()
        else:
            pkt = rnd.integers(0, 256, nbytes).astype(np.uint8)
        _run(shim, pkt, ops)


def test_lanedec_synth_celt_1_symbol_sequence(shim):
    """The symbol sequence k_synth_rangedec decodes on real SYNTH-CELT/1 payloads."""
    lm, channels, pkt_bytes = 3, 2, 160
    ops = [(OP_BIT_LOGP, 15, 0), (OP_BIT_LOGP, 1, 0)]
    pk = O.synth_fill(0, 24, 0, 2, lm, channels, pkt_bytes, 300, 1).reshape(-1, pkt_bytes)
    for p in pk:
        payload = p[1:]
        side = O.SynthStream(lm, channels).decode(payload)[0]
        ops = [(OP_BIT_LOGP, 15, 0), (OP_BIT_LOGP, 1, 0)]
        if side.postfilter:
            ops += [(OP_UINT, 6, 0), (OP_BITS, 4 + side.octave, 0), (OP_BITS, 3, 0), (OP_ICDF, 6, 2)]
        ops += [(OP_BIT_LOGP, 3, 0), (OP_BIT_LOGP, 3, 0)]
        for b in range(21):
            decay = 6000 + 400 * b
            ops += [(OP_LAPLACE, O.lib().orc_laplace_start_freq(decay), decay)] * channels
        ops += [(OP_BITS, 2, 0)] * (21 * channels)
        _run(shim, payload, np.array(ops, O.OP_DTYPE))


@pytest.mark.parametrize("lm,channels,pkt_bytes", [(3, 2, 160), (3, 1, 100), (2, 2, 130), (1, 1, 60), (0, 2, 80)])
def test_celt2_frame_decode_through_the_lane_coder(shim, lm, channels, pkt_bytes):
    """SYNTH-CELT/2: the device's decode-side coder (LaneCoder on LaneDec, celt2_lane.cuh) running celt2_frame (celt2.cuh),
    compiled for the host, against the oracle's independent restatement: the whole side record, on generated packets and
    on random bytes."""
    import opus_native_b200 as opn
    rnd = np.random.default_rng(lm * 10 + channels)
    for s in range(24):
        if s % 3 == 2:
            payload = rnd.integers(0, 256, int(rnd.integers(2, 200))).astype(np.uint8)
        else:
            payload = O.celt2_packet(s, 1, lm, channels, pkt_bytes, 300)[0][1:]
        buf = np.zeros(len(payload) + 12, np.uint8)
        off = (-buf.ctypes.data) % 4 + 4 + (s & 3)
        buf[off:off + len(payload)] = payload
        side = np.zeros(1, opn.CELT2_SIDE_DTYPE)
        parts = np.zeros(192 * 12, np.uint8)
        shim.lanedec_celt2_decode(buf.ctypes.data + off, len(payload), lm, channels, side.ctypes.data, parts.ctypes.data)
        w = O.celt2_decode_symbols(payload, lm, channels)[0]
        for f in opn.CELT2_SIDE_DTYPE.names:
            v = getattr(w, f)
            v = np.ctypeslib.as_array(v) if hasattr(v, "__len__") else v
            assert np.array_equal(side[0][f], v), (s, len(payload), f)
