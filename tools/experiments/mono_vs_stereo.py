"""Kernel times for 8192 mono streams versus 4096 stereo streams (the same number of channel-frames):
what one-channel-per-warp costs/gains in kernels 1 and 2 with the current code."""
import numpy as np
import torch

import opus_native_b200 as opn

for channels, n, pkt in ((2, 4096, 160), (1, 8192, 100)):
    steps, lm, nf = 40, 3, 960
    packets = opn.synth_fill(0, n, 0, steps + 5, lm, channels, pkt, 0, n_threads=8)
    dev = torch.device("cuda:0")
    d_arena = torch.from_numpy(packets.reshape(-1)).to(dev)
    d_off = (torch.arange(n, dtype=torch.int64, device=dev) * pkt).to(torch.int32)
    d_len = torch.full((n,), pkt, dtype=torch.int32, device=dev)
    d_res = torch.zeros(n, dtype=torch.int32, device=dev)
    dec = opn.BatchDecoder(n, opn.DecoderConfiguration(48000, channels, 0), bitstream=opn.BITSTREAM_SYNTH_CELT_1)
    flags = opn.FLAG_DEVICE_PTRS | opn.FLAG_NO_PCM_COPY | opn.FLAG_INPUTS_READY
    dec.enable_timing(True)
    for f in range(5):
        dec.decode_float_ptrs(d_arena.data_ptr() + f * n * pkt, d_off.data_ptr(), d_len.data_ptr(), None, 0, nf, d_res.data_ptr(), flags)
    dec.stats(reset=True)
    for f in range(5, 5 + steps):
        dec.decode_float_ptrs(d_arena.data_ptr() + f * n * pkt, d_off.data_ptr(), d_len.data_ptr(), None, 0, nf, d_res.data_ptr(), flags)
    st = dec.stats(reset=True)
    assert int((d_res != nf).sum().item()) == 0
    print(f"channels={channels} streams={n}: symbols {1e3 * st['ms'][0] / steps:.1f} us, kernel 1 {1e3 * st['ms'][1] / steps:.1f} us, "
          f"kernel 2 {1e3 * st['ms'][2] / steps:.1f} us")
