"""Resident-loop step time in consecutive chunks of one process: is a slow run a persistent state of the pipeline or
a transient hiccup?  usage: PYTHONPATH=. python tools/experiments/step_jitter.py [chunks] [steps_per_chunk]"""
import sys

import torch

import opus_native_b200 as opn

n, pkt, nf = 4096, 160, 960
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
per = int(sys.argv[2]) if len(sys.argv) > 2 else 200
total = 10 + chunks * per
packets = opn.synth_fill(0, n, 0, min(total, 410), 3, 2, pkt, 0, n_threads=8)
nsets = packets.shape[0]
dev = torch.device("cuda:0")
d_arena = torch.from_numpy(packets.reshape(-1)).to(dev)
d_off = (torch.arange(n, dtype=torch.int64, device=dev) * pkt).to(torch.int32)
d_len = torch.full((n,), pkt, dtype=torch.int32, device=dev)
d_res = torch.zeros(n, dtype=torch.int32, device=dev)
dec = opn.BatchDecoder(n, bitstream=opn.BITSTREAM_SYNTH_CELT_1)
stream = torch.cuda.ExternalStream(dec.cuda_stream, device=dev)
flags = opn.FLAG_DEVICE_PTRS | opn.FLAG_NO_PCM_COPY | opn.FLAG_INPUTS_READY
pa, po, pl, pr = d_arena.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_res.data_ptr()
f = 0
for _ in range(10):
    dec.decode_float_ptrs(pa + (f % nsets) * n * pkt, po, pl, None, 0, nf, pr, flags)
    f += 1
dec.synchronize()
out = []
for c in range(chunks):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(per):
            dec.decode_float_ptrs(pa + (f % nsets) * n * pkt, po, pl, None, 0, nf, pr, flags)
            f += 1
        dec.join()
        e1.record()
    dec.synchronize()
    out.append(1e3 * e0.elapsed_time(e1) / per)
print("us/step per chunk of %d:" % per, " ".join("%.1f" % x for x in out))
