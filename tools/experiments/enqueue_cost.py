"""Host time to enqueue one device-resident step (all CUDA calls of the per-step stream DAG) versus the GPU
time of the step.  If the two are close, the host is the limit (and a CUDA-graph replay would pay)."""
import time

import torch

import opus_native_b200 as opn

import sys

n, pkt, nf = 4096, 160, 960
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40  # keep below the launch-queue depth or the host blocks on the GPU
packets = opn.synth_fill(0, n, 0, steps + 10, 3, 2, pkt, 0, n_threads=8)
dev = torch.device("cuda:0")
d_arena = torch.from_numpy(packets.reshape(-1)).to(dev)
d_off = (torch.arange(n, dtype=torch.int64, device=dev) * pkt).to(torch.int32)
d_len = torch.full((n,), pkt, dtype=torch.int32, device=dev)
d_res = torch.zeros(n, dtype=torch.int32, device=dev)
dec = opn.BatchDecoder(n, bitstream=opn.BITSTREAM_SYNTH_CELT_1)
flags = opn.FLAG_DEVICE_PTRS | opn.FLAG_NO_PCM_COPY | opn.FLAG_INPUTS_READY
for f in range(10):
    dec.decode_float_ptrs(d_arena.data_ptr() + f * n * pkt, d_off.data_ptr(), d_len.data_ptr(), None, 0, nf, d_res.data_ptr(), flags)
dec.synchronize()
t0 = time.perf_counter()
for f in range(10, 10 + steps):
    dec.decode_float_ptrs(d_arena.data_ptr() + f * n * pkt, d_off.data_ptr(), d_len.data_ptr(), None, 0, nf, d_res.data_ptr(), flags)
t1 = time.perf_counter()
dec.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e6 * (t1 - t0) / steps:.1f} us/step on the host; all done after {1e6 * (t2 - t0) / steps:.1f} us/step")
