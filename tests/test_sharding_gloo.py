"""Multi-GPU plan on CPU: world_size-2 gloo run of the sharding + max-over-ranks timing plumbing
bench.py uses at N > 1.  Streams are independent: each rank generates and owns a
disjoint contiguous range of global stream ids and no data-path collective exists -- the only
collectives are the barrier and the MAX reduction of the elapsed time."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import opus_native_b200 as opn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_per_rank, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = opn.shard_range(n_per_rank * world, rank, world)
    assert hi - lo == n_per_rank
    packets = opn.synth_fill(lo, n_per_rank, 0, 2, 3, 2, 160, n_threads=1)
    # every rank decodes only its own streams (here with the oracle: no GPU on this box)
    rngs = []
    for s in range(n_per_rank):
        st = O.SynthStream(3, 2)
        for f in range(2):
            side = st.decode(packets[f, s, 1:])[0]
        rngs.append(side.final_rng)
    dist.barrier()
    t = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert abs(t.item() - 0.010 * world) < 1e-12
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([lo, hi] + rngs, dtype=np.uint64))
    dist.destroy_process_group()


def test_two_rank_sharding_is_disjoint_and_complete(tmp_path):
    world, n_per_rank = 2, 6
    mp.spawn(_worker, args=(world, _free_port(), n_per_rank, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    assert [int(g[0]) for g in got] == [0, 6] and [int(g[1]) for g in got] == [6, 12]
    # the union equals a single-process decode of the 12 global streams
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    packets = opn.synth_fill(0, 12, 0, 2, 3, 2, 160, n_threads=1)
    want = []
    for s in range(12):
        st = O.SynthStream(3, 2)
        for f in range(2):
            side = st.decode(packets[f, s, 1:])[0]
        want.append(side.final_rng)
    assert [int(x) for g in got for x in g[2:]] == want


def test_shard_range_properties():
    for n in (1, 7, 4096, 65536, 262144):
        for w in (1, 2, 4, 8):
            edges = [opn.shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in edges) - min(h - l for l, h in edges) <= 1
