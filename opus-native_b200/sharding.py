"""Stream -> GPU placement.  Streams never communicate, so multi-GPU decode is a
static partition of the stream ids with no collective on the data path: rank r of W owns the
contiguous range shard_range(n, r, W).  Placement is sticky because each stream's overlap carry,
comb history and post-filter parameters live in that GPU's HBM."""


def shard_range(n_streams: int, rank: int, world: int):
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = n_streams * rank // world
    hi = n_streams * (rank + 1) // world
    return lo, hi
