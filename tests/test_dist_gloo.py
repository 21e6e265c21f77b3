"""World-size-2 gloo test (CPU) of the multi-GPU plumbing: byte-balanced contiguous sharding, order-preserving
reassembly, max-over-ranks timing.  The per-rank "codec" here is the oracle (a checker standing in for the GPU
so the N>1 host logic can be exercised without one)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, lens, q):
    import torch.distributed as dist
    from conftest import Oracle, build_oracle
    from slow5tools_b200 import synth
    from slow5tools_b200.dist import max_over_ranks, shard_bounds, sum_over_ranks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle(build_oracle())
    bounds = shard_bounds(lens, world)
    sig = synth.nanopore_signal(int(np.sum(lens)), seed=11).numpy()      # every rank can regenerate the batch
    starts = np.concatenate([[0], np.cumsum(lens)])
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = [orc.compress(sig[starts[i]:starts[i + 1]]) for i in range(lo, hi)]
    n_all = sum_over_ranks(len(mine))
    slowest = max_over_ranks(float(rank + 1))
    both = max_over_ranks([float(rank), 10.0 - rank])
    dist.barrier()
    q.put((rank, lo, hi, mine, n_all, slowest, both))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reassembly():
    from conftest import Oracle, build_oracle
    from slow5tools_b200 import synth
    from slow5tools_b200.dist import shard_bounds
    rng = np.random.default_rng(3)
    lens = np.clip(rng.lognormal(np.log(3000), 1.0, 60), 50, 20000).astype(np.int64)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, lens, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bounds = shard_bounds(lens, world)
    assert [g[1] for g in got] == bounds[:-1] and [g[2] for g in got] == bounds[1:]
    # byte balance: no rank holds more than the ideal share plus one read
    share = [int(np.sum(lens[bounds[r]:bounds[r + 1]])) for r in range(world)]
    assert max(share) <= np.sum(lens) / world + lens.max()
    # concatenating rank outputs in rank order rebuilds the batch exactly
    orc = Oracle(build_oracle())
    sig = synth.nanopore_signal(int(np.sum(lens)), seed=11).numpy()
    starts = np.concatenate([[0], np.cumsum(lens)])
    want = [orc.compress(sig[starts[i]:starts[i + 1]]) for i in range(len(lens))]
    assert sum((g[3] for g in got), []) == want
    assert all(g[4] == len(lens) and g[5] == 2.0 and g[6] == [1.0, 10.0] for g in got)


def test_shard_bounds_edge_cases():
    from slow5tools_b200.dist import shard_bounds
    assert shard_bounds([], 4) == [0, 0, 0, 0, 0]
    assert shard_bounds([5], 2)[0] == 0 and shard_bounds([5], 2)[-1] == 1
    b = shard_bounds([1] * 8, 8)
    assert b == list(range(9))
    b = shard_bounds([100, 1, 1, 1, 1, 100], 2)
    assert b[0] == 0 and b[-1] == 6 and 1 <= b[1] <= 5
