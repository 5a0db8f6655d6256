"""Multi-GPU host logic on CPU: sharding by matrix index, LPT partition for variable sizes, and a
world-size-2 gloo run in which every rank factors ITS shard (with the oracle standing in for the GPU,
which this box does not have) and the gathered pivots equal the single-process result."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from magma_b200 import mgpu


def test_shard_ranges_cover_the_batch_once():
    for batch in (0, 1, 7, 8, 1000001):
        for world in (1, 2, 3, 8):
            seen = 0
            prev = 0
            for r in range(world):
                lo, hi = mgpu.shard_range(batch, world, r)
                assert lo == prev and hi >= lo
                prev = hi
                seen += hi - lo
            assert seen == batch and prev == batch
            sizes = [np.diff(mgpu.shard_range(batch, world, r))[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_lpt_partition_balances_config4_sizes():
    # BASELINE config 4: sizes 16 + (lcg mod 497), 20,000 matrices (bench.py run_sweep)
    x, ns = 1234, []
    for _ in range(20000):
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        ns.append(16 + (x >> 8) % 497)
    ns = np.array(ns)
    for world in (2, 4, 8):
        parts = mgpu.lpt_partition(ns, ns, world)
        assert sorted(np.concatenate(parts).tolist()) == list(range(len(ns)))
        assert mgpu.imbalance(ns, ns, parts) < 1.001
        # a contiguous split of the same batch is measurably worse
        naive = [np.arange(*mgpu.shard_range(len(ns), world, r)) for r in range(world)]
        assert mgpu.imbalance(ns, ns, parts) <= mgpu.imbalance(ns, ns, naive)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle
    from magma_b200 import mgpu as mg
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, batch = 12, 101
    A, _ = oracle.random_batch(batch, n, n)      # every rank draws the same global stream
    lo, hi = mg.shard_range(batch, world, rank)
    mine = A[lo:hi].copy()
    ipiv, info = oracle.getrf_batched(mine, n)    # stand-in for the GPU tier: same arithmetic, same pivots
    # gather (outside any timed region): pivots of every shard, padded to the largest shard
    width = max(np.diff(mg.shard_range(batch, world, r))[0] for r in range(world))
    pad = np.zeros((width, n), dtype=np.int32)
    pad[:hi - lo] = ipiv
    bufs = [torch.zeros((width, n), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(pad))
    t = mg.max_over_ranks(float(rank + 1))        # "slowest rank" reduction used by the bench
    s = mg.sum_over_ranks(float(hi - lo))
    if rank == 0:
        full = A.copy()
        ref, _ = oracle.getrf_batched(full, n)
        got = np.concatenate([bufs[r].numpy()[:np.diff(mg.shard_range(batch, world, r))[0]] for r in range(world)])
        np.save(os.path.join(out_dir, "ok.npy"), np.array([np.array_equal(got, ref), t == world, s == batch]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_run(tmp_path):
    import torch.multiprocessing as mp
    import socket
    with socket.socket() as sk:  # a port the kernel says is free right now (a pid-derived one collided once with a live process)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(os.path.join(tmp_path, "ok.npy"))
    assert ok.all(), ok
