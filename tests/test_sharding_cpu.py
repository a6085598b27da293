"""CPU (gloo, world_size 2): the host-side logic of the sharded path -- contiguous token-balanced document partition,
agreement across ranks, and the gather that rebuilds the global document order."""
import os

import numpy as np
import pytest

from bpe_tokenizer_b200.sharded import shard_bounds


def test_shard_bounds_partition_properties():
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 8):
        for _ in range(50):
            sizes = rng.integers(0, 40, size=rng.integers(0, 60)).tolist()
            b = shard_bounds(sizes, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == len(sizes)
            assert all(b[i] <= b[i + 1] for i in range(world))
            total = sum(sizes)
            if total and world > 1:
                loads = [sum(sizes[b[r]:b[r + 1]]) for r in range(world)]
                assert max(loads) <= total / world + max(sizes)  # balanced up to one document
    assert shard_bounds([], 4) == [0, 0, 0, 0, 0]
    assert shard_bounds([5, 5, 5, 5], 2) == [0, 2, 4]
    assert shard_bounds([0, 0, 3, 0], 2) == [0, 3, 4]  # the document holding the cut token starts the next shard


def _gloo_worker(rank, world, port, sizes):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        docs = [np.arange(n, dtype=np.int32) + 1000 * i for i, n in enumerate(sizes)]
        b = shard_bounds(sizes, world)
        mine = docs[b[rank]:b[rank + 1]]
        ids = np.concatenate(mine) if mine and sum(d.size for d in mine) else np.zeros(0, dtype=np.int32)
        off = np.zeros(len(mine) + 1, dtype=np.int64)
        np.cumsum([d.size for d in mine], out=off[1:])
        parts = [None] * world
        dist.all_gather_object(parts, (ids, off))
        # same reconstruction as ShardedBPETokenizer.corpusIdsAllRanks
        all_ids = np.concatenate([p[0] for p in parts])
        offs, base = [np.zeros(1, dtype=np.int64)], 0
        for p in parts:
            offs.append(p[1][1:] + base)
            base += int(p[1][-1])
        offs = np.concatenate(offs)
        want = np.concatenate(docs) if sum(sizes) else np.zeros(0, dtype=np.int32)
        assert np.array_equal(all_ids, want)
        assert offs.tolist() == np.concatenate([[0], np.cumsum(sizes)]).tolist()
        bounds = [None] * world
        dist.all_gather_object(bounds, b)
        assert all(x == b for x in bounds)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_partition_and_gather_world2_gloo():
    import torch.multiprocessing as mp

    mp.spawn(_gloo_worker, args=(2, 29541, [7, 0, 3, 12, 1, 0, 9, 4]), nprocs=2, join=True)
