"""CPU, world_size 2 over gloo: the host-side logic of the N>1 paths --
candidate sharding, flat-bucket all-reduce, rank-ordered all-gather, final EA gather."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nasrec_b200.parallel import allgather_cat, allreduce_flat, gather_results, shard_range


def test_shard_range_partitions_contiguously():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(1024, 8, 3) == (384, 512)          # BASELINE config #3: 128 candidates per GPU


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # dense gradients: one flat bucket, views replace the inputs
        g = [torch.full((3, 4), float(rank + 1)), torch.arange(5, dtype=torch.float32) * (rank + 1),
             torch.tensor(2.0 * (rank + 1))]
        out = allreduce_flat(g)
        tot = sum(r + 1 for r in range(world))
        assert [tuple(o.shape) for o in out] == [(3, 4), (5,), ()]
        assert torch.equal(out[0], torch.full((3, 4), float(tot)))
        assert torch.equal(out[1], torch.arange(5, dtype=torch.float32) * tot)
        assert float(out[2]) == 2.0 * tot
        assert out[0].data_ptr() + 12 * 4 == out[1].data_ptr()        # one contiguous bucket
        # sparse gradients: ids and rows concatenated in rank order
        ids = torch.arange(6, dtype=torch.int64).reshape(3, 2) + 100 * rank
        rows = torch.full((3, 2, 16), float(rank))
        gi, gr = allgather_cat(ids), allgather_cat(rows)
        assert gi.shape == (3 * world, 2) and gr.shape == (3 * world, 2, 16)
        for r in range(world):
            assert torch.equal(gi[3 * r:3 * r + 3], torch.arange(6).reshape(3, 2) + 100 * r)
            assert float(gr[3 * r:3 * r + 3].mean()) == float(r)
        # EA: contiguous shards, ragged sizes, gathered in candidate order on rank 0
        n_total = 5
        lo, hi = shard_range(n_total, world, rank)
        local = [[float(c), 0.5 + c, 0.25 * c] for c in range(lo, hi)]
        res = gather_results(local, n_total)
        if rank == 0:
            assert res.shape == (n_total, 3)
            assert res[:, 0].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0]
            assert res[3].tolist() == [3.0, 3.5, 0.75]
        else:
            assert res is None
        q.put((rank, "ok"))
    except Exception as e:          # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_collective_helpers_world_size_2_gloo():
    world, port = 2, 29500 + os.getpid() % 400
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(got) == [(0, "ok"), (1, "ok")], got
