"""The N > 1 path on CPU: two processes (gloo), each taking its shard of a
stream table through bjxa_shard_range, then the same reductions bench.py does
(max-over-ranks of the elapsed time, gathered per-rank records).  There is no
data-path collective to test: shards are independent."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bjxa_b200 import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, sizes, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bjxa_b200
        lib = bjxa_b200.load()
        first, count = lib.shard_range(len(sizes), rank, world, sizes)
        # every rank plans only its own shard (host logic, no GPU needed)
        my = sizes[first:first + count]
        rec = torch.tensor([first, count, int(my.sum())], dtype=torch.int64)
        gathered = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, rec)
        elapsed = torch.tensor([10.0 + rank], dtype=torch.float64)   # slowest rank wins
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        dist.barrier()
        if rank == 0:
            np.save(os.path.join(out_dir, "gathered.npy"),
                    torch.stack(gathered).numpy())
            np.save(os.path.join(out_dir, "elapsed.npy"), elapsed.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_ranks_shard_the_stream_table(tmp_path, world):
    sizes = (synth.rand_u64(3, 3, 4096) % np.uint64(5_000_000)).astype(np.uint64) + 1000
    port = _free_port()
    mp.spawn(_worker, args=(world, port, sizes, str(tmp_path)), nprocs=world, join=True)
    g = np.load(tmp_path / "gathered.npy")
    # contiguous, exhaustive, disjoint
    assert g[0, 0] == 0 and g[0, 0] + g[0, 1] == g[1, 0] and g[1, 0] + g[1, 1] == len(sizes)
    # balanced by bytes within one stream
    assert abs(int(g[0, 2]) - int(g[1, 2])) <= 2 * int(sizes.max())
    assert int(g[:, 2].sum()) == int(sizes.sum())
    assert float(np.load(tmp_path / "elapsed.npy")[0]) == 10.0 + world - 1


@pytest.mark.parametrize("world", [2, 4, 8])
def test_corpus_table_shards_within_three_percent(world):
    """bench.py's configs[4]: the 262 144-stream corpus table (the same on every
    rank) cut by bjxa_shard_range(bytes[]) -- contiguous, exhaustive, and no shard
    more than 3 % above the mean."""
    import bench
    import bjxa_b200
    lib = bjxa_b200.load()
    for tab in (bench.VarTable(5, 262144, [4, 6, 8], [2], 0.25, 4.0),
                bench.VarTable(6, 262144, [4, 6, 8], [1, 2], 0.25, 4.0)):
        again = bench.VarTable(5, 262144, [4, 6, 8], [2], 0.25, 4.0)
        assert tab.algo.sum() > 0 and np.array_equal(
            again.algo, bench.VarTable(5, 262144, [4, 6, 8], [2], 0.25, 4.0).algo)
        nxt, per = 0, []
        for r in range(world):
            first, count = lib.shard_range(tab.n, r, world, tab.algo)
            assert first == nxt
            nxt = first + count
            per.append(int(tab.algo[first:first + count].sum()))
        assert nxt == tab.n
        assert max(per) / (sum(per) / world) - 1 <= 0.03
