"""Multi-GPU host logic on CPU: row-block and month sharding, and the gloo all_gather that reassembles per-rank
results (world_size 2, spawned processes, 127.0.0.1 rendezvous).  No CUDA is touched."""

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_row_blocks_partition_the_matrix_on_tile_boundaries():
    import bench

    for N in (64800, 2592, 2600, 100):
        for world in (1, 2, 4, 8):
            blocks = [bench.row_block(N, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(b[1] for b in blocks) == N
            for (r0, n0), (r1, _) in zip(blocks, blocks[1:]):
                assert r0 + n0 == r1 and r1 % bench.TILE == 0
            sizes = [b[1] for b in blocks]
            assert max(sizes) - min(sizes) <= bench.TILE or world > N // bench.TILE


def test_month_sharding_covers_every_month_once():
    from glomargridding_b200.batched import shard

    for n, world in [(480, 1), (480, 8), (7, 4), (3, 8)]:
        got = sorted(i for r in range(world) for i in shard(n, r, world))
        assert got == list(range(n))
        assert max(len(shard(n, r, world)) for r in range(world)) - min(len(shard(n, r, world)) for r in range(world)) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, M, q):
    import torch.distributed as dist

    from glomargridding_b200.batched import gather_results, shard

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard(n_items, rank, world)
    local = np.array([[100.0 * i + j for j in range(M)] for i in mine]).reshape(len(mine), M)
    full = gather_results(local, n_items, rank, world)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 8])
def test_gather_results_gloo_world2(n_items):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, world, M = _free_port(), 2, 5
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, M, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.array([[100.0 * i + j for j in range(M)] for i in range(n_items)])
    for _, full in results:
        assert np.array_equal(full, want)
