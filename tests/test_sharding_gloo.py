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


def _assemble_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from glomargridding_b200.distributed import assemble_obs_system, local_rhs

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    N, n = 37, 11
    A = rng.standard_normal((N, N))
    C = torch.from_numpy(A @ A.T)
    idx = torch.from_numpy(np.sort(rng.choice(N, n, replace=False)))
    E = torch.from_numpy(np.diag(rng.uniform(0.1, 0.2, n)))
    bounds = [0, 20, N]                      # uneven row blocks
    r0, r1 = bounds[rank], bounds[rank + 1]
    S = assemble_obs_system(C[r0:r1].contiguous(), r0, idx, E.numpy() if rank == 0 else E)
    T = local_rhs(C[r0:r1].contiguous(), idx, torch.arange(n, dtype=torch.float64))
    ok = bool(torch.equal(S, C[idx][:, idx] + E)) and bool(torch.equal(T[:, : r1 - r0], C[idx][:, r0:r1])) \
        and bool((T[:, r1 - r0] == 1).all()) and bool(torch.equal(T[:, r1 - r0 + 1], torch.arange(n, dtype=torch.float64)))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_obs_system_assembly_gloo_world2():
    """The exchange step of the row-block sharded kriging path (distributed.py) on CPU tensors over gloo."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_assemble_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results)
