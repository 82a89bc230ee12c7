"""Peer-to-peer row blocks for the multi-GPU covariance build (one process per GPU, NVLink / NVSwitch).

`PeerRowBlocks` allocates this rank's row block of the N x N matrix, exchanges CUDA IPC handles with the other
ranks of the group and maps their blocks into this process, so that `gmg_ellipse_cov_p2p_f64` can store the
transposed half of every off-diagonal tile straight into the memory of the rank that owns those rows.  With it
every pair of grid points is evaluated once across the whole box (instead of once per row block that contains
it), which is what makes the build scale: per-rank FP64 work N^2 / (2 world) instead of
N^2 (2 world - 1) / (2 world^2).

PyTorch is only plumbing here: the allocation, `torch.distributed` for the 64-byte handles and the barriers.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from .._lib import check, lib

TILE = 64


def tile_aligned_blocks(N: int, world: int) -> np.ndarray:
    """Row-block boundaries (world + 1 entries), every boundary but the last a multiple of the 64-row tile."""
    tiles = (N + TILE - 1) // TILE
    b = [min(N, (tiles * r // world) * TILE) for r in range(world)] + [N]
    return np.asarray(b, dtype=np.int64)


class _CudaBlock:
    """Raw device allocation exposed through the CUDA array interface (so torch can view it without owning it)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}


class PeerRowBlocks:
    def __init__(self, N: int, dtype=torch.float64, group=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerRowBlocks needs torch.distributed (one process per GPU)")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.N = N
        self.blocks = tile_aligned_blocks(N, self.world)
        self.row0 = int(self.blocks[self.rank])
        self.nrows = int(self.blocks[self.rank + 1] - self.blocks[self.rank])
        dev = torch.device("cuda", torch.cuda.current_device())
        # An IPC handle names a whole cudaMalloc block; torch's caching allocator sub-allocates, so the block comes
        # from the library (gmg_ipc_alloc_f64) and is wrapped as a tensor through __cuda_array_interface__.
        itemsize = torch.empty(0, dtype=dtype).element_size()
        self._base = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        check(lib().gmg_ipc_alloc_f64(self.nrows * N * itemsize, ctypes.byref(self._base), handle),
              "gmg_ipc_alloc_f64")
        self._holder = _CudaBlock(self._base.value, (self.nrows, N), "<f8" if itemsize == 8 else "<f4")
        self.local = torch.as_tensor(self._holder, device=dev)
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(dev)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=group)
        self._opened = []
        self.ptrs = (ctypes.c_void_p * self.world)()
        for g in range(self.world):
            if g == self.rank:
                self.ptrs[g] = self.local.data_ptr()
                continue
            raw = allh[g].cpu().numpy().tobytes()
            base = ctypes.c_void_p()
            check(lib().gmg_ipc_open_f64(raw[:64], ctypes.byref(base)), "gmg_ipc_open_f64")
            self._opened.append(base.value)
            self.ptrs[g] = base.value
        self.block_row0 = (ctypes.c_int64 * (self.world + 1))(*[int(x) for x in self.blocks])
        self.barrier()

    def barrier(self) -> None:
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def close(self) -> None:
        self.barrier()
        for p in self._opened:
            lib().gmg_ipc_close_f64(p)
        self._opened = []
        self.barrier()                      # nobody maps this block any more
        if self._base.value:
            self.local = None
            self._holder = None
            lib().gmg_ipc_free_f64(self._base)
            self._base = ctypes.c_void_p()
