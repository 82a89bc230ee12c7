"""Device-buffer plumbing (PyTorch owns every allocation; the C ABI only sees raw pointers)."""

from __future__ import annotations

import numpy as np
import torch

from ._lib import GlomarB200Error


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise GlomarB200Error("glomargridding_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def host_threads(cap: int) -> int:
    """Worker threads for host-side packing / bookkeeping: the cores of the box divided by the ranks that share
    it (`LOCAL_WORLD_SIZE`, set by torchrun - one process per GPU), one left free when a rank has more than two, at
    most `cap`.  (The thread that waits for the pool is idle, so two cores per rank give two packing threads.)"""
    import os

    try:
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:
        ranks = 1
    share = (os.cpu_count() or 2) // ranks
    return max(1, min(cap, share - 1 if share > 2 else share))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def to_device(a, dtype=torch.float64) -> torch.Tensor:
    """numpy / torch (any device) -> contiguous CUDA tensor of `dtype` on the current device."""
    dev = require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    arr = np.ascontiguousarray(np.ma.getdata(a))
    return torch.from_numpy(arr).to(device=dev, dtype=dtype, non_blocking=False).contiguous()


def to_host(t: torch.Tensor, pinned: bool = True) -> np.ndarray:
    """CUDA tensor -> numpy array.  Large results land in pinned host memory (torch's caching host
    allocator reuses the block across calls), which is what keeps D2H at PCIe speed."""
    if t.device.type != "cuda":
        return t.numpy()
    if pinned and t.numel() * t.element_size() >= (1 << 20):
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()
    return t.cpu().numpy()


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()
