"""Kriging against a ROW-BLOCK SHARDED covariance (BASELINE.json configs[3] + configs[4]: the 1-degree covariance is
built in row blocks, one per GPU, and 20 000 observations are kriged against it without ever gathering the
33.6 GB matrix on one device).

One process per GPU (`torchrun`), `torch.distributed` for the plumbing.  Rank r owns rows [r0, r0 + nrows) of C.

  1. exchange - the only collective on the data path: every rank cuts C[idx_owned, idx] out of its own rows and
     one all_gather assembles S = C[idx, idx] (+ E) on every rank (n^2 8 bytes in total, 3.2 GB at n = 20 000).
  2. no exchange is needed for the right-hand sides: C is symmetric, so the columns [r0, r0 + nrows) of
     C_cross = C[idx, :] are the transpose of C[r0 : r0 + nrows, idx], which the rank already holds.  The grid
     columns - and with them the triangular solve, which is >85 % of the flops - shard across ranks for free.
  3. S is factored ONCE across the ranks (`_dense.factor_inplace_distributed`: 1-D block-cyclic tile rows, the
     owner factors the diagonal block and broadcasts it, every rank forms its rows of the panel, one NCCL all_reduce
     per block step hands everyone the whole panel, every rank updates its own tile rows) and the complete factor
     ends up on every rank; each rank then runs the forward solve on its columns and reduces them to
     field / variance / mask.  `distributed_cholesky=False` factors redundantly on every rank instead.
  4. one all_gather of the M-vectors.

A reference user would write `OrdinaryKriging(cov, idx, obs, error_cov)`; `krige_sharded` is that call for a
covariance that lives in row blocks (e.g. `EllipseCovarianceBuilder(..., rows=(r0, nrows)).cov_ns_device`).
"""

from __future__ import annotations

from warnings import warn

import numpy as np
import torch
import torch.distributed as dist

from . import _dense as dn
from .kriging import _finish_uncertainty


def assemble_obs_system(cov_rows: torch.Tensor, r0: int, idx: torch.Tensor, error_cov, group=None,
                        meanwhile=None) -> torch.Tensor:
    """S = C[idx, idx] + E on every rank from row-block shards of C.  `idx` must be sorted ascending (then the
    rows each rank contributes are contiguous in S).  `error_cov` (n x n, host numpy or tensor, identical on every
    rank, or None): each rank adds - and, for a host array, uploads - only the rows of E that belong to the
    observations it owns, so the host-to-device traffic of E is split across the ranks instead of repeated on each.
    Works on any device / backend (gloo test on CPU).

    The upload of this rank's rows of E runs on a side stream while the rank cuts its rows of C[idx, idx] out of its
    row block; the only exchange is ONE all-gather, straight into the buffer that becomes S (rows padded to the
    largest per-rank count, compacted in place when the counts differ).  `meanwhile` (optional callable) is run while
    the all-gather is in flight - `krige_sharded` builds its right-hand sides there."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = idx.numel()
    nrows = cov_rows.shape[0]
    mine = (idx >= r0) & (idx < r0 + nrows)
    own = torch.nonzero(mine).flatten()
    lo, hi = (int(own[0]), int(own[-1]) + 1) if own.numel() else (0, 0)
    contiguous = own.numel() == hi - lo                                      # always, for sorted idx and row blocks
    ldn = n + (n & 1)                                                        # even leading dimension (16-byte rows)
    counts = [hi - lo]
    if world > 1:
        ct = torch.zeros(world, dtype=torch.int64, device=cov_rows.device)
        ct[dist.get_rank(group)] = hi - lo
        dist.all_reduce(ct, group=group)
        counts = ct.cpu().tolist()
    cap = max(counts)
    rank = dist.get_rank(group) if world > 1 else 0
    # (no 3.2 GB memset at n = 20 000: every entry that is read later is written below - the pieces by their owners, the
    # rest by the all-gather - except the padding column of an odd n)
    buf = torch.empty((world * cap, ldn), dtype=cov_rows.dtype, device=cov_rows.device)
    if ldn != n:
        buf[:, n:] = 0
    piece = buf[rank * cap: rank * cap + (hi - lo), :n]
    if own.numel():
        piece.copy_(cov_rows[(idx[own] - r0).unsqueeze(1), idx.unsqueeze(0)])   # (n_mine, n), one fused gather (async)
    # ---- error-covariance rows of this rank: the upload (side stream; the host waits for it) overlaps the gather above --
    e_rows, side = None, None
    if error_cov is not None and own.numel():
        if isinstance(error_cov, torch.Tensor):
            e_rows = (error_cov[lo:hi] if contiguous else error_cov[own.to(error_cov.device)]).to(piece.device)
        elif piece.is_cuda:
            # host array: the rows of this rank's observations are one contiguous slice - a VIEW, packed into pinned
            # staging buffers by a thread pool while the previous buffer is in flight (`_dense._h2d_concat`); a
            # boolean-mask copy + pageable upload of the same 400 MB cost 0.13 s per rank at n = 20 000 on 8 GPUs
            E = np.asarray(error_cov)
            block = E[lo:hi] if contiguous else E[own.cpu().numpy()]
            slabs = [[block[r:r + 128]] for r in range(0, block.shape[0], 128)]      # one packing task per 128 rows
            side = _upload_stream(piece.device)
            with torch.cuda.stream(side):
                e_rows = dn._h2d_concat(slabs, int(block.size), piece.device).view(block.shape)
        else:
            E = np.asarray(error_cov, dtype=np.float64)
            e_rows = torch.from_numpy(np.ascontiguousarray(E[lo:hi] if contiguous else E[own.cpu().numpy()])).to(piece.device)
    if e_rows is not None:
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        piece += e_rows.to(piece.dtype)
    work = None
    if world > 1:
        work = dist.all_gather_into_tensor(buf, buf[rank * cap:(rank + 1) * cap].clone(), group=group, async_op=True)
    if meanwhile is not None:
        meanwhile()
    if work is not None:
        work.wait()
    if sum(counts) != n:
        raise ValueError("the row blocks of the ranks do not cover every observation index")
    if all(c == cap for c in counts):
        return buf[:n, :n]
    # compact the padded blocks in place, front to back (block r moves up by the padding of the blocks before it)
    pos = counts[0]
    for r in range(1, world):
        if counts[r]:
            buf[pos:pos + counts[r]] = buf[r * cap:r * cap + counts[r]].clone() if r * cap < pos + counts[r] else buf[r * cap:r * cap + counts[r]]
        pos += counts[r]
    return buf[:n, :n]                                                       # view: row stride ldn (even)


_UPLOAD_STREAMS: dict = {}


def _upload_stream(device) -> torch.cuda.Stream:
    key = (device.type, device.index)
    if key not in _UPLOAD_STREAMS:
        _UPLOAD_STREAMS[key] = torch.cuda.Stream(device)
    return _UPLOAD_STREAMS[key]


def local_rhs(cov_rows: torch.Tensor, idx: torch.Tensor, obs: torch.Tensor) -> torch.Tensor:
    """[C_cross[:, r0 : r0 + nrows] | 1 | y] from the rank's own rows: C_cross[a, r0 + j] = C[r0 + j, idx[a]]."""
    nrows, n = cov_rows.shape[0], idx.numel()
    T = torch.empty((n, nrows + 2 + (nrows & 1)), dtype=cov_rows.dtype, device=cov_rows.device)
    T[:, :nrows] = cov_rows[:, idx].t()
    T[:, nrows] = 1.0
    T[:, nrows + 1] = obs
    return T


def krige_sharded(cov_rows: torch.Tensor, r0: int, idx, obs, error_cov=None, method: str = "ordinary", mean=0.0,
                  group=None, distributed_cholesky: bool = True) -> dict:
    """Field, uncertainty and constraint mask on the FULL grid (every rank returns the same numpy arrays) for a
    covariance sharded by contiguous row blocks in rank order.  `idx` (sorted), `obs`, `error_cov` (n x n or None)
    are host arrays, identical on every rank."""
    if method not in ("ordinary", "simple"):
        raise ValueError(f"Unknown kriging method: {method}")
    import os
    import time

    if cov_rows.dtype != torch.float64:
        # the device kernels read double*; the builder's default precision is float32 (the reference's default)
        raise ValueError(f"krige_sharded needs a float64 covariance row block, got {cov_rows.dtype}: build it with "
                         "EllipseCovarianceBuilder(..., precision=np.float64)")
    device = cov_rows.device
    timing = bool(os.environ.get("GLOMAR_B200_TIMING")) and cov_rows.is_cuda
    marks = []

    def mark(name):
        if timing:
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    mark("start")
    idx_np = np.asarray(idx, dtype=np.int64)
    if np.any(np.diff(idx_np) <= 0):
        raise ValueError("idx must be sorted ascending and unique for the sharded path")
    idx_t = torch.from_numpy(idx_np).to(device)
    y = torch.from_numpy(np.asarray(obs, dtype=np.float64)).to(device)
    rhs = []
    S = assemble_obs_system(cov_rows, r0, idx_t, error_cov, group,       # rows of an even-stride buffer
                            meanwhile=lambda: rhs.append(local_rhs(cov_rows, idx_t, y)))
    T = rhs[0]
    mark("assemble S (gather + E upload + all_gather) with the local right-hand sides built meanwhile")
    nrows, N = cov_rows.shape
    # local column j has its prior variance at cov_rows[j, r0 + j]
    try:
        out = dn.solve_prepared(S, T, nrows, cov_rows.view(-1)[r0:], cov_rows.stride(0), y, method == "ordinary",
                                group=group if distributed_cholesky else False)
    except dn.NotPositiveDefinite as err:
        # The reference solves by LU whatever the matrix (np.linalg.solve, kriging.py:343, :585): an indefinite
        # C[idx, idx] + E - the default displacement method's covariance without eigenvalue_clip - goes through the
        # blocked LU with partial pivoting instead.  Every rank saw the same status (it is reduced over the group), so
        # all of them re-assemble the system (the factorisation overwrote it) and take this route together.
        warn(f"{err}; solving by LU with partial pivoting like the reference does - slow path")
        del S, T
        S = assemble_obs_system(cov_rows, r0, idx_t, error_cov, group)
        T = local_rhs(cov_rows, idx_t, y)
        out = dn.solve_prepared_lu(S, T, nrows, cov_rows.view(-1)[r0:], cov_rows.stride(0), y, method == "ordinary")
    mark("factor + forward solve + apply")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        sizes = torch.zeros(world, dtype=torch.int64, device=device)
        sizes[dist.get_rank(group)] = nrows
        dist.all_reduce(sizes, group=group)
        sizes = sizes.cpu().tolist()
        cap = max(sizes)
        buf = torch.zeros((5, cap), dtype=torch.float64, device=device)
        buf[:, :nrows] = out
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        out = torch.cat([p[:, :c] for p, c in zip(parts, sizes)], dim=1)
    o = out.cpu().numpy()
    mark("gather results")
    if timing and (not dist.is_initialized() or dist.get_rank(group) == 0):
        print("[krige_sharded] " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.1f} ms" for a, b in zip(marks, marks[1:])),
              flush=True)
    return {"field": o[dn.OUT_FIELD] + (mean if method == "simple" else 0.0),
            "uncertainty": _finish_uncertainty(o[dn.OUT_U2], np.float64),
            "constraint_mask": o[dn.OUT_MASK], "alpha": o[dn.OUT_ALPHA]}
