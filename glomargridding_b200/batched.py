"""Many kriging problems against one covariance matrix in one go (BASELINE.json config 2: 480 months of 5-degree
ordinary kriging, sharded by month across the GPUs of a box).

The reference has no batched entry point - a user loops `OrdinaryKriging(cov, idx_m, obs_m, E_m)` over months
(`glomar_gridding/kriging.py:494-812`, notebook cells 31-40).  `krige_batch` is that loop, run as ONE sequence of
batched launches per chunk of months (`_dense.KrigeBatch`: the month index is `blockIdx.z` of every kernel).
Each problem goes through exactly the constructor logic of the classes (error-covariance subsetting and the
NaN / zero-diagonal filter with its warning), so results are identical to the per-month classes.

Multi-GPU: problems are independent, so rank r of `world` takes problems r, r + world, ... (`shard`), with no
data-path collective; `gather_results` collects the M-vectors on every rank with one all_gather (gloo or nccl).
"""

from __future__ import annotations

import os

import numpy as np
import torch

from . import _dense as dn
from . import _device as _dev
from .kriging import OrdinaryKriging, SimpleKriging, _as_device_cov, _finish_uncertainty


def shard(n_items: int, rank: int, world: int) -> list[int]:
    """Round-robin ownership of independent problems (months, ensemble members)."""
    return list(range(rank, n_items, world))


def _bytes_per_problem(n: int, M: int) -> int:
    lin, pdb, _ = dn.solve_ws(n)
    slabs = (n + dn.SLAB - 1) // dn.SLAB
    return 8 * (n * (n + 2) + 2 * n * (M + 2) + lin + pdb + 4 * slabs * M + 5 * M + n * n)


_SIDE_STREAMS: dict = {}
_ARENAS: dict = {}


def release_workspaces() -> None:
    """Give the device arenas of the pipelined driver back to the allocator (they are kept between calls)."""
    _ARENAS.clear()


def _side_stream(device) -> torch.cuda.Stream:
    """One upload stream per device for the life of the process: the caching allocator keeps a pool per stream, so a
    fresh stream per call would cudaMalloc every buffer of every call anew (~50 ms on the first chunk)."""
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


def krige_batch(covariance, problems, method: str = "ordinary", mean=0.0, max_bytes: int | None = None,
                return_device: bool = False, symmetry: str = "auto"):
    """Kriged field, uncertainty and constraint mask for every (idx, obs, error_cov) in `problems`.

    Returns a dict of (len(problems), M) arrays: "field", "uncertainty", "constraint_mask" (numpy, or CUDA
    tensors "field", "variance", "constraint_mask" if `return_device`).  `max_bytes` bounds the device
    memory of one chunk of problems (default: 40 % of the free memory).

    Batches of 32 or more problems run as a three-stage pipeline over chunks (an eighth of the batch each, the first
    two smaller; `GLOMAR_B200_CHUNKS` overrides the eighth): constructor logic, packing + H2D on a side stream, device
    work.  The two device arenas the chunks live in, the pinned staging buffers and the side stream are kept between
    calls (a second call allocates nothing); `release_workspaces()` returns the arenas to the allocator.

    `symmetry` ("auto" | "exact" | "sample" | "assume", see `_dense.KrigeBatch`): how the error covariances are tested
    for symmetry.  A batch of 8 or more problems defaults to the SAMPLED test (96 mirrored entry pairs per matrix):
    an error covariance that is asymmetric in only a handful of entries is then treated as symmetric - its upper
    triangle is never read - whereas the reference would LU-solve it as given; pass "exact" when that matters."""
    if method not in ("ordinary", "simple"):
        raise ValueError(f"Unknown kriging method: {method}")
    cls = OrdinaryKriging if method == "ordinary" else SimpleKriging
    cov = _as_device_cov(covariance)
    M = cov.shape[0]
    # the classes' constructor logic (subset E, drop NaN / zero-diagonal observations, warn) per problem
    def prepare(problem):
        idx, obs, E = problem
        k = cls.__new__(cls)
        k.covariance, k.idx, k.obs, k.error_cov = covariance, np.asarray(idx), np.asarray(obs), E
        k._batch = k._user_weights = None
        k._nan_to_zero_on_device = True
        k.subset_error_covariance()
        return k._problem()

    if max_bytes is None:
        free, _ = torch.cuda.mem_get_info(cov.device)
        max_bytes = int(0.4 * free)
    # Chunks of problems: bounded by `max_bytes` of device memory (sized from the unfiltered observation counts, an
    # upper bound) and - for large batches - at most an eighth of the batch (but 24 problems or more), so that there is something to overlap:
    # while chunk i runs on the device, a background thread does the host work of chunk i+1 (the classes'
    # constructor logic - an O(n^2) NaN scan per error covariance, numpy releases the GIL for it -, the packing into
    # pinned staging buffers and the H2D copies, on a side stream).  End to end the three stages cost about the
    # same (profiles/: 480 months, 5 GB of error covariances), so the pipeline is worth ~2x.
    n_all = len(problems)
    if n_all == 0:
        empty = np.empty((0, M)) if not return_device else torch.empty((0, M), dtype=torch.float64, device=cov.device)
        return {"field": empty, ("variance" if return_device else "uncertainty"): empty, "constraint_mask": empty}
    n_chunks = max(1, int(os.environ.get("GLOMAR_B200_CHUNKS", "8")))
    cap = n_all if n_all < 32 else max(24, (n_all + n_chunks - 1) // n_chunks)
    bounds, start = [], 0
    while start < n_all:
        used, stop = 0, start
        # the first chunk has nothing to hide behind: a quarter-size one, then a half-size one, fill the pipeline
        cap_now = cap if n_all < 32 else max(8, cap >> max(0, 2 - len(bounds)))
        while stop < n_all and stop - start < cap_now:
            need = _bytes_per_problem(len(problems[stop][0]), M)
            if stop > start and used + need > max_bytes // 2:      # two chunks are alive at a time
                break
            used += need
            stop += 1
        bounds.append((start, stop))
        start = stop
    from concurrent.futures import ThreadPoolExecutor

    main_stream = torch.cuda.current_stream(cov.device)
    side = _side_stream(cov.device) if len(bounds) > 1 else main_stream
    workers = ThreadPoolExecutor(max_workers=_dev.host_threads(8)) if n_all >= 8 else None

    timing = os.environ.get("GLOMAR_B200_TIMING")
    import time as _time

    # Two arenas, used alternately by consecutive chunks (chunk i + 1 is uploaded while chunk i runs; chunk i is dead
    # before the upload of chunk i + 2 starts), kept for the life of the process like the pinned staging buffers.
    chunk_no = {lo: i for i, (lo, hi) in enumerate(bounds)}

    def arena_for(lo, prepared):
        if len(bounds) == 1:
            return None
        need = dn.krige_batch_doubles([len(p[0]) for p in prepared], M, any(p[2] is not None for p in prepared))
        key = (cov.device.index, chunk_no[lo] & 1)
        arena = _ARENAS.get(key)
        if arena is None or arena.buf.numel() < need:
            _ARENAS.pop(key, None)
            arena = None                                               # free the old one before asking for more
            _ARENAS[key] = arena = dn.Arena(int(1.15 * need), cov.device)
        return arena.reset()

    def prep(lo, hi):
        chunk = problems[lo:hi]
        return list(workers.map(prepare, chunk)) if workers is not None else [prepare(p) for p in chunk]

    def upload(prepared_or_future, lo, hi):
        torch.cuda.set_device(cov.device)
        t0 = _time.perf_counter()
        prepared = prepared_or_future.result() if hasattr(prepared_or_future, "result") else prepared_or_future
        t1 = _time.perf_counter()
        with torch.cuda.stream(side):
            kb = dn.KrigeBatch(cov, prepared, ordinary=method == "ordinary", alloc=arena_for(lo, prepared),
                               symmetry=symmetry if symmetry != "auto" or n_all < 8 else "sample")
            side.synchronize()                                     # uploads complete before the main stream reads them
        if timing:
            print(f"[krige_batch] upload {lo}:{hi} waited {1e3 * (t1 - t0):.1f} ms for prepare, pack + H2D "
                  f"{1e3 * (_time.perf_counter() - t1):.1f} ms", flush=True)
        return kb

    assert (dn.OUT_FIELD, dn.OUT_U2, dn.OUT_MASK) == (0, 1, 2)
    if return_device:
        out = torch.empty((n_all, 5, M), dtype=torch.float64, device=cov.device)
    else:
        # field, variance and mask planes of every chunk go home as soon as the chunk is done (pinned, asynchronous)
        host = torch.empty((3, n_all, M), dtype=torch.float64, pin_memory=True)

    def collect(batch, lo, hi):
        if return_device:
            out[lo:hi] = batch.out
        else:
            planes = batch.out[:, :3].permute(1, 0, 2).contiguous()
            for p in range(3):
                host[p, lo:hi].copy_(planes[p], non_blocking=True)

    try:
        if len(bounds) == 1:
            batch = upload(prep(*bounds[0]), *bounds[0]).run()
            collect(batch, 0, n_all)
            main_stream.synchronize()
            del batch
        else:
            # three stages in flight: constructor logic of chunk i+2.. (one thread feeding the worker pool), packing +
            # H2D of chunk i+1 (one thread, side stream), device work of chunk i (this thread, current stream)
            with ThreadPoolExecutor(max_workers=1) as prepper, ThreadPoolExecutor(max_workers=1) as uploader:
                prepped = [prepper.submit(prep, lo, hi) for lo, hi in bounds]
                fut = uploader.submit(upload, prepped[0], *bounds[0])
                for i, (lo, hi) in enumerate(bounds):
                    tw = _time.perf_counter()
                    batch = fut.result()
                    prepped[i] = None
                    tr = _time.perf_counter()
                    if i + 1 < len(bounds):
                        fut = uploader.submit(upload, prepped[i + 1], *bounds[i + 1])
                    batch.run()
                    collect(batch, lo, hi)
                    main_stream.synchronize()                      # its memory was allocated on the side stream
                    del batch
                    if timing:
                        print(f"[krige_batch] chunk {i}: waited {1e3 * (tr - tw):.1f} ms for its inputs, device "
                              f"{1e3 * (_time.perf_counter() - tr):.1f} ms", flush=True)
    finally:
        if workers is not None:
            workers.shutdown(wait=True)
    t_out = _time.perf_counter()
    if return_device:
        field = out[:, dn.OUT_FIELD] + (mean if method == "simple" else 0.0)
        return {"field": field, "variance": out[:, dn.OUT_U2], "constraint_mask": out[:, dn.OUT_MASK]}
    o = host.numpy()
    field = o[dn.OUT_FIELD] + mean if method == "simple" else o[dn.OUT_FIELD]
    unc = _finish_uncertainty(o[dn.OUT_U2], np.float64)            # elementwise: all problems at once
    if timing:
        print(f"[krige_batch] uncertainty clamp {1e3 * (_time.perf_counter() - t_out):.1f} ms", flush=True)
    return {"field": field, "uncertainty": unc, "constraint_mask": o[dn.OUT_MASK]}


def gather_results(local: np.ndarray, n_items: int, rank: int, world: int) -> np.ndarray:
    """Reassemble per-rank result rows (sharded by `shard`) into problem order on every rank."""
    if world == 1:
        return local
    import torch.distributed as dist

    M = local.shape[1]
    per = (n_items + world - 1) // world
    use_cuda = dist.get_backend() == "nccl"
    buf = torch.zeros((per, M), dtype=torch.float64, device="cuda" if use_cuda else "cpu")
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(buf.device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = np.empty((n_items, M), dtype=np.float64)
    for r in range(world):
        rows = shard(n_items, r, world)
        out[rows] = parts[r][: len(rows)].cpu().numpy()
    return out


def stochastic_ensemble(covariance, idx, obs, error_cov, n_members: int, seed: int | None = None,
                        return_device: bool = False):
    """`n_members` perturbed ordinary-kriging fields in one go (BASELINE.json config 5: 200 ensemble members).

    The reference produces one member per `StochasticKriging(...).solve()` call, each drawing a full state with an
    SVD of the M x M covariance (`stochastic.py:371-383`, 3.2 s per member in its notebook).  Here the three
    factorisations are done ONCE and every member is a column of a GEMM:

        L_C = chol(C), L_E = chol(E), L_S = chol(C[idx, idx] + E)              (DMMA Cholesky, K4)
        X = Z_1 L_C^T  (members x M)    noise = Z_2 L_E^T  (members x n)        (Philox normals + DMMA GEMM, K9)
        simulated_obs = X[:, idx] + noise
        epsilon = simulated_obs S^-1 C_cross - X                                (one solve with `members` columns + GEMM)
        result = gridded_field (ordinary kriging of the real obs) + epsilon     (`stochastic.py:386-392`)

    Returns {"result": (members, M), "gridded_field": (M,), "epsilon": (members, M), "uncertainty": (M,)} as
    numpy arrays (CUDA tensors if `return_device`).  The random stream is the device Philox generator keyed by
    `seed` (default: drawn from numpy's global state), not numpy's - see `stochastic.py` of this package.
    """
    from . import _device as dev
    from ._lib import check, lib
    from .stochastic import StochasticKriging, _chol_with_jitter

    import time as _time

    timing = bool(os.environ.get("GLOMAR_B200_TIMING"))
    marks = []

    def mark(name):
        if timing:
            torch.cuda.synchronize()
            marks.append((name, _time.perf_counter()))

    mark("start")
    sk = StochasticKriging(covariance, np.asarray(idx), np.asarray(obs), error_cov)   # E subsetting / filter / warning
    cov = sk._cov_device()
    M, n, nm = cov.shape[0], len(sk.idx), int(n_members)
    mark("constructor")
    if seed is None:
        seed = int(np.random.randint(0, 2**31 - 1)) | (int(np.random.randint(0, 2**31 - 1)) << 31)
    # The three factorisations - C (states), E (observation noise), S = C[idx, idx] + E (weights) - as ONE batch of three
    # systems through the overlapped pipeline: their block steps run in lockstep (21 latency-bound steps instead of
    # 21 + 11 + 11), the forward solve of S rides along, one status read-back at the end.  E goes up once (packed, with
    # the exact symmetry test) and is expanded to its lower triangle on the device for its own factorisation.
    L, s = lib(), dev.stream_ptr()
    batch = dn.KrigeBatch(cov, [sk._problem()], ordinary=True)
    if batch._asym:
        raise np.linalg.LinAlgError("stochastic_ensemble needs a symmetric error covariance")
    mark("upload (KrigeBatch)")
    C_ptr, ldc = cov.data_ptr(), cov.stride(0)
    check(L.gmg_krige_gather_batched_f64(C_ptr, M, ldc, batch._kd.data_ptr(), 1, n, s), "gather")
    L_C = dn.padded(cov[:, :M])
    L_E = torch.empty((n, n + (n & 1)), dtype=torch.float64, device=cov.device)
    check(L.gmg_krige_error_cov_lower_f64(batch._kd.data_ptr(), 0, n, L_E.data_ptr(), L_E.stride(0), s), "error_cov_lower")
    sd3 = np.zeros(3, dtype=dn.SOLVE_DESC)
    sd3[0] = batch._sd_np[0]
    info2 = torch.zeros(2, dtype=torch.int32, device=cov.device)
    keep = []
    for j, (mat, order) in enumerate(((L_C, M), (L_E, n)), start=1):
        lin, pdb, ldp = dn.solve_ws(order)
        ws_l = torch.empty(lin, dtype=torch.float64, device=cov.device)
        ws_p = torch.empty(pdb, dtype=torch.float64, device=cov.device)
        keep += [ws_l, ws_p]
        sd3[j] = (mat.data_ptr(), mat.stride(0), ws_l.data_ptr(), 0, 2, ws_p.data_ptr(), ldp,
                  info2.data_ptr() + 4 * (j - 1), order, 0)
    sd3_t = dn._desc_tensor(sd3, cov.device)
    check(L.gmg_potrf_solve_f64(None, sd3_t.data_ptr(), 3, M, batch.nrhs, None, 0, s), "gmg_potrf_solve_f64")
    check(L.gmg_krige_apply_batched_f64(C_ptr, M, ldc, batch._kd.data_ptr(), 1, n, 1, s), "apply")
    check(L.gmg_tril_f64(L_C.data_ptr(), M, L_C.stride(0), s), "gmg_tril_f64")
    check(L.gmg_tril_f64(L_E.data_ptr(), n, L_E.stride(0), s), "gmg_tril_f64")
    info = dn._d2h_info(torch.cat([batch.info, info2]))          # synchronises
    mark("gather + three factorisations + forward solve + apply")
    del keep
    if info[0] != 0:
        raise np.linalg.LinAlgError("stochastic_ensemble needs a positive definite C[idx, idx] + E")
    if info[1] != 0:       # the covariance itself is not positive definite: the jitter / |C| route of the class path
        L_C = _chol_with_jitter(cov, M)
    if info[2] != 0:
        E_low = torch.empty_like(L_E)
        check(L.gmg_krige_error_cov_lower_f64(batch._kd.data_ptr(), 0, n, E_low.data_ptr(), E_low.stride(0), s), "error_cov_lower")
        E_full = E_low[:, :n] + torch.tril(E_low[:, :n], -1).T
        L_E = _chol_with_jitter(dn.padded(E_full), n)
    sk._batch, sk._user_weights, sk._user_simple_weights = batch, None, None
    z = torch.empty(nm * (M + n), dtype=torch.float64, device=cov.device)
    check(lib().gmg_normal_f64(z.data_ptr(), z.numel(), seed, 0, dev.stream_ptr()), "gmg_normal_f64")
    Z1 = dn.padded(z[: nm * M].view(nm, M))
    Z2 = dn.padded(z[nm * M:].view(nm, n))
    X = dn.gemm(Z1[:, :M], L_C[:, :M], transb=True)            # (members, M) simulated states ~ N(0, C)
    noise = dn.gemm(Z2[:, :n], L_E[:, :n], transb=True)        # (members, n) ~ N(0, E)
    idx_t = dev.to_device(np.asarray(sk.idx), torch.int64)
    sim_obs = X[:, idx_t] + noise                              # plumbing gather + add of (members, n)
    V = batch.solve_vectors(0, sim_obs.t().contiguous())       # S^-1 simulated_obs^T : (n, members)
    sim_grid = dn.gemm(dn.padded(V), batch.Cx(0), transa=True)[:nm, :M]   # (members, M) = V^T C_cross
    eps = sim_grid - X
    field = batch.out[0, dn.OUT_FIELD]
    result = field.unsqueeze(0) + eps
    mark("normals + draws + epsilon")
    if timing:
        print("[stochastic_ensemble] " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f} ms" for a, b in zip(marks, marks[1:])),
              flush=True)
    if return_device:
        return {"result": result, "gridded_field": field, "epsilon": eps, "variance": batch.out[0, dn.OUT_U2]}
    return {"result": result.cpu().numpy(), "gridded_field": field.cpu().numpy(), "epsilon": eps.cpu().numpy(),
            "uncertainty": _finish_uncertainty(batch.out[0, dn.OUT_U2].cpu().numpy(), np.float64)}
