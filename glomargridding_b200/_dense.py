"""Host-side driver of the device kriging pipeline (gather -> Cholesky -> two triangular solves -> apply).

PyTorch owns the device buffers; every arithmetic step is a `libglomar_b200.so` entry point
(`include/glomar_b200.h`).  `KrigeBatch` runs any number of independent kriging problems that share one
covariance matrix - one problem for the `SimpleKriging` / `OrdinaryKriging` / `StochasticKriging` classes,
hundreds of months for the batched driver (`glomargridding_b200.batched`) - as ONE sequence of launches
(the batch index is `blockIdx.z`), not as per-problem streams of small kernels.
"""

from __future__ import annotations

import ctypes
import os
from warnings import warn

import numpy as np
import torch

from . import _device as dev
from ._lib import GlomarB200Error, check, lib

NB = 128      # factorisation block = GEMM tile of the library
SLAB = 256    # rows per partial sum of the apply pass

SOLVE_DESC = np.dtype([("S", "u8"), ("lds", "i8"), ("Linv", "u8"), ("R", "u8"), ("ldr", "i8"), ("P", "u8"),
                       ("ldp", "i8"), ("info", "u8"), ("n", "i4"), ("nrhs", "i4")])
KRIGE_DESC = np.dtype([("idx", "u8"), ("E", "u8"), ("lde", "i8"), ("y", "u8"), ("S", "u8"), ("lds", "i8"),
                       ("Cx", "u8"), ("ldx", "i8"), ("T", "u8"), ("ldt", "i8"), ("partial", "u8"), ("out", "u8"),
                       ("n", "i4"), ("ordinary", "i4")])
assert SOLVE_DESC.itemsize == 72 and KRIGE_DESC.itemsize == 104

OUT_FIELD, OUT_U2, OUT_MASK, OUT_ALPHA, OUT_ROWDOT = range(5)


def _even(x: int) -> int:
    return x + (x & 1)


def solve_ws(n: int) -> tuple[int, int, int]:
    """(doubles for the inverted diagonal blocks, doubles for the coefficient strip, its leading dimension)."""
    a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    check(lib().gmg_solve_ws_doubles(n, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "gmg_solve_ws_doubles")
    return a.value, b.value, c.value


def gesv_ws(n: int) -> tuple[int, int, int]:
    """Workspace of the blocked LU (`gmg_gesv_blocked_f64`) for systems of order up to n: (doubles for the inverted
    diagonal blocks of L and U, doubles for the coefficient strip + pivots, leading dimension of the strip)."""
    a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    check(lib().gmg_gesv_ws_doubles(n, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "gmg_gesv_ws_doubles")
    return a.value, b.value, c.value


def _desc_tensor(arr: np.ndarray, device) -> torch.Tensor:
    return _h2d(np.ascontiguousarray(arr).view(np.uint8).reshape(-1), device)


def _h2d(a: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


_INFO_HOST = None


def _d2h_info(t: torch.Tensor) -> np.ndarray:
    """Factorisation status words -> numpy through a persistent pinned buffer, waiting on the stream only: a pageable
    cudaMemcpy is synchronous inside the driver and stalls the CUDA calls of the other pipeline thread
    (`batched.krige_batch`) for as long as the factorisation runs."""
    global _INFO_HOST
    if _INFO_HOST is None or _INFO_HOST.numel() < t.numel():
        _INFO_HOST = torch.empty(max(4096, t.numel()), dtype=torch.int32, pin_memory=True)
    host = _INFO_HOST[: t.numel()]
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy().copy()


_STAGE = None


def _group_doubles(g, raw_f32: bool) -> int:
    """Doubles a group occupies in the blob: float32 arrays kept as float32 count 4 bytes per element."""
    return (sum(int(a.size) * (4 if raw_f32 and a.dtype == np.float32 else 8) for a in g) + 7) // 8


def _h2d_concat(groups, total: int, device, chunk_bytes: int = 128 << 20, out=None, post=None,
                raw_f32: bool = False) -> torch.Tensor:
    """Flattened concatenation of host floating-point arrays on the device, as float64.  `groups` is a list of
    lists of arrays (one list per kriging problem: the row blocks of its error covariance); every group is ONE task
    of a small thread pool that packs it into one of two pinned staging buffers (numpy copies release the GIL; a
    single thread packs at ~6 GB/s, far below PCIe) while the previous buffer is in flight, so packing overlaps the
    H2D copy - also inside one call, which is why the staging buffers are much smaller (128 MB) than a chunk of months
    (~500 MB).  The per-month error covariances are the bulk of the kriging path's host-to-device traffic.
    `post` (optional, one entry per group): a callable (or None) run by the packing task right after it has packed
    its group - the exact symmetry check of `KrigeBatch` rides here, on cache-warm data.
    `raw_f32`: float32 arrays are NOT widened - they are packed as float32, back to back inside their group, and a
    group is padded to a whole number of doubles (the device then reads them as float32, GMG_KRIGE_E_FLOAT32)."""
    from concurrent.futures import ThreadPoolExecutor

    if out is None:
        out = torch.empty(total, dtype=torch.float64, device=device)
    sizes = [_group_doubles(g, raw_f32) for g in groups]
    cap = max(chunk_bytes // 8, max(sizes))
    global _STAGE, _POOL
    if _STAGE is None or _STAGE[0].numel() < cap:       # pinned allocations are slow (~0.3 s): keep them
        _STAGE = [torch.empty(cap, dtype=torch.float64, pin_memory=True) for _ in range(2)]
    _pool()
    L = lib()
    stage = [t.numpy() for t in _STAGE]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    used = [False, False]
    # chunks of whole groups that fit one staging buffer
    chunks, cur, fill = [], [], 0
    for gi, (g, size) in enumerate(zip(groups, sizes)):
        if fill + size > cap and cur:
            chunks.append((cur, fill))
            cur, fill = [], 0
        cur.append((fill, g, post[gi] if post is not None else None))
        fill += size
    if cur:
        chunks.append((cur, fill))
    pos = 0
    for ci, (chunk, n) in enumerate(chunks):
        which = ci & 1
        if used[which]:
            done[which].synchronize()                    # the copy that last read this buffer has finished
        buf = stage[which]

        def pack(item, buf=buf):
            off, g, after = item
            at = 8 * off                                 # byte position in the staging buffer
            for a in g:                                  # one (strided) copy per block, any float dtype
                keep32 = raw_f32 and a.dtype == np.float32
                native = (a.ndim == 2 and a.dtype in (np.float64, np.float32) and a.strides[1] == a.itemsize
                          and a.strides[0] % a.itemsize == 0 and a.strides[0] >= a.shape[1] * a.itemsize)
                if keep32 and native:
                    L.gmg_host_pack_f32(a.ctypes.data, a.shape[0], a.shape[1], a.strides[0] // 4, buf.ctypes.data + at)
                elif keep32:
                    np.copyto(buf.view(np.float32)[at // 4:at // 4 + a.size].reshape(a.shape), a)
                elif native:
                    # native packer: streaming stores into the pinned buffer, float32 widened on the way
                    L.gmg_host_pack_f64(a.ctypes.data, int(a.dtype == np.float32), a.shape[0], a.shape[1],
                                        a.strides[0] // a.itemsize, buf.ctypes.data + at)
                else:
                    np.copyto(buf[at // 8:at // 8 + a.size].reshape(a.shape), a, casting="unsafe")
                at += a.size * (4 if keep32 else 8)
            if after is not None:
                after()

        list(_POOL.map(pack, chunk))
        out[pos:pos + n].copy_(_STAGE[which][:n], non_blocking=True)
        done[which].record()
        used[which] = True
        pos += n
    torch.cuda.current_stream().synchronize()
    return out


_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor

        _POOL = ThreadPoolExecutor(max_workers=dev.host_threads(14))   # memcpy-bound packing
    return _POOL


GMG_ERR_UNSUPPORTED = 4             # gmg_common.cuh
FUSED_RHS_GATHER = True             # tests switch it off to compare with the separate gather + solve calls
E_FLOAT32 = 2                       # GMG_KRIGE_E_FLOAT32 (include/glomar_b200.h): the descriptor's E holds float32 values
PACK_SYMMETRIC_ERROR_COV = True     # tests switch the trapezoid packing off to compare it with whole uploads
_SYM_PAIRS: dict[int, tuple[np.ndarray, np.ndarray]] = {}


def _looks_symmetric(E: np.ndarray, samples: int = 96) -> bool:
    """Cheap PRE-FILTER of the symmetry test: E[i, j] == E[j, i] on a fixed pseudo-random sample of index pairs.
    A matrix that fails it is uploaded whole straight away; one that passes is packed as its lower block trapezoid
    and verified EXACTLY by `_block_row_symmetric` inside the packing tasks (every mirrored pair of entries)."""
    n = E.shape[0]
    if n not in _SYM_PAIRS:
        rng = np.random.default_rng(n)
        _SYM_PAIRS[n] = (rng.integers(0, n, samples), rng.integers(0, n, samples))
    i, j = _SYM_PAIRS[n]
    return bool(np.array_equal(E[i, j], E[j, i], equal_nan=True))


def _block_row_symmetric(E: np.ndarray, r0: int, r1: int) -> bool:
    """Rows r0:r1 of E against columns r0:r1, up to and including the diagonal block - exact, entry by entry, NaN
    equal to NaN (both become 0 on the device).  Native (`gmg_host_rows_symmetric`, blocked, GIL released); called
    by the packing task that has just read E[r0:r1, :r1]."""
    if E.dtype in (np.float64, np.float32) and E.strides[1] == E.itemsize and E.strides[0] % E.itemsize == 0:
        rc = lib().gmg_host_rows_symmetric(E.ctypes.data, int(E.dtype == np.float32), E.shape[0],
                                           E.strides[0] // E.itemsize, r0, r1)
        if rc < 0:
            raise GlomarB200Error(f"gmg_host_rows_symmetric failed with status {rc}")
        return bool(rc)
    a, b = E[r0:r1, :r1], E[:r1, r0:r1].T                       # other dtypes / layouts: numpy
    return bool((a == b).all() or np.array_equal(a, b, equal_nan=True))


def is_symmetric(E: np.ndarray) -> bool:
    """Exact symmetry test of a host matrix (row blocks on the packing pool)."""
    n = E.shape[0]
    spans = [(r0, min(n, r0 + 4 * NB)) for r0 in range(0, n, 4 * NB)]
    if n * n < (1 << 22):
        return all(_block_row_symmetric(E, a, b) for a, b in spans)
    return all(_pool().map(lambda s: _block_row_symmetric(E, *s), spans))


def packed_error_cov_doubles(n: int) -> int:
    """Doubles uploaded for one symmetric n x n error covariance (block trapezoid, see `KrigeBatch`)."""
    if n <= NB:
        return n * n
    return sum((min(n, r0 + NB) - r0) * min(n, r0 + NB) for r0 in range(0, n, NB))


def factor_inplace(S: torch.Tensor, n: int, what: str = "matrix") -> dict:
    """S[:n, :n] (row-major, leading dimension S.stride(0), even) <- lower Cholesky factor L (the strict upper
    triangle keeps its old contents).  Returns the solve workspace (inverted diagonal blocks etc.).

    Raises numpy.linalg.LinAlgError if a pivot is not positive."""
    assert S.is_cuda and S.dtype == torch.float64 and S.stride(1) == 1 and S.stride(0) % 2 == 0
    lin, pdb, ldp = solve_ws(n)
    ws = {"Linv": torch.empty(lin, dtype=torch.float64, device=S.device),
          "P": torch.empty(pdb, dtype=torch.float64, device=S.device),
          "info": torch.zeros(1, dtype=torch.int32, device=S.device), "ldp": ldp}
    d = np.zeros(1, dtype=SOLVE_DESC)
    d["S"], d["lds"], d["Linv"], d["P"], d["ldp"], d["info"], d["n"] = (
        S.data_ptr(), S.stride(0), ws["Linv"].data_ptr(), ws["P"].data_ptr(), ldp, ws["info"].data_ptr(), n)
    dt = _desc_tensor(d, S.device)
    # look-ahead factorisation (gmg_potrf_solve_f64 with no right-hand sides): same tiles, same arithmetic as
    # gmg_potrf_batched_f64, the diagonal-block / panel chain overlapped with the bulk of the trailing update
    check(lib().gmg_potrf_solve_f64(None, dt.data_ptr(), 1, n, 0, None, 0, dev.stream_ptr()), "gmg_potrf_solve_f64")
    code = int(ws["info"].item())
    if code != 0:
        raise NotPositiveDefinite(
            f"{what} is not positive definite (Cholesky pivot {code} of {n} is not positive); "
            "apply glomar_gridding.covariance_tools.eigenvalue_clip or use delta_x_method='Met_Office'")
    return ws


def cholesky_solve_inplace(S: torch.Tensor, R: torch.Tensor, what: str = "matrix", lu_fallback: bool = True) -> None:
    """R <- S^-1 R for one symmetric S (n x n, overwritten by its factors): Cholesky on DMMA tiles, or - if S is
    not positive definite and `lu_fallback` - LU with partial pivoting like the reference's np.linalg.solve."""
    n = S.shape[0]
    assert R.is_cuda and R.dtype == torch.float64 and R.stride(1) == 1 and R.stride(0) % 2 == 0
    keep = S.clone() if lu_fallback else None
    d = np.zeros(1, dtype=SOLVE_DESC)
    d["R"], d["ldr"], d["n"], d["nrhs"] = R.data_ptr(), R.stride(0), n, R.shape[1]
    try:
        ws = factor_inplace(S, n, what)
    except np.linalg.LinAlgError:
        if not lu_fallback:
            raise
        warn(f"{what} is not positive definite; solving by LU with partial pivoting")
        S.copy_(keep)
        lin, pdb, ldp = gesv_ws(n)
        linv = torch.empty(lin, dtype=torch.float64, device=S.device)
        strip = torch.empty(pdb, dtype=torch.float64, device=S.device)
        info = torch.zeros(1, dtype=torch.int32, device=S.device)
        d["S"], d["lds"], d["Linv"], d["P"], d["ldp"], d["info"] = (
            S.data_ptr(), S.stride(0), linv.data_ptr(), strip.data_ptr(), ldp, info.data_ptr())
        dt = _desc_tensor(d, S.device)
        check(lib().gmg_gesv_blocked_f64(dt.data_ptr(), 1, n, R.shape[1], dev.stream_ptr()), "gmg_gesv_blocked_f64")
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError("Singular matrix")
        return
    d["S"], d["lds"], d["Linv"], d["P"], d["ldp"], d["info"] = (
        S.data_ptr(), S.stride(0), ws["Linv"].data_ptr(), ws["P"].data_ptr(), ws["ldp"], ws["info"].data_ptr())
    dt = _desc_tensor(d, S.device)
    check(lib().gmg_potrs_batched_f64(dt.data_ptr(), 1, n, R.shape[1], 3, dev.stream_ptr()), "gmg_potrs_batched_f64")
    torch.cuda.current_stream().synchronize()   # the descriptor / workspace tensors die with this frame


def gemm(A: torch.Tensor, B: torch.Tensor, transa: bool = False, transb: bool = False,
         out: torch.Tensor | None = None, alpha: float = 1.0, beta: float = 0.0) -> torch.Tensor:
    """alpha op(A) op(B) + beta out on DMMA tiles (row-major, contiguous, even leading dimensions)."""
    m, k = (A.shape[1], A.shape[0]) if transa else A.shape
    k2, n = (B.shape[1], B.shape[0]) if transb else B.shape
    assert k == k2, (A.shape, B.shape, transa, transb)
    assert A.stride(1) == 1 and B.stride(1) == 1 and A.stride(0) % 2 == 0 and B.stride(0) % 2 == 0
    if out is None:
        out = torch.empty((m, n), dtype=torch.float64, device=A.device)
    check(lib().gmg_gemm_f64(int(transa), int(transb), m, n, k, alpha, A.data_ptr(), A.stride(0), B.data_ptr(),
                             B.stride(0), beta, out.data_ptr(), out.stride(0), dev.stream_ptr()), "gmg_gemm_f64")
    return out


def apply_reduce(cov: torch.Tensor, Cx: torch.Tensor, T: torch.Tensor, y: torch.Tensor, ordinary: bool) -> torch.Tensor:
    """One apply pass (`gmg_krige_apply_batched_f64`) for given device operands.

    cov (M, M): only its diagonal is read; Cx (n, >=M) = C_cross; T (n, >=M [+1]) transposed weights
    (column M = S^-1 1 if `ordinary`); y (n,).  Returns the (5, M) result rows (OUT_* indices)."""
    M, n = cov.shape[0], T.shape[0]
    assert Cx.stride(1) == 1 and T.stride(1) == 1 and y.is_contiguous()
    for name, t in (("cov", cov), ("Cx", Cx), ("T", T), ("y", y)):          # the kernels read double*
        if t.dtype != torch.float64:
            raise TypeError(f"apply_reduce: {name} must be float64, got {t.dtype}")
    slabs = (n + SLAB - 1) // SLAB
    partial = torch.empty(4 * slabs * M, dtype=torch.float64, device=cov.device)
    out = torch.empty((5, M), dtype=torch.float64, device=cov.device)
    kd = np.zeros(1, dtype=KRIGE_DESC)
    kd[0] = (0, 0, n, y.data_ptr(), 0, 0, Cx.data_ptr(), Cx.stride(0), T.data_ptr(), T.stride(0),
             partial.data_ptr(), out.data_ptr(), n, int(ordinary))
    kdt = _desc_tensor(kd, cov.device)
    check(lib().gmg_krige_apply_batched_f64(cov.data_ptr(), M, cov.stride(0), kdt.data_ptr(), 1, n, 0,
                                            dev.stream_ptr()), "gmg_krige_apply_batched_f64")
    torch.cuda.current_stream().synchronize()   # descriptor / workspace tensors die with this frame
    return out


def padded(a, rows: int | None = None, cols: int | None = None) -> torch.Tensor:
    """Host or device matrix -> float64 CUDA tensor with an even leading dimension (zero padded)."""
    t = dev.to_device(a, torch.float64) if not isinstance(a, torch.Tensor) else a.to(dev.require_cuda(), torch.float64)
    r, c = t.shape
    out = torch.zeros((rows or r, _even(cols or c)), dtype=torch.float64, device=t.device)
    out[:r, :c] = t
    return out


class Arena:
    """Bump allocator over one float64 device buffer: `KrigeBatch` carves its pools (S, T, inverted blocks, error
    covariances ...) out of it instead of asking the caching allocator for five multi-hundred-MB blocks per chunk of
    months (5-6 ms of cudaMalloc per chunk in the pipelined driver, `batched.krige_batch`).  256-byte granularity."""

    def __init__(self, numel: int, device):
        self.buf = torch.empty(max(int(numel), 32), dtype=torch.float64, device=device)
        self.off = 0

    def reset(self) -> "Arena":
        self.off = 0
        return self

    def __call__(self, numel: int) -> torch.Tensor:
        numel = int(numel)
        if self.off + numel > self.buf.numel():
            return torch.empty(numel, dtype=torch.float64, device=self.buf.device)    # does not fit: ordinary allocation
        out = self.buf[self.off:self.off + numel]
        self.off += (numel + 31) // 32 * 32
        return out


def krige_batch_doubles(ns, M: int, has_E=True) -> int:
    """Upper bound of the doubles a `KrigeBatch` of problems with `ns` observations takes from its allocator."""
    total = 0
    for n in ns:
        lin, pdb, _ = solve_ws(n)
        slabs = (n + SLAB - 1) // SLAB
        total += n * _even(n) + n * _even(M + 2) + lin + pdb + 4 * slabs * M + (n * n if has_E else 0)
    return total + sum(2 * n + 5 * M for n in ns) + 32 * len(ns) + 32 * 8     # + idx, y, results, descriptors, alignment


class KrigeBatch:
    """`B` kriging problems against one covariance matrix, solved together on the current device.

    problems: sequence of (idx int64[n], obs float64[n], error_cov float64[n, n] | None), already filtered
    by `Kriging.subset_error_covariance`.  After `run()`:
      out[b]  (5, M) tensor rows: field, variance before clamping, constraint mask, alpha, rowdot(W, C_cross^T)
    `run()` stops after the FORWARD triangular solve: with S^-1 = L^-T L^-1 every result row is a column dot
    product of Y = L^-1 [C_cross | 1 | y], so the second solve - half of the flops - is only done by
    `ensure_full()` (when the weights matrix itself is wanted):
      T(b)    (n, ldt) tensor: S^-1 [C_cross | 1 | y]: simple-kriging weights transposed, column M = S^-1 1
    """

    def __init__(self, cov: torch.Tensor, problems, ordinary: bool, alloc=None, symmetry: str = "auto"):
        """`symmetry` - how it is decided whether an error covariance is symmetric (a symmetric one travels as its lower
        block trapezoid and is factored by Cholesky; an asymmetric one is uploaded whole and solved by LU on the matrix
        as given, like the reference, kriging.py:342-343, :576-585):
          "exact"   every mirrored pair of entries is compared, inside the packing tasks (one more pass over the host
                    matrix: ~0.1 s for one 20 000 x 20 000 system, ~0.2 ms per 5-degree month on 14 threads);
          "sample"  96 pseudo-random mirrored pairs per matrix: catches any asymmetry that is not confined to a few
                    entries; a batch of hundreds of months stays bound by its upload, not by host memory bandwidth;
          "assume"  the caller vouches for symmetry: no test, the upper triangles are never read;
          "auto"    (default) "exact" for fewer than 8 problems - the class calls -, "sample" for larger batches."""
        if symmetry not in ("auto", "exact", "sample", "assume"):
            raise ValueError(f"Unknown symmetry policy: {symmetry}")
        assert cov.is_cuda and cov.dtype == torch.float64 and cov.dim() == 2 and cov.stride(1) == 1
        import time

        marks = [("start", time.perf_counter())]
        self.cov, self.ordinary = cov, bool(ordinary)
        self.M = M = cov.shape[0]
        device = cov.device
        B = self.B = len(problems)
        ns = [int(len(p[0])) for p in problems]
        if min(ns) <= 0:
            raise ValueError("a kriging problem needs at least one observation")
        self.ns, self.n_max = ns, max(ns)
        self.ldx, self.ldt = _even(M), _even(M + 2)
        self.nrhs = M + 2
        lds = [_even(n) for n in ns]
        ws = [solve_ws(n) for n in ns]
        slabs = [(n + SLAB - 1) // SLAB for n in ns]

        if alloc is None:
            def alloc(numel):
                return torch.empty(int(numel), dtype=torch.float64, device=device)

        def pool(sizes):
            offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
            return alloc(int(offs[-1])), offs

        self._S, oS = pool([n * l for n, l in zip(ns, lds)])
        self._T, oT = pool([n * self.ldt for n in ns])
        self._Linv, oL = pool([w[0] for w in ws])
        self._P, oP = pool([w[1] for w in ws])
        self._partial, oPa = pool([4 * s * M for s in slabs])
        self.out = alloc(B * 5 * M).view(B, 5, M)
        marks.append(("pools", time.perf_counter()))
        offs_n = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
        # observation cells index C like numpy indexes an array (`self.covariance[self.idx, :]`, kriging.py:337, :571):
        # negative values count from the end, anything outside [-M, M) is an IndexError - checked here, on the host,
        # because the gather kernel dereferences C + idx * ldc unchecked
        idx_all = np.concatenate([np.asarray(p[0]).astype(np.int64, copy=False).reshape(-1) for p in problems])
        lo, hi = int(idx_all.min()), int(idx_all.max())
        if lo < -M or hi >= M:
            bad = lo if lo < -M else hi
            raise IndexError(f"index {bad} is out of bounds for axis 0 with size {M}")
        if lo < 0:
            idx_all = np.where(idx_all < 0, idx_all + M, idx_all)
        has_E = [p[2] is not None for p in problems]
        Es = []
        for p, n, h in zip(problems, ns, has_E):
            E = np.asarray(p[2]) if h else None
            if h and E.shape != (n, n):
                raise ValueError(f"error covariance of shape {E.shape} does not match {n} observations")
            Es.append(E.astype(np.float64) if h and E.dtype.kind != "f" else E)
        # A symmetric error covariance (what a covariance is; checked on a fixed sample of mirrored entry pairs, an
        # asymmetric one is uploaded whole) travels as the block trapezoid below and on its 128 x 128 diagonal
        # blocks - row block b of 128 rows with its first 128 (b + 1) columns, a strided VIEW of the host array.
        # That is everything the Cholesky path reads and halves the H2D bytes of the kriging path
        # (GmgKrigeDesc.lde < 0; the LU fallback reads the mirrored entries).
        # float32 error covariances (HadCRUT's are) stay float32 all the way to the gather kernel, which widens them -
        # exactly what numpy's `C[idx][:, idx] + E` does with a float32 E - when every matrix of the batch is float32
        f32 = any(has_E) and all(E is None or E.dtype == np.float32 for E in Es)
        self.error_cov_dtype = np.float32 if f32 else np.float64
        groups, posts, sizes_E, ldes, self._asym = [], [], [], [], []
        late_asym: set[int] = set()                              # found asymmetric by the exact test while packing
        task_cap = (4 << 20) if B >= 8 else (1 << 17)          # doubles per packing task: a lone problem still uses the pool
        for b, (E, n) in enumerate(zip(Es, ns)):
            size, lde = 0, n
            if E is not None:
                sym = True if symmetry == "assume" else _looks_symmetric(E)
                if not sym:
                    self._asym.append(b)                 # no Cholesky for this system: LU like the reference
                exact = sym and (symmetry == "exact" or (symmetry == "auto" and B < 8))
                rows = [(r0, min(n, r0 + NB)) for r0 in range(0, n, NB)]
                if sym and n > NB and PACK_SYMMETRIC_ERROR_COV:
                    lde = -NB
                    blocks = [E[r0:r1, :r1] for r0, r1 in rows]
                else:
                    blocks = [E[r0:r1] for r0, r1 in rows]                   # whole matrix, in row slabs
                # packing tasks of at most ~32 MB (1 MB when there are few problems): one per 5-degree month of a batch,
                # many for one large system (whose 1.6 - 3.2 GB would otherwise be ONE task and need staging buffers of
                # that size).  The task that packs rows r0:r1 then tests them against columns r0:r1 - the EXACT symmetry
                # test, on the pool threads and on data the task has just pulled through the cache.
                def checker(spans, E=E, b=b):
                    def run():
                        if b not in late_asym and not all(_block_row_symmetric(E, r0, r1) for r0, r1 in spans):
                            late_asym.add(b)
                    return run

                # (float32 mode: every block but a matrix's last has an even number of entries - 128 rows - so tasks start
                # on 8-byte boundaries and only the end of a matrix is padded)
                task, spans, fill = [], [], 0
                for blk, span in zip(blocks, rows):
                    if task and fill + blk.size > task_cap:
                        groups.append(task)
                        posts.append(checker(spans) if exact else None)
                        task, spans, fill = [], [], 0
                    task.append(blk)
                    spans.append(span)
                    fill += blk.size
                groups.append(task)
                posts.append(checker(spans) if exact else None)
                size = sum(blk.size for blk in blocks)
                if f32:
                    size = (size + 1) // 2
            sizes_E.append(size)
            ldes.append(lde)
        offs_E = np.concatenate([[0], np.cumsum(sizes_E)]).astype(np.int64)
        marks.append(("symmetry + blocks", time.perf_counter()))
        # ALL inputs - error covariances, observation cells, observations and the two descriptor arrays - go up as one
        # blob through the pinned, pipelined staging of `_h2d_concat`: [E | idx | y | krige descs | solve descs].
        # (Separate small pageable copies stall against the CUDA calls of the other pipeline thread, ~5 ms per chunk.)
        nE, nn = int(offs_E[-1]), int(offs_n[-1])
        o_idx, o_y, o_kd = nE, nE + nn, nE + 2 * nn
        o_sd = o_kd + B * KRIGE_DESC.itemsize // 8
        o_info = o_sd + B * SOLVE_DESC.itemsize // 8
        blob = alloc(o_info + (B + 1) // 2)
        self.info = blob[o_info:].view(torch.int32)[:B]       # zeroed by the upload
        self._E = blob[:nE] if nE else None
        self._idx = blob[o_idx:o_y].view(torch.int64)
        self._y = blob[o_y:o_kd]
        self._kd, self._sd = blob[o_kd:o_sd], blob[o_sd:o_info]
        self.h2d_bytes = 8 * blob.numel()

        kd = np.zeros(B, dtype=KRIGE_DESC)
        sd = np.zeros(B, dtype=SOLVE_DESC)
        f8 = 8
        for b in range(B):
            n = ns[b]
            S_ptr = self._S.data_ptr() + f8 * int(oS[b])
            T_ptr = self._T.data_ptr() + f8 * int(oT[b])
            kd[b] = (self._idx.data_ptr() + f8 * int(offs_n[b]),
                     (blob.data_ptr() + f8 * int(offs_E[b])) if has_E[b] else 0, ldes[b],
                     self._y.data_ptr() + f8 * int(offs_n[b]), S_ptr, lds[b], 0, self.ldx, T_ptr, self.ldt,
                     self._partial.data_ptr() + f8 * int(oPa[b]), self.out.data_ptr() + f8 * 5 * M * b, n,
                     int(self.ordinary) | (E_FLOAT32 if f32 and has_E[b] else 0))
            sd[b] = (S_ptr, lds[b], self._Linv.data_ptr() + f8 * int(oL[b]), T_ptr, self.ldt,
                     self._P.data_ptr() + f8 * int(oP[b]), ws[b][2], self.info.data_ptr() + 4 * b, n, self.nrhs)
        marks.append(("descriptors", time.perf_counter()))
        y_all = np.concatenate([np.asarray(p[1], dtype=np.float64).reshape(-1) for p in problems])
        groups += [[idx_all.view(np.float64)], [y_all], [kd.view(np.uint8).view(np.float64)],
                   [sd.view(np.uint8).view(np.float64)], [np.zeros((B + 1) // 2)]]
        posts += [None] * 5
        _h2d_concat(groups, blob.numel(), device, out=blob, post=posts, raw_f32=f32)
        marks.append(("pack + H2D of all inputs", time.perf_counter()))
        self._E_whole = []
        if late_asym:
            # The exact test found mirrored entries that differ in matrices the sample had passed: these systems are
            # solved by LU on the matrix as given, like the reference (kriging.py:342-343, :576-585) - their whole
            # matrix goes up now and their descriptors are patched.  Rare by construction (a covariance is symmetric).
            for b in sorted(late_asym):
                self._asym.append(b)
                whole = _h2d(np.ascontiguousarray(Es[b], dtype=np.float64), device)
                self._E_whole.append(whole)
                kd[b]["E"], kd[b]["lde"], kd[b]["ordinary"] = whole.data_ptr(), ns[b], int(self.ordinary)
            self._asym.sort()
            self._kd.copy_(_desc_tensor(kd, device).view(torch.float64))
            torch.cuda.current_stream().synchronize()
            marks.append(("re-upload of asymmetric error covariances", time.perf_counter()))
        if os.environ.get("GLOMAR_B200_TIMING") and B > 1:
            print(f"[KrigeBatch] {B} problems: " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.1f} ms"
                                                              for a, b in zip(marks, marks[1:])), flush=True)
        self._kd_np, self._sd_np = kd, sd
        self._oS, self._oT, self._lds, self._offs_n = oS, oT, lds, offs_n
        self._Cx = {}                       # C_cross copies, made on demand (LU fallback, weights from outside)
        self.full = np.zeros(B, dtype=bool)  # True once T(b) holds S^-1 [Cx | 1 | y] (both solves done)
        self.lu_fallback: list[int] = []

    # -- views -----------------------------------------------------------------------------------------
    def T(self, b: int) -> torch.Tensor:
        """S^-1 [C_cross | 1 | y] of problem b (runs the backward solve if it has not been done)."""
        self.ensure_full([b])
        return self._T_view(b)

    def _T_view(self, b: int) -> torch.Tensor:
        return self._T[int(self._oT[b]):int(self._oT[b + 1])].view(self.ns[b], self.ldt)

    def Cx(self, b: int) -> torch.Tensor:
        """C[idx, :] of problem b, (n, ldx); gathered on first use (plumbing copy)."""
        if b not in self._Cx:
            idx = self._idx[int(self._offs_n[b]):int(self._offs_n[b + 1])]
            cx = torch.zeros((self.ns[b], self.ldx), dtype=torch.float64, device=self.cov.device)
            cx[:, : self.M] = self.cov[idx, : self.M]
            self._Cx[b] = cx
        return self._Cx[b]

    def S(self, b: int) -> torch.Tensor:
        return self._S[int(self._oS[b]):int(self._oS[b + 1])].view(self.ns[b], self._lds[b])

    # -- pipeline --------------------------------------------------------------------------------------
    def run(self, allow_lu_fallback: bool = True) -> "KrigeBatch":
        """gather -> Cholesky -> forward triangular solve -> apply.  Systems whose Cholesky meets a non-positive
        pivot are re-gathered and solved by LU with partial pivoting (what the reference does for every system),
        with a warning; an exactly singular system raises numpy.linalg.LinAlgError like numpy does."""
        L, s = lib(), dev.stream_ptr()
        C, M, ldc = self.cov.data_ptr(), self.M, self.cov.stride(0)
        self.full[:] = False
        # Cholesky and the forward triangular solve as ONE overlapped pipeline (the solve step of block row k starts as
        # soon as diagonal block k exists, the serial diagonal-block kernel of the factorisation hides behind the GEMMs
        # of the solve), then the apply pass - all enqueued before the factorisation status is read, so the common case
        # (every system positive definite) has no host synchronisation inside it.  A system whose Cholesky failed
        # produced garbage that the LU pass below overwrites.
        # gmg_krige_solve_f64 also folds the gather of the right-hand sides C[idx, :] into the solve steps (T is never
        # written by a gather kernel: 23 MB less to write and to read back per 5-degree month); a covariance with an odd
        # order or leading dimension takes the two separate calls - same bits either way.
        rc = GMG_ERR_UNSUPPORTED
        if FUSED_RHS_GATHER:
            rc = L.gmg_krige_solve_f64(C, M, ldc, self._kd.data_ptr(), self._sd.data_ptr(), self.B, self.n_max, s)
        if rc == GMG_ERR_UNSUPPORTED:
            check(L.gmg_krige_gather_batched_f64(C, M, ldc, self._kd.data_ptr(), self.B, self.n_max, s), "gather")
            check(L.gmg_potrf_solve_f64(None, self._sd.data_ptr(), self.B, self.n_max, self.nrhs, None, 0, s),
                  "gmg_potrf_solve_f64")
        else:
            check(rc, "gmg_krige_solve_f64")
        check(L.gmg_krige_apply_batched_f64(C, M, ldc, self._kd.data_ptr(), self.B, self.n_max, 1, s), "apply")
        info = _d2h_info(self.info)                  # synchronises: the factorisation outcome decides the path
        asym = np.asarray(self._asym, dtype=np.int64)
        if asym.size:
            warn(f"the error covariance of {asym.size} of {self.B} kriging system(s) is not symmetric (first: system "
                 f"{int(asym[0])}); solving those by LU with partial pivoting like the reference does - slow path")
        not_spd = np.setdiff1d(np.flatnonzero(info), asym)
        bad = np.union1d(not_spd, asym)
        self.lu_fallback = bad.tolist()
        if bad.size == 0:
            return self
        if not allow_lu_fallback and not_spd.size:
            self.raise_if_not_spd(info)
        if not_spd.size:
            warn(f"C[idx, idx] + E is not positive definite for {not_spd.size} of {self.B} kriging system(s) "
                 f"(first: system {int(not_spd[0])}, Cholesky pivot {int(info[not_spd[0]])}); solving those by LU "
                 "with partial pivoting like the reference does - slow path, consider "
                 "glomar_gridding.covariance_tools.eigenvalue_clip or delta_x_method='Met_Office'")
        good = np.setdiff1d(np.arange(self.B), bad)
        self.info.zero_()
        kd_bad_np = self._kd_np[bad].copy()
        for i, b in enumerate(bad):
            kd_bad_np[i]["Cx"] = self.Cx(int(b)).data_ptr()
        kd_bad = _desc_tensor(kd_bad_np, self.cov.device)
        n_bad = max(self.ns[int(b)] for b in bad)
        # blocked LU with partial pivoting (gmg_gesv_blocked_f64): its workspaces - inverted diagonal blocks of L and U,
        # coefficient strip + pivots - are sized for the largest of these systems and allocated here, on the rare path
        lin, pdb, ldp = gesv_ws(n_bad)
        lu_linv = torch.empty(int(bad.size) * lin, dtype=torch.float64, device=self.cov.device)
        lu_strip = torch.empty(int(bad.size) * pdb, dtype=torch.float64, device=self.cov.device)
        sd_bad_np = self._sd_np[bad].copy()
        for i in range(int(bad.size)):
            sd_bad_np[i]["Linv"] = lu_linv.data_ptr() + 8 * i * lin
            sd_bad_np[i]["P"] = lu_strip.data_ptr() + 8 * i * pdb
            sd_bad_np[i]["ldp"] = ldp
        sd_bad = _desc_tensor(sd_bad_np, self.cov.device)
        check(L.gmg_krige_gather_batched_f64(C, M, ldc, kd_bad.data_ptr(), int(bad.size), n_bad, s), "gather")
        check(L.gmg_gesv_blocked_f64(sd_bad.data_ptr(), int(bad.size), n_bad, self.nrhs, s), "gmg_gesv_blocked_f64")
        check(L.gmg_krige_apply_batched_f64(C, M, ldc, kd_bad.data_ptr(), int(bad.size), n_bad, 0, s), "apply")
        self.full[bad] = True                        # (the good systems were completed by the overlapped pass above)
        if np.any(_d2h_info(self.info)):             # also synchronises before the temporaries die
            raise np.linalg.LinAlgError("Singular matrix")
        return self

    def ensure_full(self, which=None) -> None:
        """Second triangular solve (T <- L^-T T) for the listed problems (default: all) that still need it."""
        todo = np.array([b for b in (range(self.B) if which is None else which) if not self.full[b]], dtype=np.int64)
        if todo.size == 0:
            return
        sd = _desc_tensor(self._sd_np[todo], self.cov.device)
        n_max = max(self.ns[int(b)] for b in todo)
        check(lib().gmg_potrs_batched_f64(sd.data_ptr(), int(todo.size), n_max, self.nrhs, 2, dev.stream_ptr()),
              "gmg_potrs_batched_f64")
        torch.cuda.current_stream().synchronize()
        self.full[todo] = True

    def solve_vectors(self, b: int, V: torch.Tensor) -> torch.Tensor:
        """S_b^-1 V for extra right-hand sides V (n, k) with the factorisation of problem b (Cholesky path)."""
        if b in self.lu_fallback:
            raise GlomarB200Error("extra solves are not available for a system that took the LU fallback")
        R = padded(V)
        d = self._sd_np[[b]].copy()
        d["R"], d["ldr"], d["nrhs"] = R.data_ptr(), R.stride(0), V.shape[1]
        dt = _desc_tensor(d, self.cov.device)
        check(lib().gmg_potrs_batched_f64(dt.data_ptr(), 1, self.ns[b], V.shape[1], 3, dev.stream_ptr()),
              "gmg_potrs_batched_f64")
        torch.cuda.current_stream().synchronize()
        return R[:, : V.shape[1]]

    def raise_if_not_spd(self, info=None) -> None:
        info = self.info.cpu().numpy() if info is None else info
        bad = np.flatnonzero(info)
        if bad.size:
            b = int(bad[0])
            raise np.linalg.LinAlgError(
                f"observation covariance C[idx, idx] + E of problem {b} is not positive definite "
                f"(Cholesky pivot {int(info[b])} of {self.ns[b]} is not positive); apply "
                "glomar_gridding.covariance_tools.eigenvalue_clip to the covariance or use "
                "delta_x_method='Met_Office'")

    def weights(self, b: int, ordinary: bool | None = None) -> torch.Tensor:
        """The reference's `kriging_weights` for problem b: (M, n) simple or (M, n + 1) ordinary."""
        ordinary = self.ordinary if ordinary is None else ordinary
        if ordinary and not self.ordinary:
            raise ValueError("ordinary weights need an ordinary-kriging batch")
        n, M = self.ns[b], self.M
        T = self.T(b)
        W = torch.empty((M, n + (1 if ordinary else 0)), dtype=torch.float64, device=T.device)
        o_ptr = T.data_ptr() + 8 * M if ordinary else None
        a_ptr = self.out[b, OUT_ALPHA].data_ptr() if ordinary else None
        check(lib().gmg_krige_weights_f64(T.data_ptr(), self.ldt, o_ptr, self.ldt, a_ptr, n, M, W.data_ptr(),
                                          W.stride(0), dev.stream_ptr()), "gmg_krige_weights_f64")
        return W


_DIST_CTX: dict = {}


def dist_context(group=None):
    """The library's communicator handle (`gmg_dist_init`) for the ranks of `group` (default: the world group), created
    on first use - one rank draws the NCCL unique id (`gmg_dist_unique_id`), `torch.distributed` carries its 128 bytes
    to the others (plumbing), every rank joins - and kept for the life of the process.  None for a single rank."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    key = id(group) if group is not None else 0
    if key not in _DIST_CTX:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        on_gpu = dist.get_backend(group) == "nccl"
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = ctypes.create_string_buffer(128)
            check(lib().gmg_dist_unique_id(raw), "gmg_dist_unique_id")
            uid = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        if on_gpu:
            uid = uid.cuda()
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(uid, src=src, group=group)
        handle = ctypes.c_void_p()
        torch.cuda.synchronize()
        check(lib().gmg_dist_init(uid.cpu().numpy().tobytes(), world, rank, ctypes.byref(handle)), "gmg_dist_init")
        _DIST_CTX[key] = handle
    return _DIST_CTX[key]


def _dist_workspace(n: int, world: int, device) -> torch.Tensor | None:
    need = ctypes.c_int64()
    check(lib().gmg_dist_ws_doubles(n, world, ctypes.byref(need)), "gmg_dist_ws_doubles")
    return torch.empty(max(need.value, 2), dtype=torch.float64, device=device) if need.value else None


def factor_inplace_distributed(S: torch.Tensor, n: int, group=None, what: str = "matrix") -> dict:
    """The blocked Cholesky of `factor_inplace` spread over the ranks of `group` (one process per GPU).

    S is REPLICATED (every rank holds the same n x n matrix) and ends up holding the complete factor on every rank,
    which is what the column-sharded triangular solve needs.  Ownership is 1-D block-cyclic by tile row (row block r
    belongs to rank r % world).  The whole block loop runs inside the library (`gmg_potrf_solve_f64` on a
    `gmg_dist_init` handle): per block step the owner factors + inverts the 128 x 128 diagonal block and broadcasts it
    (256 KB), every rank forms and packs ITS tile rows of the panel, one all-gather of the packed tiles gives everyone
    the whole panel, every rank updates its own tile rows - the next panel column first (look-ahead), the rest on a
    second stream.  Same kernels, same per-tile arithmetic as the single-GPU path: the factor is bit-identical."""
    import torch.distributed as dist

    assert S.is_cuda and S.dtype == torch.float64 and S.stride(1) == 1 and S.stride(0) % 2 == 0
    lin, pdb, ldp = solve_ws(n)
    ws = {"Linv": torch.empty(lin, dtype=torch.float64, device=S.device),
          "P": torch.empty(pdb, dtype=torch.float64, device=S.device),
          "info": torch.zeros(1, dtype=torch.int32, device=S.device), "ldp": ldp}
    d = np.zeros(1, dtype=SOLVE_DESC)
    d["S"], d["lds"], d["Linv"], d["P"], d["ldp"], d["info"], d["n"] = (
        S.data_ptr(), S.stride(0), ws["Linv"].data_ptr(), ws["P"].data_ptr(), ldp, ws["info"].data_ptr(), n)
    dt = _desc_tensor(d, S.device)
    ctx = dist_context(group)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    xbuf = _dist_workspace(n, world, S.device)
    check(lib().gmg_potrf_solve_f64(ctx, dt.data_ptr(), 1, n, 0, xbuf.data_ptr() if xbuf is not None else None,
                                    xbuf.numel() if xbuf is not None else 0, dev.stream_ptr()), "gmg_potrf_solve_f64")
    _raise_unless_spd(ws["info"], n, group, what)
    return ws


class NotPositiveDefinite(np.linalg.LinAlgError):
    """Cholesky met a non-positive pivot (callers with an LU route catch exactly this)."""


def _raise_unless_spd(info_t: torch.Tensor, n: int, group, what: str) -> None:
    """Factorisation status of a (possibly distributed) factorisation: only the owner of a block sees its pivots."""
    import torch.distributed as dist

    info = info_t.clone()
    if dist.is_initialized() and group is not False and dist.get_world_size(group) > 1:
        dist.all_reduce(info, op=dist.ReduceOp.MAX, group=group)
    code = int(info.item())
    if code != 0:
        raise NotPositiveDefinite(
            f"{what} is not positive definite (Cholesky pivot {code} of {n} is not positive); "
            "apply glomar_gridding.covariance_tools.eigenvalue_clip or use delta_x_method='Met_Office'")


def solve_prepared(S: torch.Tensor, T: torch.Tensor, Mc: int, diag_base: torch.Tensor, ldc: int, y: torch.Tensor,
                   ordinary: bool, group=None) -> torch.Tensor:
    """Factor + forward solve + apply for ONE system whose operands are already assembled on this device:
    S (n, n) = C_obs + E (overwritten by its Cholesky factor), T (n, >= Mc + 2) = [C_cross columns | 1 | y]
    (overwritten by L^-1 T).  `diag_base` points at the element that holds C[m0, m0] for local column 0 and
    `ldc` is the stride to the next diagonal element minus one (the matrix leading dimension), so that local
    column m reads its prior variance at diag_base[m * ldc + m].  Used by the column-sharded multi-GPU path.
    With an initialised process group of more than one rank the factorisation is the distributed one
    (`factor_inplace_distributed`; `group=False` forces the local, replicated factorisation).
    Returns the (5, Mc) result rows.  Raises `NotPositiveDefinite` (a LinAlgError) if S is not positive definite - on
    every rank alike; `solve_prepared_lu` is the route for such a system."""
    n = S.shape[0]
    for name, t in (("S", S), ("T", T), ("diag_base", diag_base), ("y", y)):  # the kernels read double*
        if t.dtype != torch.float64:
            raise TypeError(f"solve_prepared: {name} must be float64, got {t.dtype}")
    assert S.stride(1) == 1 and T.stride(1) == 1 and S.stride(0) % 2 == 0 and T.stride(0) % 2 == 0
    assert T.shape[1] >= Mc + 2
    import torch.distributed as dist

    import time

    timing = bool(os.environ.get("GLOMAR_B200_TIMING"))
    if timing:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
    distributed = group is not False and dist.is_initialized() and dist.get_world_size(group) > 1
    lin, pdb, ldp = solve_ws(n)
    ws = {"Linv": torch.empty(lin, dtype=torch.float64, device=S.device),
          "P": torch.empty(pdb, dtype=torch.float64, device=S.device),
          "info": torch.zeros(1, dtype=torch.int32, device=S.device), "ldp": ldp}
    slabs = (n + SLAB - 1) // SLAB
    partial = torch.empty(4 * slabs * Mc, dtype=torch.float64, device=S.device)
    out = torch.empty((5, Mc), dtype=torch.float64, device=S.device)
    sd = np.zeros(1, dtype=SOLVE_DESC)
    sd[0] = (S.data_ptr(), S.stride(0), ws["Linv"].data_ptr(), T.data_ptr(), T.stride(0), ws["P"].data_ptr(),
             ws["ldp"], ws["info"].data_ptr(), n, Mc + 2)
    kd = np.zeros(1, dtype=KRIGE_DESC)
    kd[0] = (0, 0, n, y.data_ptr(), 0, 0, 0, 0, T.data_ptr(), T.stride(0), partial.data_ptr(), out.data_ptr(), n,
             int(ordinary))
    sdt, kdt = _desc_tensor(sd, S.device), _desc_tensor(kd, S.device)
    L, s = lib(), dev.stream_ptr()
    ctx = dist_context(group) if distributed else None
    xbuf = _dist_workspace(n, dist.get_world_size(group), S.device) if distributed else None
    # factorisation (distributed over the ranks, or local) and the forward solve of this rank's columns as one
    # overlapped pipeline, then the apply pass; the status is read afterwards
    check(L.gmg_potrf_solve_f64(ctx, sdt.data_ptr(), 1, n, Mc + 2, xbuf.data_ptr() if xbuf is not None else None,
                                xbuf.numel() if xbuf is not None else 0, s), "gmg_potrf_solve_f64")
    check(L.gmg_krige_apply_batched_f64(diag_base.data_ptr(), Mc, ldc, kdt.data_ptr(), 1, n, 1, s), "apply")
    _raise_unless_spd(ws["info"], n, group if distributed else False, "observation covariance C[idx, idx] + E")
    if timing:
        torch.cuda.synchronize()
        if not dist.is_initialized() or dist.get_rank() == 0:
            print(f"[solve_prepared] factor + forward solve + apply, n = {n}: {1e3 * (time.perf_counter() - t0):.1f} ms",
                  flush=True)
    torch.cuda.current_stream().synchronize()
    return out


def solve_prepared_lu(S: torch.Tensor, T: torch.Tensor, Mc: int, diag_base: torch.Tensor, ldc: int, y: torch.Tensor,
                      ordinary: bool) -> torch.Tensor:
    """`solve_prepared` for a system Cholesky rejected: S (FULL n x n matrix, both triangles; overwritten by its LU
    factors) is factored by the blocked LU with partial pivoting on THIS GPU (every rank of a sharded solve factors its
    own replica - 0.35 s at n = 20 000 - and solves its own columns), T <- S^-1 T, and the apply pass is the full-solve
    one (it needs the untouched C_cross columns: a copy of T[:, :Mc] is kept).  What np.linalg.solve does
    (reference kriging.py:343, :585).  Raises LinAlgError("Singular matrix") for an exactly zero pivot."""
    n = S.shape[0]
    for name, t in (("S", S), ("T", T), ("diag_base", diag_base), ("y", y)):
        if t.dtype != torch.float64:
            raise TypeError(f"solve_prepared_lu: {name} must be float64, got {t.dtype}")
    assert S.stride(1) == 1 and T.stride(1) == 1 and S.stride(0) % 2 == 0 and T.stride(0) % 2 == 0
    assert T.shape[1] >= Mc + 2
    Cx = torch.empty((n, Mc + (Mc & 1)), dtype=torch.float64, device=S.device)
    Cx[:, :Mc] = T[:, :Mc]
    lin, pdb, ldp = gesv_ws(n)
    linv = torch.empty(lin, dtype=torch.float64, device=S.device)
    strip = torch.empty(pdb, dtype=torch.float64, device=S.device)
    info = torch.zeros(1, dtype=torch.int32, device=S.device)
    slabs = (n + SLAB - 1) // SLAB
    partial = torch.empty(4 * slabs * Mc, dtype=torch.float64, device=S.device)
    out = torch.empty((5, Mc), dtype=torch.float64, device=S.device)
    sd = np.zeros(1, dtype=SOLVE_DESC)
    sd[0] = (S.data_ptr(), S.stride(0), linv.data_ptr(), T.data_ptr(), T.stride(0), strip.data_ptr(), ldp,
             info.data_ptr(), n, Mc + 2)
    kd = np.zeros(1, dtype=KRIGE_DESC)
    kd[0] = (0, 0, n, y.data_ptr(), 0, 0, Cx.data_ptr(), Cx.stride(0), T.data_ptr(), T.stride(0), partial.data_ptr(),
             out.data_ptr(), n, int(ordinary))
    sdt, kdt = _desc_tensor(sd, S.device), _desc_tensor(kd, S.device)
    L, s = lib(), dev.stream_ptr()
    check(L.gmg_gesv_blocked_f64(sdt.data_ptr(), 1, n, Mc + 2, s), "gmg_gesv_blocked_f64")
    check(L.gmg_krige_apply_batched_f64(diag_base.data_ptr(), Mc, ldc, kdt.data_ptr(), 1, n, 0, s), "apply")
    if int(info.item()) != 0:                       # (synchronises before the temporaries die)
        raise np.linalg.LinAlgError("Singular matrix")
    return out
