"""The reference's OWN golden vector for the ellipse covariance path, re-packed as a small .npz:

    test/data/Atlantic_Ocean_07.nc   fitted ellipse parameters (Lx, Ly, theta, standard_deviation on a 120 x 120 grid)
    test/data/cov_no_hfix.nc         the covariance (262 x 262, float32) the reference's test expects for the
                                     [50:70, 50:70] window, nu = 0.5 (`test/test_ellipse.py:177-217`, rtol 1e-5)

Both files are netCDF-4 = HDF5 (superblock v2, version-2 object headers with `OCHK` continuation blocks); every
variable is stored CONTIGUOUSLY and unfiltered, so a few dozen lines of stdlib code read them - no HDF5 / netCDF library
exists in this image.  Variables are identified by their `long_name` attribute (the reference's loader uses the link
name; `Lx` is linked as "lx", ...).  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_atlantic.py

The script also checks that the unmodified reference reproduces its golden here (numpy 2.3.5 / scipy 1.18.1).
"""

from __future__ import annotations

import contextlib
import io
import os
import re
import struct
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def _messages(b: bytes, off: int):
    """(type, body) of every header message of the version-2 object header at `off`, continuation blocks included."""
    assert b[off:off + 4] == b"OHDR" and b[off + 4] == 2
    flags = b[off + 5]
    p = off + 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
    szlen = 1 << (flags & 3)
    chunk0 = int.from_bytes(b[p:p + szlen], "little")
    p += szlen
    out = []

    def walk(p, end):
        while p + 4 <= end:
            mtype, size = b[p], struct.unpack_from("<H", b, p + 1)[0]
            p += 4 + (2 if flags & 0x04 else 0)
            body = b[p:p + size]
            if mtype == 0x10:                                   # object header continuation
                addr, length = struct.unpack_from("<QQ", body, 0)
                assert b[addr:addr + 4] == b"OCHK"
                walk(addr + 4, addr + length - 4)
            else:
                out.append((mtype, body))
            p += size

    walk(p, p + chunk0)
    return out


def datasets(path: str) -> dict:
    """long_name (or "shape/dtype@offset" when there is none) -> ndarray for every contiguous dataset of the file."""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 2
    out = {}
    for m in re.finditer(b"OHDR", b):
        shape = dtype = layout = name = None
        for mtype, body in _messages(b, m.start()):
            if mtype == 0x01:                                   # dataspace
                rank = body[1]
                shape = struct.unpack_from("<" + "Q" * rank, body, 4 if body[0] == 2 else 8)
            elif mtype == 0x03:                                 # datatype: class 0 fixed point, 1 floating point
                cls, size = body[0] & 0x0F, struct.unpack_from("<I", body, 4)[0]
                dtype = {(1, 4): "<f4", (1, 8): "<f8", (0, 8): "<i8", (0, 4): "<i4"}.get((cls, size))
            elif mtype == 0x08 and body[1] == 1:                # data layout, contiguous: address, size
                layout = struct.unpack_from("<QQ", body, 2)
            elif mtype == 0x0C and body[0] == 3:                # attribute (version 3): fixed-length string values
                nlen, dts, dss = struct.unpack_from("<HHH", body, 2)
                if body[9:9 + nlen - 1] == b"long_name":
                    name = body[9 + nlen + dts + dss:].split(b"\x00")[0].decode()
        if shape is None or dtype is None or layout is None or layout[0] >= len(b):
            continue
        arr = np.frombuffer(b, dtype=dtype, count=int(np.prod(shape)) if shape else 1, offset=layout[0]).reshape(shape)
        out[name or f"{shape}/{dtype}@{m.start()}"] = arr
    return out


def main():
    from oracle.reference_shim import REFERENCE_ROOT, load_reference

    data = os.path.join(REFERENCE_ROOT, "test", "data")
    atl = datasets(os.path.join(data, "Atlantic_Ocean_07.nc"))
    gold = datasets(os.path.join(data, "cov_no_hfix.nc"))
    coords = sorted((k for k in atl if k.startswith("(120,)/<f8")), key=lambda k: int(k.split("@")[1]))
    lats, lons = atl[coords[0]], atl[coords[1]]                 # latitude -59.875 .. 59.125, longitude 260.125 .. 379.125
    assert lats[0] < -59 and lons[0] > 260
    win = (slice(50, 70), slice(50, 70))                        # the window of test/test_ellipse.py:189-199
    Lx, Ly, theta, stdev = (np.array(atl[k][win]) for k in ("Lx", "Ly", "theta", "standard_deviation"))
    cov = np.array(gold["covariance"])
    mask = Lx > 1e5
    ref = load_reference()
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = ref.EllipseCovarianceBuilder(np.ma.masked_where(mask, Lx), np.ma.masked_where(mask, Ly),
                                         np.ma.masked_where(mask, theta), np.ma.masked_where(mask, stdev),
                                         lats[50:70], lons[50:70], v=0.5)
    assert b.cov_ns.shape == cov.shape and np.allclose(b.cov_ns, cov, rtol=1e-5), "reference does not reproduce its golden"
    print("reference vs its golden: max rel", float(np.max(np.abs(b.cov_ns - cov) / np.abs(cov))))
    iu = np.triu_indices(cov.shape[0])
    assert np.array_equal(cov, cov.T)
    np.savez_compressed(os.path.join(HERE, "ellipse_atlantic.npz"), Lx=Lx, Ly=Ly, theta=theta, stdev=stdev,
                        lats=np.array(lats[50:70]), lons=np.array(lons[50:70]), cov_triu=cov[iu],
                        cor_diag=np.array(np.diag(gold["correlation"])), matern_nu=np.float64(gold["matern_nu"]))
    print("wrote ellipse_atlantic.npz", cov.shape, int((~mask).sum()), "unmasked cells")


if __name__ == "__main__":
    main()
