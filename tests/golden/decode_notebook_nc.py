"""Stdlib-only decoder for the chunked float32 variables of the reference's example-notebook netCDF-4 files
(`/root/reference/notebooks/*.nc`, HDF5 superblock v2) - no HDF5 / netCDF library exists in this image.

Used ONLY by `make_golden_notebook.py` to build the configs[0] fixture (the HadSST.4.0.1.0 2014-03 ensemble-member
field the notebook `notebooks/Ellipse_to_grid.ipynb` grids, cells 19-34).  Each of these variables is stored as
chunks indexed by a version-1 B-tree leaf (`TREE`, node type 1, level 0); every chunk went through the shuffle
filter (id 2) and deflate (id 1).  Entry layout: chunk_size u32, filter_mask u32, offsets u64 x (rank + 1),
child address u64.
"""
import struct
import zlib

import numpy as np


def _tree_leaf(buf: bytes):
    i = buf.find(b"TREE")
    if i < 0:
        raise ValueError("no chunk B-tree in file")
    ntype, level, nent = struct.unpack_from("<BBH", buf, i + 4)
    if ntype != 1 or level != 0:
        raise ValueError("expected a single level-0 chunk B-tree leaf")
    return i + 24, nent          # signature 4 + type/level/entries 4 + left/right siblings 16


def read_chunked_f32(path: str, shape, chunk) -> np.ndarray:
    """Variable of `shape` stored in chunks of `chunk` (both tuples of the same rank), float32, shuffle + deflate."""
    buf = open(path, "rb").read()
    rank = len(shape)
    pos, nent = _tree_leaf(buf)
    out = np.zeros(shape, dtype="<f4")
    key = 8 + 8 * (rank + 1)
    for _ in range(nent):
        size, mask = struct.unpack_from("<II", buf, pos)
        offs = struct.unpack_from("<" + "Q" * (rank + 1), buf, pos + 8)
        (addr,) = struct.unpack_from("<Q", buf, pos + key)
        pos += key + 8
        raw = buf[addr:addr + size]
        if not mask & 1:                       # bit i set = filter i of the pipeline was skipped
            raw = zlib.decompress(raw)
        n = int(np.prod(chunk))
        data = np.frombuffer(raw, dtype=np.uint8)
        if not mask & 2:
            data = data.reshape(4, n).T.copy()  # un-shuffle: byte planes -> little-endian words
        block = data.reshape(-1).view("<f4").reshape(chunk)
        sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs[:rank], chunk, shape))
        out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
    return out
