"""ABMP flat problem container (reader/writer).

Python twin of ``include/abmcmc_b200/abmp_io.hpp``: a list of named little-endian arrays holding what
the reference's ``SparseBasisSampler`` constructor consumes (SparseBasisSampler.h:107-185).  Files may be
xz-compressed (``*.abmp.xz``), which is how fixtures are committed under ``tests/golden``.
"""
from __future__ import annotations

import lzma
import struct
from pathlib import Path
from typing import Dict

import numpy as np

_DTYPES = {0: np.dtype("<i4"), 1: np.dtype("<f8"), 2: np.dtype("u1"), 3: np.dtype("i1")}
_CODES = {np.dtype("int32"): 0, np.dtype("float64"): 1, np.dtype("uint8"): 2, np.dtype("int8"): 3}
MAGIC = b"ABMPROB1"


def read_abmp(path) -> Dict[str, np.ndarray]:
    raw = Path(path).read_bytes()
    if str(path).endswith(".xz"):
        raw = lzma.decompress(raw)
    return parse_abmp(raw)


def parse_abmp(raw: bytes) -> Dict[str, np.ndarray]:
    if raw[:8] != MAGIC:
        raise ValueError("not an ABMP file")
    (n,) = struct.unpack_from("<I", raw, 8)
    pos = 12
    out: Dict[str, np.ndarray] = {}
    for _ in range(n):
        name = raw[pos:pos + 24].split(b"\0", 1)[0].decode()
        dtype, count = struct.unpack_from("<IQ", raw, pos + 24)
        pos += 36
        dt = _DTYPES[dtype]
        nbytes = count * dt.itemsize
        out[name] = np.frombuffer(raw, dtype=dt, count=count, offset=pos).copy()
        pos += (nbytes + 7) & ~7
    return out


def dump_abmp(arrays: Dict[str, np.ndarray]) -> bytes:
    parts = [MAGIC, struct.pack("<I", len(arrays))]
    for name, arr in arrays.items():
        arr = np.ascontiguousarray(arr)
        code = _CODES[arr.dtype]
        data = arr.tobytes()
        parts.append(name.encode()[:23].ljust(24, b"\0"))
        parts.append(struct.pack("<IQ", code, arr.size))
        parts.append(data)
        parts.append(b"\0" * (((len(data) + 7) & ~7) - len(data)))
    return b"".join(parts)


def write_abmp(path, arrays: Dict[str, np.ndarray]) -> None:
    raw = dump_abmp(arrays)
    if str(path).endswith(".xz"):
        raw = lzma.compress(raw, preset=9 | lzma.PRESET_EXTREME)
    Path(path).write_bytes(raw)
