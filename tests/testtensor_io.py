"""Reader/writer for the reference's .testtensor container (tensor.h:201-253, utils.py:7-53).

Test infrastructure: used by tests/ and bench.py to load the golden fixtures under tests/golden/
and to build weight blobs for the per-layer parity suite.
"""
import struct
from collections import OrderedDict

import numpy as np


def load_testtensor(path_or_bytes):
    """Returns an OrderedDict name -> float32 ndarray (positional order preserved)."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    version, count = struct.unpack_from("<ii", data, 0)
    assert version == 1, version
    off = 8
    names = []
    for _ in range(count):
        (n,) = struct.unpack_from("<i", data, off)
        off += 4
        names.append(data[off:off + n].decode("utf8"))
        off += n
    out = OrderedDict()
    for i, name in enumerate(names):
        (ndim,) = struct.unpack_from("<i", data, off)
        off += 4
        dims = struct.unpack_from("<%di" % ndim, data, off) if ndim else ()
        off += 4 * ndim
        size, nbytes = struct.unpack_from("<ii", data, off)
        off += 8
        arr = np.frombuffer(data, dtype="<f4", count=size, offset=off).reshape(dims).copy()
        off += nbytes
        key = name if name not in out else "%s#%d" % (name, i)
        out[key] = arr
    assert off == len(data)
    return out


def load_list(path_or_bytes):
    return list(load_testtensor(path_or_bytes).values())


def dump_testtensor(arrays):
    """arrays: list of (name, ndarray) -> bytes in the container format."""
    blob = struct.pack("<ii", 1, len(arrays))
    for name, _ in arrays:
        enc = name.encode("utf8")
        blob += struct.pack("<i", len(enc)) + enc
    for _, a in arrays:
        a = np.ascontiguousarray(a, dtype="<f4")
        blob += struct.pack("<i", a.ndim)
        if a.ndim:
            blob += struct.pack("<%di" % a.ndim, *a.shape)
        blob += struct.pack("<ii", a.size, a.nbytes)
        blob += a.tobytes()
    return blob
