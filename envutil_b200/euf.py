"""Reader/writer for the trivial ".euf" float raster container used by the host tools,
the oracle builds and the tests (OpenImageIO is not available in this image).

Layout: b"EUF1", int32 width, int32 height, int32 nchannels, then height*width*nchannels
little-endian float32 values, row-major, top row first, channels interleaved - the same
in-memory layout envutil hands to OIIO (reference envutil_basic.h:760-775).
"""
import struct

import numpy as np

MAGIC = b"EUF1"


def write_euf(path, img):
    a = np.ascontiguousarray(img, dtype=np.float32)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<iii", w, h, c))
        a.tofile(f)


def read_euf(path):
    with open(path, "rb") as f:
        if f.read(4) != MAGIC:
            raise ValueError(f"{path}: not an EUF1 file")
        w, h, c = struct.unpack("<iii", f.read(12))
        a = np.fromfile(f, dtype=np.float32, count=w * h * c)
    if a.size != w * h * c:
        raise ValueError(f"{path}: truncated")
    return a.reshape(h, w, c)
