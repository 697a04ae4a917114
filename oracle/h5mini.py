"""Minimal HDF5 reader -- TEST INFRASTRUCTURE.

Reads the files Jexpresso's write_hdf5 produces (src/io/write_output.jl:939-979): version-2
superblock, one root object header (OHDR v2) holding link messages to a handful of
contiguous little-endian Float64 datasets.  No h5py exists in this image.
Only what those golden files need is implemented.
"""
from __future__ import annotations

import struct

import numpy as np


def _ohdr_messages(buf, addr):
    assert buf[addr:addr + 4] == b"OHDR", "not a v2 object header"
    ver, flags = buf[addr + 4], buf[addr + 5]
    assert ver == 2
    p = addr + 6
    if flags & 0x20:
        p += 16                       # access/mod/change/birth times
    if flags & 0x10:
        p += 4                        # max compact / min dense attributes
    szlen = 1 << (flags & 0x3)
    chunk0 = int.from_bytes(buf[p:p + szlen], "little")
    p += szlen
    end = p + chunk0
    msgs = []
    while p + 4 <= end:
        mtype = buf[p]
        msize = struct.unpack_from("<H", buf, p + 1)[0]
        p += 4
        if flags & 0x04:
            p += 2                    # creation order
        msgs.append((mtype, buf[p:p + msize]))
        p += msize
    return msgs


def _parse_link(body):
    ver, flags = body[0], body[1]
    p = 2
    ltype = 0
    if flags & 0x08:
        ltype = body[p]; p += 1
    if flags & 0x04:
        p += 8
    if flags & 0x10:
        p += 1
    nlen_size = 1 << (flags & 0x3)
    nlen = int.from_bytes(body[p:p + nlen_size], "little")
    p += nlen_size
    name = body[p:p + nlen].decode()
    p += nlen
    assert ltype == 0, "only hard links"
    return name, struct.unpack_from("<Q", body, p)[0]


def read_h5(path):
    """Return {dataset name: np.ndarray(float64)} for a flat Jexpresso output file."""
    buf = open(path, "rb").read()
    assert buf[:8] == b"\x89HDF\r\n\x1a\n" and buf[8] in (2, 3), "unsupported superblock"
    assert buf[9] == 8 and buf[10] == 8
    root = struct.unpack_from("<Q", buf, 12 + 8 * 3)[0]
    out = {}
    for mtype, body in _ohdr_messages(buf, root):
        if mtype != 0x06:
            continue
        name, addr = _parse_link(body)
        shape, daddr, dsize, dt_ok = None, None, None, False
        for mt, b in _ohdr_messages(buf, addr):
            if mt == 0x01:            # dataspace v2
                assert b[0] == 2
                rank = b[1]
                shape = struct.unpack_from("<%dQ" % rank, b, 4)
            elif mt == 0x03:          # datatype: class 1 (float), 8 bytes
                dt_ok = (b[0] & 0x0F) == 1 and struct.unpack_from("<I", b, 4)[0] == 8
            elif mt == 0x08:          # layout v3/v4 contiguous
                assert b[0] in (3, 4) and b[1] == 1, "only contiguous layout"
                daddr, dsize = struct.unpack_from("<QQ", b, 2)
        assert dt_ok and daddr is not None
        arr = np.frombuffer(buf, dtype="<f8", count=dsize // 8, offset=daddr).copy()
        out[name] = arr.reshape(shape[::-1]).T if shape and len(shape) > 1 else arr
    return out
