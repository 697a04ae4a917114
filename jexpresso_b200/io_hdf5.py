"""Restart files in the reference's on-disk format (src/io/write_output.jl:939-1014).

``write_hdf5`` / ``read_hdf5`` mirror the reference's functions of the same names: one ``var_<ivar>_<rank>.h5`` per
variable and rank holding the datasets ``q`` (that variable's slice of the flat state vector) and ``qe`` (the reference
state column), plus ``t.h5`` with the scalar ``time`` written by rank 0.  No HDF5 library exists in this image, so the
container format is written and parsed here directly -- exactly the subset HDF5.jl/libhdf5 emits for these files:
version-2 superblock, version-2 object headers (with Jenkins lookup3 checksums), compact link messages, contiguous
little-endian Float64 datasets, raw data from byte 2048 on.  Given the same arrays and creation time the writer reproduces
the reference's CI output files byte for byte (tests/test_restart_hdf5_cpu.py), which is the check that a real libhdf5
reads what it writes.  Host-side I/O: nothing here is on the RHS path.
"""
from __future__ import annotations

import os
import struct
import time as _time

import numpy as np

__all__ = ["write_hdf5", "read_hdf5", "read_h5_file", "write_h5_file", "lookup3"]

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_META_BLOCK = 2048           # libhdf5's default metadata block: the first raw-data byte of these small files
_ROOT_CHUNK, _DSET_CHUNK = 120, 256      # first-chunk sizes libhdf5 allocates for a new-style group / a dataset header


def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data: bytes, initval: int = 0) -> int:
    """Bob Jenkins' lookup3 ``hashlittle`` -- the metadata checksum of HDF5 (H5_checksum_lookup3)."""
    n = len(data)
    a = b = c = (0xDEADBEEF + n + initval) & 0xFFFFFFFF
    p = 0
    while n > 12:
        a = (a + int.from_bytes(data[p:p + 4], "little")) & 0xFFFFFFFF
        b = (b + int.from_bytes(data[p + 4:p + 8], "little")) & 0xFFFFFFFF
        c = (c + int.from_bytes(data[p + 8:p + 12], "little")) & 0xFFFFFFFF
        a = (a - c) & 0xFFFFFFFF; a ^= _rot(c, 4); c = (c + b) & 0xFFFFFFFF
        b = (b - a) & 0xFFFFFFFF; b ^= _rot(a, 6); a = (a + c) & 0xFFFFFFFF
        c = (c - b) & 0xFFFFFFFF; c ^= _rot(b, 8); b = (b + a) & 0xFFFFFFFF
        a = (a - c) & 0xFFFFFFFF; a ^= _rot(c, 16); c = (c + b) & 0xFFFFFFFF
        b = (b - a) & 0xFFFFFFFF; b ^= _rot(a, 19); a = (a + c) & 0xFFFFFFFF
        c = (c - b) & 0xFFFFFFFF; c ^= _rot(b, 4); b = (b + a) & 0xFFFFFFFF
        p += 12
        n -= 12
    if n == 0:
        return c
    tail = data[p:p + n] + b"\x00" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & 0xFFFFFFFF
    b = (b + int.from_bytes(tail[4:8], "little")) & 0xFFFFFFFF
    c = (c + int.from_bytes(tail[8:12], "little")) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 14)) & 0xFFFFFFFF
    a ^= c; a = (a - _rot(c, 11)) & 0xFFFFFFFF
    b ^= a; b = (b - _rot(a, 25)) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 16)) & 0xFFFFFFFF
    a ^= c; a = (a - _rot(c, 4)) & 0xFFFFFFFF
    b ^= a; b = (b - _rot(a, 14)) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 24)) & 0xFFFFFFFF
    return c


# ---- writer ------------------------------------------------------------------------------------------------------------
def _msg(mtype, body, flags=0):
    return struct.pack("<BHB", mtype, len(body), flags) + body


def _ohdr(flags, stamp, chunk, msgs):
    """Version-2 object header: times stored (flag 0x20), first chunk padded to ``chunk`` bytes by one NIL message."""
    body = b"".join(msgs)
    assert len(body) + 4 <= chunk, "object header messages exceed the first chunk"
    body += _msg(0x00, b"\x00" * (chunk - len(body) - 4))
    szlen = 1 << (flags & 3)
    head = b"OHDR" + bytes([2, flags]) + struct.pack("<IIII", stamp, stamp, stamp, stamp) + chunk.to_bytes(szlen, "little") + body
    return head + struct.pack("<I", lookup3(head))


_F64 = bytes.fromhex("11203f000800000000004000340b0034ff030000")     # datatype message: IEEE binary64, little endian


def write_h5_file(path, datasets, stamp=None):
    """``datasets``: ordered {name: 1-D float array | Python float (scalar dataspace)}.  One flat root group."""
    stamp = int(_time.time()) if stamp is None else int(stamp)
    names = list(datasets)
    raw = []
    for nme in names:
        v = datasets[nme]
        raw.append(np.ascontiguousarray(np.atleast_1d(np.asarray(v, dtype="<f8"))).tobytes() if np.ndim(v) else struct.pack("<d", float(v)))
    root_at = 48
    root_len = 4 + 2 + 16 + 1 + _ROOT_CHUNK + 4
    dset_len = 4 + 2 + 16 + 2 + _DSET_CHUNK + 4
    dset_at = [root_at + root_len + i * dset_len for i in range(len(names))]
    assert dset_at[-1] + dset_len <= _META_BLOCK, "too many datasets for one metadata block"
    data_at, off = [], _META_BLOCK
    for r in raw:
        data_at.append(off)
        off += len(r)
    eof = off
    links = [_msg(0x06, bytes([1, 0x10, 1, len(n.encode())]) + n.encode() + struct.pack("<Q", a)) for n, a in zip(names, dset_at)]
    root = _ohdr(0x20, stamp, _ROOT_CHUNK, [_msg(0x02, bytes([0, 0]) + struct.pack("<QQ", _UNDEF, _UNDEF)),      # link info
                                           _msg(0x0A, bytes([0, 0]), flags=1)] + links)                         # group info
    out = bytearray()
    sb = _SIG + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, _UNDEF, eof, root_at)
    out += sb + struct.pack("<I", lookup3(sb))
    out += root
    for nme, r, da in zip(names, raw, data_at):
        v = datasets[nme]
        if np.ndim(v):
            n = len(r) // 8
            space = bytes([2, 1, 1, 1]) + struct.pack("<QQ", n, n)          # version 2, rank 1, max dims present, simple
        else:
            space = bytes([2, 0, 0, 0])                                      # scalar
        layout = bytes([3, 1]) + struct.pack("<QQ", da, len(r))             # version 3, contiguous
        out += _ohdr(0x21, stamp, _DSET_CHUNK, [_msg(0x01, space), _msg(0x03, _F64, flags=1), _msg(0x05, bytes([3, 0x0A]), flags=1),
                                                _msg(0x08, layout)])
    out += b"\x00" * (_META_BLOCK - len(out))
    for r in raw:
        out += r
    assert len(out) == eof
    with open(path, "wb") as f:
        f.write(bytes(out))


# ---- reader ------------------------------------------------------------------------------------------------------------
def _messages(buf, addr):
    if buf[addr:addr + 4] != b"OHDR" or buf[addr + 4] != 2:
        raise ValueError("not a version-2 object header")
    flags = buf[addr + 5]
    p = addr + 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
    szlen = 1 << (flags & 3)
    chunk = int.from_bytes(buf[p:p + szlen], "little")
    p += szlen
    end = p + chunk
    if struct.unpack_from("<I", buf, end)[0] != lookup3(buf[addr:end]):
        raise ValueError("object header checksum mismatch")
    out = []
    while p + 4 <= end:
        mtype, msize = buf[p], struct.unpack_from("<H", buf, p + 1)[0]
        p += 4 + (2 if flags & 0x04 else 0)
        out.append((mtype, buf[p:p + msize]))
        p += msize
    return out


def read_h5_file(path):
    """{dataset name: float64 array (or 0-d array for a scalar)} of a flat file as write_hdf5 / the reference produce it."""
    buf = open(path, "rb").read()
    if buf[:8] != _SIG or buf[8] not in (2, 3) or buf[9] != 8 or buf[10] != 8:
        raise ValueError(f"{path}: unsupported HDF5 superblock")
    if struct.unpack_from("<I", buf, 44)[0] != lookup3(buf[:44]):
        raise ValueError(f"{path}: superblock checksum mismatch")
    root = struct.unpack_from("<Q", buf, 36)[0]
    out = {}
    for mtype, body in _messages(buf, root):
        if mtype != 0x06:
            continue
        fl, p = body[1], 2
        if fl & 0x08:
            if body[p] != 0:
                raise ValueError("only hard links")
            p += 1
        p += (8 if fl & 0x04 else 0) + (1 if fl & 0x10 else 0)
        nl = 1 << (fl & 3)
        n = int.from_bytes(body[p:p + nl], "little")
        p += nl
        name = body[p:p + n].decode()
        addr = struct.unpack_from("<Q", body, p + n)[0]
        shape, where = None, None
        for mt, b in _messages(buf, addr):
            if mt == 0x01:
                shape = struct.unpack_from("<%dQ" % b[1], b, 4) if b[1] else ()
            elif mt == 0x03 and not ((b[0] & 0x0F) == 1 and struct.unpack_from("<I", b, 4)[0] == 8):
                raise ValueError(f"{path}:{name}: not a Float64 dataset")
            elif mt == 0x08:
                if not (b[0] in (3, 4) and b[1] == 1):
                    raise ValueError(f"{path}:{name}: only contiguous layout")
                where = struct.unpack_from("<QQ", b, 2)
        arr = np.frombuffer(buf, dtype="<f8", count=where[1] // 8, offset=where[0]).copy()
        out[name] = arr.reshape(shape[::-1]).T if shape is not None and len(shape) > 1 else (arr[0] if shape == () else arr)
    return out


# ---- the reference's two functions -------------------------------------------------------------------------------------
def write_hdf5(npoin, q, qe, t, output_dir, *, nvar, rank=0, stamp=None):
    """write_hdf5 (write_output.jl:939-979): ``q`` is the flat state vector Float64[npoin*nvar] (``u`` of the ODE problem),
    ``qe`` the reference state Float64[npoin, >= nvar]; rank 0 also writes ``t.h5``."""
    q = np.asarray(q, dtype=np.float64).reshape(-1)
    qe = np.asarray(qe, dtype=np.float64)
    os.makedirs(output_dir, exist_ok=True)
    if rank == 0:
        write_h5_file(os.path.join(output_dir, "t.h5"), {"time": float(t)}, stamp)
    for ivar in range(1, nvar + 1):
        idx = (ivar - 1) * npoin
        write_h5_file(os.path.join(output_dir, f"var_{ivar}_{rank}.h5"),
                      {"q": q[idx:ivar * npoin], "qe": np.ascontiguousarray(qe[:npoin, ivar - 1])}, stamp)


def read_hdf5(input_dir, npoin, nvar, *, rank=0):
    """read_hdf5 (write_output.jl:981-1014): returns (q, qe, time) with q, qe Float64[npoin, nvar+1] (Fortran order; the last
    column stays zero, as in the reference).  ``time`` is what the reference stores into inputs[:tinit]."""
    q = np.zeros((npoin, nvar + 1), order="F")
    qe = np.zeros((npoin, nvar + 1), order="F")
    t = float(read_h5_file(os.path.join(input_dir, "t.h5"))["time"])
    for ivar in range(1, nvar + 1):
        d = read_h5_file(os.path.join(input_dir, f"var_{ivar}_{rank}.h5"))
        if d["q"].size != npoin or d["qe"].size != npoin:
            raise ValueError(f"var_{ivar}_{rank}.h5 holds {d['q'].size} nodes, expected {npoin}")
        q[:, ivar - 1] = d["q"]
        qe[:, ivar - 1] = d["qe"]
    return q, qe, t
