"""Pure-Python reader for TensorFlow "bundle v2" checkpoints (TEST INFRASTRUCTURE ONLY).

The reference ships TF-1.14 checkpoints under ``save/`` but TensorFlow itself
cannot be installed here.  The ``.index`` file is an uncompressed LevelDB-format table whose
values are ``BundleEntryProto`` messages; the ``.data-00000-of-00001`` file is raw
little-endian tensor bytes.  This module decodes both with nothing but ``struct`` and numpy so
that the reference's saved tensors can be used as golden vectors (``tests/golden/make_golden.py``).

Nothing under ``oracle/`` is imported by the product path (``multimodaltraj_2_b200``).
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_MAGIC = 0xDB4775248B80FB57  # leveldb table magic, little-endian at the end of the footer

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf: bytes, pos: int):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block_entries(buf: bytes, off: int, size: int):
    """Yield (key, value) of one table block (prefix-compressed keys, restart array at the tail)."""
    blk = buf[off:off + size]
    n_restarts = struct.unpack_from("<I", blk, size - 4)[0]
    end = size - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(blk, pos)
        non_shared, pos = _varint(blk, pos)
        vlen, pos = _varint(blk, pos)
        key = key[:shared] + blk[pos:pos + non_shared]
        pos += non_shared
        yield key, blk[pos:pos + vlen]
        pos += vlen


def _parse_proto(msg: bytes):
    """Minimal protobuf wire decoder -> {field: [values]} (varint / 64-bit / bytes / 32-bit)."""
    out, pos = {}, 0
    while pos < len(msg):
        tag, pos = _varint(msg, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(msg, pos)
        elif wt == 1:
            v = msg[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(msg, pos)
            v = msg[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = msg[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _shape(tsp: bytes):
    dims = []
    for d in _parse_proto(tsp).get(2, []):
        f = _parse_proto(d)
        dims.append(f.get(1, [0])[0])
    return tuple(dims)


def read_index(prefix: str | Path):
    """Return {tensor_name: (np.dtype, shape, offset, size)} for checkpoint ``prefix``."""
    buf = Path(str(prefix) + ".index").read_bytes()
    footer = buf[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
        raise ValueError("not a leveldb table")
    _, p = _varint(footer, 0)          # metaindex offset
    _, p = _varint(footer, p)          # metaindex size
    ioff, p = _varint(footer, p)
    isz, p = _varint(footer, p)
    entries = {}
    for _, handle in _block_entries(buf, ioff, isz):
        doff, q = _varint(handle, 0)
        dsz, q = _varint(handle, q)
        for key, val in _block_entries(buf, doff, dsz):
            if not key:                # "" -> BundleHeaderProto
                continue
            f = _parse_proto(val)
            dt = _DTYPES[f.get(1, [1])[0]]
            shape = _shape(f[2][0]) if 2 in f else ()
            entries[key.decode()] = (np.dtype(dt), shape, f.get(4, [0])[0], f.get(5, [0])[0])
    return entries


def read_checkpoint(prefix: str | Path):
    """Return {tensor_name: ndarray} for every tensor in checkpoint ``prefix``."""
    idx = read_index(prefix)
    data = Path(str(prefix) + ".data-00000-of-00001").read_bytes()
    out = {}
    for name, (dt, shape, off, size) in idx.items():
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(data, dtype=dt, count=n, offset=off) if n else np.zeros(0, dt)
        assert n * dt.itemsize == size, (name, shape, size)
        out[name] = arr.reshape(shape).copy()
    return out


if __name__ == "__main__":
    import sys
    for k, v in sorted(read_checkpoint(sys.argv[1]).items()):
        print(f"{k:60s} {str(v.dtype):8s} {v.shape}  std={v.std() if v.size else 0:.4g}")
