"""TensorFlow "bundle v2" checkpoints without TensorFlow: writer and reader.

The reference saves and restores its variables with ``tf.train.Saver`` (train.py:330-343 save every
``save_every`` batches, train.py:383-402 / sample.py:150-170 restore).  A checkpoint ``prefix`` is two files:

``prefix.data-00000-of-00001``  the tensors' raw little-endian bytes, in ascending key order, back to back;
``prefix.index``                a LevelDB-format table (no compression, restart interval 16, one filter-less
                                metaindex block) mapping "" -> BundleHeaderProto and every tensor name ->
                                BundleEntryProto (dtype, shape, offset, size, masked CRC-32C of the bytes).

``write_checkpoint`` reproduces the files TF 1.14 writes byte for byte (tests: the reference's own checkpoints under
``save/`` are read, written again and compared), so checkpoints written here load in the reference's ``Saver`` and
vice versa; ``save_state`` / ``latest_checkpoint`` keep the ``checkpoint`` state file next to them.
Host-side only (numpy + struct): checkpoint I/O is not on the device hot path.
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_RESTART_INTERVAL = 16
# tensorflow/core/framework/types.proto
_DT_OF = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}
_NP_OF = {v: k for k, v in _DT_OF.items()}


# ---------------------------------------------------------------------------------------------- CRC-32C (Castagnoli)
def _make_table():
    tab = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_TAB = _make_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    tab = _TAB
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    """leveldb / TF mask: rotate right by 15 and add a constant (CRCs of data that embeds CRCs)."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- varint / protobuf
def _put_varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _field_varint(field: int, v: int) -> bytes:
    return _put_varint(field << 3) + _put_varint(v)


def _field_bytes(field: int, b: bytes) -> bytes:
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


def _shape_proto(shape) -> bytes:
    # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }; proto3: a zero size is not serialised
    return b"".join(_field_bytes(2, _field_varint(1, int(d)) if d else b"") for d in shape)


def _entry_proto(dtype: np.dtype, shape, offset: int, size: int, crc: int) -> bytes:
    # BundleEntryProto { dtype = 1; shape = 2; shard_id = 3; offset = 4; size = 5; fixed32 crc32c = 6 }
    out = _field_varint(1, _DT_OF[np.dtype(dtype)]) + _field_bytes(2, _shape_proto(shape))
    if offset:
        out += _field_varint(4, offset)
    if size:
        out += _field_varint(5, size)
    return out + _put_varint((6 << 3) | 5) + struct.pack("<I", crc)


# BundleHeaderProto { num_shards = 1; endianness = 2 (LITTLE = 0, omitted); VersionDef version = 3 { producer = 1 } }
_HEADER = _field_varint(1, 1) + _field_bytes(3, _field_varint(1, 1))


def _parse_proto(msg: bytes):
    out, pos = {}, 0
    while pos < len(msg):
        tag, pos = _get_varint(msg, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(msg, pos)
        elif wt == 1:
            v, pos = msg[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _get_varint(msg, pos)
            v, pos = msg[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = msg[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


# ---------------------------------------------------------------------------------------------- table blocks
def _build_block(items) -> bytes:
    """items: [(key, value)] in ascending key order -> block contents (prefix-compressed keys + restart array)."""
    out, restarts, last = bytearray(), [], b""
    for n, (key, val) in enumerate(items):
        if n % _RESTART_INTERVAL == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            lim = min(len(last), len(key))
            while shared < lim and last[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val)) + key[shared:] + val
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _with_trailer(block: bytes) -> bytes:
    return block + b"\x00" + struct.pack("<I", masked_crc32c(block + b"\x00"))   # type 0 = no compression


def _short_successor(key: bytes) -> bytes:
    """leveldb BytewiseComparator::FindShortSuccessor: first byte that can be incremented, truncated after it."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def _read_block(buf: bytes, off: int, size: int):
    blk = buf[off:off + size]
    if struct.unpack_from("<I", buf, off + size + 1)[0] != masked_crc32c(blk + buf[off + size:off + size + 1]):
        raise ValueError("table block checksum mismatch")
    n_restarts = struct.unpack_from("<I", blk, size - 4)[0]
    end = size - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(blk, pos)
        non_shared, pos = _get_varint(blk, pos)
        vlen, pos = _get_varint(blk, pos)
        key = key[:shared] + blk[pos:pos + non_shared]
        pos += non_shared
        yield key, blk[pos:pos + vlen]
        pos += vlen


# ---------------------------------------------------------------------------------------------- public API
def write_checkpoint(prefix, tensors: dict) -> None:
    """Write ``{name: ndarray}`` as ``prefix.index`` + ``prefix.data-00000-of-00001`` (what ``Saver.save`` produces)."""
    prefix = str(prefix)
    names = sorted(tensors, key=lambda s: s.encode())
    data, items, offset = bytearray(), [(b"", _HEADER)], 0
    for name in names:
        a = np.asarray(tensors[name])                 # (ascontiguousarray would turn a scalar into shape (1,))
        if not a.flags.c_contiguous:
            a = a.copy()
        if a.dtype not in _DT_OF:
            raise TypeError(f"{name}: dtype {a.dtype} not supported (float32 / float64 / int32 / int64)")
        raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
        items.append((name.encode(), _entry_proto(a.dtype, a.shape, offset, len(raw), masked_crc32c(raw))))
        data += raw
        offset += len(raw)
    out = bytearray()
    block = _build_block(items)                       # index files are far below the 256 KB block size: one data block
    data_handle = _put_varint(0) + _put_varint(len(block))
    out += _with_trailer(block)
    meta_off = len(out)
    meta = _build_block([])
    out += _with_trailer(meta)
    index_off = len(out)
    index = _build_block([(_short_successor(items[-1][0]), data_handle)])
    out += _with_trailer(index)
    footer = _put_varint(meta_off) + _put_varint(len(meta)) + _put_varint(index_off) + _put_varint(len(index))
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    Path(prefix + ".data-00000-of-00001").write_bytes(bytes(data))
    Path(prefix + ".index").write_bytes(bytes(out))


def read_index(prefix) -> dict:
    """{name: (dtype, shape, offset, size, masked crc)} of checkpoint ``prefix`` (block checksums verified)."""
    buf = Path(str(prefix) + ".index").read_bytes()
    footer = buf[-48:]
    if len(buf) < 48 or struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
        raise ValueError(f"{prefix}.index is not a TensorFlow bundle index")
    _, p = _get_varint(footer, 0)
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isz, p = _get_varint(footer, p)
    entries = {}
    for _, handle in _read_block(buf, ioff, isz):
        doff, q = _get_varint(handle, 0)
        dsz, q = _get_varint(handle, q)
        for key, val in _read_block(buf, doff, dsz):
            if not key:
                continue
            f = _parse_proto(val)
            dims = tuple(_parse_proto(d).get(1, [0])[0] for d in _parse_proto(f[2][0]).get(2, [])) if 2 in f else ()
            entries[key.decode()] = (_NP_OF[f.get(1, [1])[0]], dims, f.get(4, [0])[0], f.get(5, [0])[0],
                                     struct.unpack("<I", f[6][0])[0] if 6 in f else None)
    return entries


def read_checkpoint(prefix, verify: bool = True) -> dict:
    """{name: ndarray} of checkpoint ``prefix``; ``verify`` checks every tensor's CRC-32C."""
    idx = read_index(prefix)
    data = Path(str(prefix) + ".data-00000-of-00001").read_bytes()
    out = {}
    for name, (dt, shape, off, size, crc) in idx.items():
        raw = data[off:off + size]
        n = int(np.prod(shape)) if shape else 1
        if len(raw) != size or n * dt.itemsize != size:
            raise ValueError(f"{name}: {size} bytes recorded for shape {shape} {dt}")
        if verify and crc is not None and masked_crc32c(raw) != crc:
            raise ValueError(f"{name}: tensor checksum mismatch")
        out[name] = np.frombuffer(raw, dtype=dt.newbyteorder("<")).astype(dt).reshape(shape)
    return out


def save_state(directory, prefix_name: str, keep=()) -> None:
    """The ``checkpoint`` state file ``Saver.save`` maintains (train.py:341 passes ``global_step``)."""
    lines = [f'model_checkpoint_path: "{prefix_name}"']
    lines += [f'all_model_checkpoint_paths: "{p}"' for p in (*keep, prefix_name)]
    (Path(directory) / "checkpoint").write_text("\n".join(lines) + "\n")


def latest_checkpoint(directory):
    """``tf.train.latest_checkpoint`` / ``get_checkpoint_state`` (train.py:383-386): prefix path or None."""
    f = Path(directory) / "checkpoint"
    if not f.exists():
        return None
    for line in f.read_text().splitlines():
        if line.startswith("model_checkpoint_path:"):
            name = line.split(":", 1)[1].strip().strip('"')
            p = Path(name)
            return str(p if p.is_absolute() else Path(directory) / p)
    return None
