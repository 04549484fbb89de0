"""TensorFlow tensor-bundle ("Saver V2") checkpoints, read and written without TensorFlow.

The reference keeps its weights in `tf.train.Saver` checkpoints (train.py:180-185 builds the saver over the trainable
variables plus every `moving_mean` / `moving_variance`; train.py:266-267 writes `p3d_<step>.ckpt`; gen_pred.py:57-64 and
test.py restore the newest one named by the directory's `checkpoint` state file).  This module reads and writes that
on-disk format directly so that weights move between the reference and this implementation by VARIABLE NAME (the builders
create the reference's names, see p3d.py), with no TensorFlow in the process:

  <prefix>.index                 an SSTable (LevelDB table format): key "" -> BundleHeaderProto, key <variable name> ->
                                 BundleEntryProto {dtype, shape, shard_id, offset, size, masked crc32c}; prefix-compressed
                                 keys with a restart point every 16 entries, each block followed by a 1-byte compression
                                 tag and a masked CRC-32C, a 48-byte footer ending in the magic 0xdb4775248b80fb57
  <prefix>.data-00000-of-00001   the raw little-endian tensor bytes back to back, in key order
  checkpoint                     text state file: model_checkpoint_path / all_model_checkpoint_paths

Parity status: format restated from TensorFlow 1.x's tensor_bundle / table sources as published (the dependency is absent
from /root/reference and from this image) -> **parity unpinned**: there is no TF here to round-trip against.  The tests
pin what can be pinned: CRC-32C known answers, byte-exact header/footer constants, write->read round trips, and a
hand-assembled index block that uses prefix compression and several restart points.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

from . import _abi as A

TABLE_MAGIC = 0xDB4775248B80FB57
BLOCK_RESTART_INTERVAL = 16
BLOCK_SIZE = 256 << 10          # tensor_bundle's table option; any value reads back
MASK_DELTA = 0xA282EAD8
HEADER_BYTES = b"\x08\x01\x1a\x02\x08\x01"   # BundleHeaderProto{num_shards: 1, endianness: LITTLE(0, omitted), version{producer: 1}}

# tensorflow DataType enum values <-> numpy
_DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"), 6: np.dtype("i1"),
       9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 19: np.dtype("<f2"), 22: np.dtype("<u4"), 23: np.dtype("<u8")}
_DT_BFLOAT16 = 14
_NP2DT = {v: k for k, v in _DT.items()}


class CheckpointError(RuntimeError):
    pass


# ---------------------------------------------------------------------------------------------------------------------
# checksums and varints
# ---------------------------------------------------------------------------------------------------------------------
def crc32c(data, crc: int = 0) -> int:
    """CRC-32C through the library's host helper (sap3d_crc32c)."""
    if isinstance(data, np.ndarray):
        data = np.ascontiguousarray(data)
        return A.lib.sap3d_crc32c(crc, data.ctypes.data_as(C.c_void_p), data.nbytes)
    b = bytes(data)
    return A.lib.sap3d_crc32c(crc, C.cast(C.c_char_p(b), C.c_void_p), len(b))


def mask_crc(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(masked: int) -> int:
    rot = (masked - MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


def _put_varint(out: bytearray, v: int):
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = v = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if b < 0x80:
            return v, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


# ---------------------------------------------------------------------------------------------------------------------
# the two protobuf messages, hand-encoded (fields in tag order, proto3 zero defaults omitted)
# ---------------------------------------------------------------------------------------------------------------------
def _encode_entry(dtype: int, shape: Tuple[int, ...], offset: int, size: int, crc_masked: int) -> bytes:
    out = bytearray()
    out += b"\x08"
    _put_varint(out, dtype)                                    # 1: dtype
    shp = bytearray()
    for d in shape:                                            # TensorShapeProto.dim (field 2) {size: field 1}
        dim = bytearray(b"\x08")
        _put_varint(dim, d)
        shp += b"\x12"
        _put_varint(shp, len(dim))
        shp += dim
    out += b"\x12"
    _put_varint(out, len(shp))                                 # 2: shape (present even for scalars)
    out += shp
    if offset:                                                 # 3: shard_id = 0 omitted; 4: offset
        out += b"\x20"
        _put_varint(out, offset)
    if size:
        out += b"\x28"
        _put_varint(out, size)                                 # 5: size
    out += b"\x35" + struct.pack("<I", crc_masked)             # 6: crc32c (fixed32)
    return bytes(out)


def _skip_field(buf: bytes, pos: int, wire: int) -> int:
    if wire == 0:
        return _get_varint(buf, pos)[1]
    if wire == 1:
        return pos + 8
    if wire == 2:
        n, pos = _get_varint(buf, pos)
        return pos + n
    if wire == 5:
        return pos + 4
    raise CheckpointError(f"unsupported protobuf wire type {wire}")


def _decode_shape(buf: bytes) -> Tuple[int, ...]:
    dims: List[int] = []
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        if tag == 0x12:                                         # dim
            n, pos = _get_varint(buf, pos)
            sub, end = buf[pos:pos + n], pos + n
            size, q = 0, 0
            while q < len(sub):
                t, q = _get_varint(sub, q)
                if t == 0x08:
                    size, q = _get_varint(sub, q)
                    if size >= 1 << 63:
                        size -= 1 << 64
                else:
                    q = _skip_field(sub, q, t & 7)
            dims.append(size)
            pos = end
        else:
            pos = _skip_field(buf, pos, tag & 7)
    return tuple(dims)


def _decode_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if field == 1 and wire == 0:
            e["dtype"], pos = _get_varint(buf, pos)
        elif field == 2 and wire == 2:
            n, pos = _get_varint(buf, pos)
            e["shape"] = _decode_shape(buf[pos:pos + n])
            pos += n
        elif field == 3 and wire == 0:
            e["shard_id"], pos = _get_varint(buf, pos)
        elif field == 4 and wire == 0:
            e["offset"], pos = _get_varint(buf, pos)
        elif field == 5 and wire == 0:
            e["size"], pos = _get_varint(buf, pos)
        elif field == 6 and wire == 5:
            e["crc32c"] = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif field == 7:
            e["sliced"] = True
            pos = _skip_field(buf, pos, wire)
        else:
            pos = _skip_field(buf, pos, wire)
    return e


def _decode_header(buf: bytes) -> dict:
    h = {"num_shards": 0, "endianness": 0, "producer": 0}
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if field == 1 and wire == 0:
            h["num_shards"], pos = _get_varint(buf, pos)
        elif field == 2 and wire == 0:
            h["endianness"], pos = _get_varint(buf, pos)
        elif field == 3 and wire == 2:
            n, pos = _get_varint(buf, pos)
            sub, q = buf[pos:pos + n], 0
            while q < len(sub):
                t, q = _get_varint(sub, q)
                if t == 0x08:
                    h["producer"], q = _get_varint(sub, q)
                else:
                    q = _skip_field(sub, q, t & 7)
            pos += n
        else:
            pos = _skip_field(buf, pos, wire)
    return h


# ---------------------------------------------------------------------------------------------------------------------
# SSTable blocks
# ---------------------------------------------------------------------------------------------------------------------
class _BlockBuilder:
    def __init__(self, restart_interval: int = BLOCK_RESTART_INTERVAL):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last_key = b""
        self.interval = restart_interval
        self.entries = 0

    def add(self, key: bytes, value: bytes):
        assert self.entries == 0 or key > self.last_key, "keys must be added in increasing order"
        shared = 0
        if self.count < self.interval:
            m = min(len(key), len(self.last_key))
            while shared < m and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        _put_varint(self.buf, shared)
        _put_varint(self.buf, len(key) - shared)
        _put_varint(self.buf, len(value))
        self.buf += key[shared:]
        self.buf += value
        self.last_key = key
        self.count += 1
        self.entries += 1

    def size_estimate(self) -> int:
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _block_with_trailer(contents: bytes) -> bytes:
    trailer_type = b"\x00"                                      # kNoCompression
    crc = crc32c(trailer_type, crc32c(contents))
    return contents + trailer_type + struct.pack("<I", mask_crc(crc))


def _handle(offset: int, size: int) -> bytes:
    out = bytearray()
    _put_varint(out, offset)
    _put_varint(out, size)
    return bytes(out)


def build_table(items: Iterable[Tuple[bytes, bytes]], block_size: int = BLOCK_SIZE) -> bytes:
    """serialises sorted (key, value) pairs as an SSTable: data blocks, an empty metaindex block, the index block, footer."""
    out = bytearray()
    index = _BlockBuilder(restart_interval=1)
    cur = _BlockBuilder()

    def flush():
        nonlocal cur
        if cur.entries == 0:
            return
        contents = cur.finish()
        index.add(cur.last_key, _handle(len(out), len(contents)))   # separator = the block's last key
        out.extend(_block_with_trailer(contents))
        cur = _BlockBuilder()

    for k, v in items:
        cur.add(k, v)
        if cur.size_estimate() >= block_size:
            flush()
    flush()
    meta = _BlockBuilder().finish()
    meta_h = _handle(len(out), len(meta))
    out.extend(_block_with_trailer(meta))
    idx = index.finish()
    idx_h = _handle(len(out), len(idx))
    out.extend(_block_with_trailer(idx))
    footer = (meta_h + idx_h).ljust(40, b"\x00") + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    return bytes(out)


def _snappy_decompress(src: bytes) -> bytes:
    """Snappy raw format (TF never compresses bundle indices, LevelDB tools may): literals + back-references."""
    n, pos = _get_varint(src, 0)
    out = bytearray()
    while pos < len(src):
        tag = src[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(src[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += src[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = 4 + ((tag >> 2) & 7)
            off = ((tag >> 5) << 8) | src[pos]
            pos += 1
        elif kind == 2:
            ln = 1 + (tag >> 2)
            off = int.from_bytes(src[pos:pos + 2], "little")
            pos += 2
        else:
            ln = 1 + (tag >> 2)
            off = int.from_bytes(src[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise CheckpointError("corrupt snappy block")
        for _ in range(ln):                                     # may overlap its own output
            out.append(out[-off])
    if len(out) != n:
        raise CheckpointError("snappy length mismatch")
    return bytes(out)


def _read_block(data: bytes, offset: int, size: int, verify: bool = True) -> bytes:
    if offset + size + 5 > len(data):
        raise CheckpointError("block handle points past the end of the index file")
    contents = data[offset:offset + size]
    ctype = data[offset + size]
    stored = struct.unpack_from("<I", data, offset + size + 1)[0]
    if verify and unmask_crc(stored) != crc32c(data[offset:offset + size + 1]):
        raise CheckpointError("index block checksum mismatch")
    if ctype == 0:
        return contents
    if ctype == 1:
        return _snappy_decompress(contents)
    raise CheckpointError(f"unknown block compression type {ctype}")


def _iter_block(block: bytes):
    if len(block) < 4:
        raise CheckpointError("block too small")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    if end < 0:
        raise CheckpointError("bad restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > end:
            raise CheckpointError("corrupt block entry")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(data: bytes, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise CheckpointError("not an SSTable (bad magic number): is this a V1 checkpoint or a .data file?")
    footer = data[-48:]
    _, pos = _get_varint(footer, 0)
    _, pos = _get_varint(footer, pos)                           # metaindex handle: unused
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    items: List[Tuple[bytes, bytes]] = []
    for _, h in _iter_block(_read_block(data, ioff, isize, verify)):
        boff, q = _get_varint(h, 0)
        bsize, q = _get_varint(h, q)
        items.extend(_iter_block(_read_block(data, boff, bsize, verify)))
    return items


# ---------------------------------------------------------------------------------------------------------------------
# bundles
# ---------------------------------------------------------------------------------------------------------------------
def _to_numpy(v) -> np.ndarray:
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.asarray(v)


def save(prefix: str, tensors: Dict[str, "np.ndarray"], update_state: bool = True, max_to_keep: int = 10) -> str:
    """writes `<prefix>.index` + `<prefix>.data-00000-of-00001` (what `saver.save(sess, prefix)` leaves on disk) and, when
    `update_state`, the directory's `checkpoint` file (train.py:185 keeps the 10 newest).  Returns prefix."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    names = sorted(tensors, key=lambda s: s.encode())
    if "" in names:
        raise CheckpointError("the empty name is reserved for the bundle header")
    items: List[Tuple[bytes, bytes]] = [(b"", HEADER_BYTES)]
    offset = 0
    tmp = prefix + ".data-00000-of-00001.tmp"
    with open(tmp, "wb") as f:
        for n in names:
            a = _to_numpy(tensors[n])
            if a.dtype not in _NP2DT:
                raise CheckpointError(f"{n}: unsupported dtype {a.dtype}")
            shape = tuple(a.shape)                               # (ascontiguousarray promotes 0-d to 1-d)
            a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<"), copy=False))
            items.append((n.encode(), _encode_entry(_NP2DT[a.dtype], shape, offset, a.nbytes, mask_crc(crc32c(a)))))
            f.write(a.reshape(-1).view(np.uint8).data if a.size else b"")
            offset += a.nbytes
    os.replace(tmp, prefix + ".data-00000-of-00001")
    with open(prefix + ".index.tmp", "wb") as f:
        f.write(build_table(items))
    os.replace(prefix + ".index.tmp", prefix + ".index")
    if update_state:
        update_checkpoint_state(os.path.dirname(os.path.abspath(prefix)), prefix, max_to_keep)
    return prefix


def list_variables(prefix: str) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every tensor in the bundle (tf.train.list_variables)."""
    with open(prefix + ".index", "rb") as f:
        items = read_table(f.read())
    return [(k.decode(), _decode_entry(v)["shape"]) for k, v in items if k != b""]


def load(prefix: str, names: Optional[Iterable[str]] = None, verify: bool = True) -> Dict[str, np.ndarray]:
    """reads a bundle into {variable name: ndarray}; `names` restricts it.  Every tensor's CRC-32C is checked."""
    if not os.path.exists(prefix + ".index"):
        raise CheckpointError(f"{prefix}.index not found (V1 checkpoints and bare .data files are not supported)")
    with open(prefix + ".index", "rb") as f:
        items = read_table(f.read(), verify)
    if not items or items[0][0] != b"":
        raise CheckpointError("bundle header entry missing")
    hdr = _decode_header(items[0][1])
    if hdr["endianness"] != 0:
        raise CheckpointError("big-endian bundles are not supported")
    n_shards = max(hdr["num_shards"], 1)
    want = None if names is None else set(names)
    files: Dict[int, "np.memmap"] = {}
    out: Dict[str, np.ndarray] = {}
    for k, v in items[1:]:
        name = k.decode()
        if want is not None and name not in want:
            continue
        e = _decode_entry(v)
        if e["sliced"]:
            raise CheckpointError(f"{name}: partitioned variables are not supported")
        if e["dtype"] == _DT_BFLOAT16:
            dt, bf16 = np.dtype("<u2"), True
        elif e["dtype"] in _DT:
            dt, bf16 = _DT[e["dtype"]], False
        else:
            raise CheckpointError(f"{name}: unsupported TensorFlow dtype enum {e['dtype']}")
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if count * dt.itemsize != e["size"]:
            raise CheckpointError(f"{name}: size {e['size']} does not match shape {e['shape']} of {dt}")
        sid = e["shard_id"]
        if sid not in files:
            path = f"{prefix}.data-{sid:05d}-of-{n_shards:05d}"
            if not os.path.exists(path):
                raise CheckpointError(f"{path} not found")
            files[sid] = np.memmap(path, dtype=np.uint8, mode="r") if os.path.getsize(path) else np.zeros(0, np.uint8)
        raw = files[sid][e["offset"]:e["offset"] + e["size"]]
        if raw.size != e["size"]:
            raise CheckpointError(f"{name}: data file is truncated")
        raw = np.array(raw)                                     # own copy
        if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise CheckpointError(f"{name}: tensor checksum mismatch")
        a = raw.view(dt).reshape(e["shape"])
        if bf16:
            a = (a.astype(np.uint32) << 16).view(np.float32)
        out[name] = a
    if want is not None and want - set(out):
        raise CheckpointError(f"variables not in the checkpoint: {sorted(want - set(out))[:5]}")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# the `checkpoint` state file (CheckpointState text proto) -- tf.train.get_checkpoint_state / latest_checkpoint
# ---------------------------------------------------------------------------------------------------------------------
def _state_path(directory: str) -> str:
    return os.path.join(directory, "checkpoint")


def get_checkpoint_state(directory: str) -> Optional[dict]:
    """{'model_checkpoint_path': str, 'all_model_checkpoint_paths': [str]} with paths made absolute, or None
    (gen_pred.py:60: `ckpt = tf.train.get_checkpoint_state(model_path)`)."""
    p = _state_path(directory)
    if not os.path.exists(p):
        return None
    cur, allp = None, []
    for line in open(p):
        m = re.match(r'\s*(model_checkpoint_path|all_model_checkpoint_paths)\s*:\s*"((?:[^"\\]|\\.)*)"', line)
        if not m:
            continue
        path = m.group(2).encode().decode("unicode_escape")
        if not os.path.isabs(path):
            path = os.path.join(directory, path)
        if m.group(1) == "model_checkpoint_path":
            cur = path
        else:
            allp.append(path)
    if cur is None:
        return None
    return {"model_checkpoint_path": cur, "all_model_checkpoint_paths": allp}


def latest_checkpoint(directory: str) -> Optional[str]:
    st = get_checkpoint_state(directory)
    if st and os.path.exists(st["model_checkpoint_path"] + ".index"):
        return st["model_checkpoint_path"]
    return None


def update_checkpoint_state(directory: str, prefix: str, max_to_keep: int = 10):
    """appends `prefix` as the newest checkpoint, deletes the files of those beyond `max_to_keep` (Saver(max_to_keep=10))."""
    st = get_checkpoint_state(directory)
    prefix = os.path.abspath(prefix)
    directory = os.path.abspath(directory)
    paths = [os.path.abspath(p) for p in (st["all_model_checkpoint_paths"] if st else [])]
    paths = [p for p in paths if p != prefix] + [prefix]
    if max_to_keep and len(paths) > max_to_keep:
        for old in paths[:-max_to_keep]:
            for suffix in (".index", ".data-00000-of-00001", ".meta"):
                if os.path.exists(old + suffix):
                    os.remove(old + suffix)
        paths = paths[-max_to_keep:]

    def rel(p):
        return os.path.relpath(p, directory) if os.path.dirname(p) == directory else p

    with open(_state_path(directory) + ".tmp", "w") as f:
        f.write(f'model_checkpoint_path: "{rel(prefix)}"\n')
        for p in paths:
            f.write(f'all_model_checkpoint_paths: "{rel(p)}"\n')
    os.replace(_state_path(directory) + ".tmp", _state_path(directory))


# ---------------------------------------------------------------------------------------------------------------------
# Adam slot naming of tf.train.AdamOptimizer (train.py:168), for full-state resume
# ---------------------------------------------------------------------------------------------------------------------
def adam_slot_names(var_name: str) -> Tuple[str, str]:
    return var_name + "/Adam", var_name + "/Adam_1"
