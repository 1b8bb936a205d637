"""Pure-Python reader (and minimal writer) of TensorFlow's V2 checkpoint format, the "tensor bundle":
`<prefix>.index` + `<prefix>.data-00000-of-0000N`.

Why it exists: the two ImageBert scorers restore TF-1.12 checkpoints --
`saver.restore(sess, ckpt)` over `variable_averages.variables_to_restore()` (imagebert_zk/evaluate_normal.py:204-212)
and `tf.train.init_from_checkpoint` / `tf.train.list_variables` (imagebert_lds/src/run_pretraining_predict_score.py:
347-362, 558-563) -- and TensorFlow is installable neither here nor on the GPU box.  With this module a competition
checkpoint goes straight into `checkpoints.select_tf_variables` and from there into `mmr_create`, no TF involved.

The format is TensorFlow's, not the reference's; it is restated from the published sources of the pinned dependency
tensorflow(-gpu)==1.12.0 (requirements.txt:68,70):
  * tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc} and tensorflow/core/protobuf/tensor_bundle.proto:
    the index maps "" -> BundleHeaderProto and every tensor name -> BundleEntryProto {dtype = 1, shape = 2,
    shard_id = 3, offset = 4, size = 5, crc32c = 6 (fixed32, masked), slices = 7}; tensor bytes sit raw,
    little-endian, at [offset, offset + size) of data shard `shard_id`.
  * tensorflow/core/lib/io/{format,block,table}.cc (the LevelDB table format): blocks of prefix-compressed entries
    (varint32 shared, non_shared, value_len; key tail; value) with a restart array, each followed by a 1-byte
    compression type (0 none, 1 snappy) and a masked crc32c; a 48-byte footer = metaindex handle, index handle,
    padding, magic 0xdb4775248b80fb57.
No TF-written file is available offline, so the reader is pinned by round trips through the writer below, by
hand-assembled blocks (prefix compression, restarts, snappy) in tests/test_tf_bundle.py, and by the CRCs the format
carries (index blocks are always verified; tensor data on request -- it is pure-Python speed).
Partitioned variables (entries with `slices`) are refused: BERT checkpoints do not use them.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
FOOTER_LEN = 48
BLOCK_TRAILER = 5

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
           6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 19: np.dtype("<f2"),
           22: np.dtype("<u4"), 23: np.dtype("<u8")}
_DT_BFLOAT16 = 14
_DT_OF = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("int64"): 9,
          np.dtype("float16"): 19, np.dtype("uint8"): 4, np.dtype("bool"): 10}


class BundleError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------- crc32c (Castagnoli)
def _make_crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _make_crc_table()


_native_crc = None


def _native():
    """mmr_crc32c of libmmrecall.so (host code, hardware crc32 instruction) when the library is built; else None."""
    global _native_crc
    if _native_crc is None:
        try:
            from . import _lib
            _native_crc = _lib.load().mmr_crc32c
        except Exception:
            _native_crc = False
    return _native_crc or None


def crc32c(data, crc: int = 0) -> int:
    fn = _native()
    if fn is not None:
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8)
        return int(fn(buf.ctypes.data, buf.size, crc))
    return crc32c_py(bytes(data), crc)


def crc32c_py(data: bytes, crc: int = 0) -> int:
    """Reference implementation (byte-wise table); ~1 us per byte, used only when the library is not built."""
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    """tensorflow/core/lib/hash/crc32c.h: Mask()."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- varints / protobuf
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise BundleError("truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 63:
            raise BundleError("varint too long")


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf: bytes) -> Iterable[Tuple[int, int, object]]:
    """(field number, wire type, value) of one message: varint -> int, 64-bit / 32-bit -> raw bytes, length-delimited
    -> bytes."""
    pos = 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise BundleError(f"unsupported protobuf wire type {wt}")
        yield field, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for field, wt, v in _proto_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:                                   # TensorShapeProto
            for f2, _, v2 in _proto_fields(v):
                if f2 == 2:                                # Dim
                    size = 0
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = _signed64(v3)
                    e["shape"].append(size)
                elif f2 == 3 and v2:
                    raise BundleError("tensor of unknown rank in a checkpoint")
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
        elif field == 7:
            e["sliced"] = True
    return e


def _parse_header(buf: bytes) -> dict:
    h = {"num_shards": 1, "endianness": 0}
    for field, _, v in _proto_fields(buf):
        if field == 1:
            h["num_shards"] = v
        elif field == 2:
            h["endianness"] = v
    return h


# ---------------------------------------------------------------------------------------------- snappy (raw format)
def snappy_decompress(buf: bytes) -> bytes:
    """Raw snappy block (format_description.txt of google/snappy): varint uncompressed length, then literal / copy
    elements.  TF writes bundle indices uncompressed; kept so that a snappy-compressed table still reads."""
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise BundleError("corrupt snappy copy offset")
        for _ in range(ln):                                # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise BundleError("snappy length mismatch")
    return bytes(out)


# ---------------------------------------------------------------------------------------------- table reader
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    if offset + size + BLOCK_TRAILER > len(data):
        raise BundleError("block handle points outside the file")
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        want = struct.unpack("<I", data[offset + size + 1:offset + size + 5])[0]
        if mask_crc(crc32c(data[offset:offset + size + 1])) != want:
            raise BundleError(f"crc mismatch in index block at offset {offset}")
    if ctype == 0:
        return raw
    if ctype == 1:
        return snappy_decompress(raw)
    raise BundleError(f"unknown block compression type {ctype}")


def _block_entries(block: bytes) -> Iterable[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise BundleError("block too small")
    n_restarts = struct.unpack("<I", block[-4:])[0]
    limit = len(block) - 4 - 4 * n_restarts
    if limit < 0:
        raise BundleError("corrupt restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise BundleError("corrupt block entry")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path: str, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < FOOTER_LEN:
        raise BundleError(f"{path}: too short for a table footer")
    footer = data[-FOOTER_LEN:]
    if struct.unpack("<Q", footer[-8:])[0] != TABLE_MAGIC:
        raise BundleError(f"{path}: bad table magic (not a TF checkpoint index)")
    pos = 0
    _, pos = _varint(footer, pos)          # metaindex handle (unused by TF bundles)
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p = _varint(handle, 0)
        size, _ = _varint(handle, p)
        out.extend(_block_entries(_read_block(data, off, size, verify)))
    return out


# ---------------------------------------------------------------------------------------------- bundle reader
class BundleReader:
    """tf.train.load_checkpoint / tf.train.list_variables without TensorFlow."""

    def __init__(self, prefix: str, verify_data: bool = False):
        self.prefix = prefix
        self.verify_data = verify_data
        index = prefix + ".index"
        if not os.path.exists(index):
            raise FileNotFoundError(f"{index}: no such checkpoint (pass the prefix, e.g. .../model.ckpt-251)")
        self.header, self.entries = None, {}
        for key, value in read_table(index):
            if key == b"":
                self.header = _parse_header(value)
            else:
                self.entries[key.decode("utf-8")] = _parse_entry(value)
        if self.header is None:
            raise BundleError(f"{index}: no bundle header entry")
        if self.header["endianness"] != 0:
            raise BundleError("big-endian checkpoint")
        self._shards: Dict[int, np.memmap] = {}

    def list_variables(self) -> List[Tuple[str, List[int]]]:
        """[(name, shape)] like tf.train.list_variables (run_pretraining_predict_score.py:352)."""
        return [(n, list(e["shape"])) for n, e in sorted(self.entries.items())]

    def _shard(self, sid: int) -> np.memmap:
        if sid not in self._shards:
            path = f"{self.prefix}.data-{sid:05d}-of-{self.header['num_shards']:05d}"
            self._shards[sid] = np.memmap(path, dtype=np.uint8, mode="r")
        return self._shards[sid]

    def get_tensor(self, name: str) -> np.ndarray:
        e = self.entries.get(name)
        if e is None:
            raise KeyError(f"'{name}' is not in checkpoint {self.prefix}")
        if e["sliced"]:
            raise BundleError(f"'{name}' is a partitioned variable; not supported")
        shard = self._shard(e["shard_id"])
        if e["offset"] + e["size"] > shard.shape[0]:
            raise BundleError(f"'{name}': data range outside shard {e['shard_id']}")
        raw = np.asarray(shard[e["offset"]:e["offset"] + e["size"]])
        if self.verify_data and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
            raise BundleError(f"'{name}': data crc mismatch")
        if e["dtype"] == _DT_BFLOAT16:
            arr = (raw.view("<u2").astype(np.uint32) << 16).view(np.float32)
        else:
            dt = _DTYPES.get(e["dtype"])
            if dt is None:
                raise BundleError(f"'{name}': unsupported dtype enum {e['dtype']}")
            arr = raw.view(dt)
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        if arr.size != n:
            raise BundleError(f"'{name}': {arr.size} elements on disk, shape {e['shape']} wants {n}")
        return np.array(arr).reshape(e["shape"])

    def read_all(self, float_only: bool = True) -> Dict[str, np.ndarray]:
        out = {}
        for name, e in self.entries.items():
            if e["sliced"]:
                continue
            if float_only and e["dtype"] not in (1, 2, 19, _DT_BFLOAT16):
                continue                                   # global_step and friends
            out[name] = self.get_tensor(name)
        return out


def latest_checkpoint(directory: str) -> Optional[str]:
    """tf.train.latest_checkpoint: the prefix named by the `checkpoint` state file (evaluate_normal.py:207-208)."""
    state = os.path.join(directory, "checkpoint")
    if not os.path.exists(state):
        return None
    for line in open(state):
        if line.startswith("model_checkpoint_path:"):
            p = line.split(":", 1)[1].strip().strip('"')
            return p if os.path.isabs(p) else os.path.join(directory, p)
    return None


# ---------------------------------------------------------------------------------------------- writer
def _entry_bytes(dtype_enum: int, shape, shard: int, offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype_enum)
    out += b"\x12" + _put_varint(len(dims)) + dims
    if shard:
        out += b"\x18" + _put_varint(shard)
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size)
    out += b"\x35" + struct.pack("<I", crc)
    return out


def _build_block(items: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(prev)) and k[shared] == prev[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts.append(0)
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], block_entries: int = 64) -> None:
    """Writes `{name: array}` as a one-shard V2 checkpoint that BundleReader -- and TensorFlow -- read: data file,
    then an uncompressed index table (data blocks of `block_entries` entries, an index block, an empty metaindex).
    Used by the tests and as the export side of a round trip; float64 / int arrays keep their dtype."""
    names = sorted(tensors)
    entries, offset = [], 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for n in names:
            a = np.asarray(tensors[n])
            if a.dtype not in _DT_OF:
                a = a.astype(np.float32)
            shape = a.shape                                 # (np.ascontiguousarray would turn a scalar into [1])
            raw = np.ascontiguousarray(a).astype(a.dtype.newbyteorder("<")).tobytes()
            f.write(raw)
            entries.append((n.encode("utf-8"), _entry_bytes(_DT_OF[a.dtype], shape, 0, offset, len(raw),
                                                             mask_crc(crc32c(raw)))))
            offset += len(raw)
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"             # num_shards = 1, little endian (default), version.producer = 1
    items = [(b"", header)] + entries
    body, index_items = bytearray(), []

    def emit(block: bytes) -> bytes:
        off = len(body)
        body.extend(block)
        body.extend(b"\x00")
        body.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    for i in range(0, len(items), block_entries):
        chunk = items[i:i + block_entries]
        handle = emit(_build_block(chunk))
        index_items.append((chunk[-1][0], handle))                # a separator >= the block's last key: the key itself
    meta = emit(_build_block([]))
    index = emit(_build_block(index_items, restart_interval=1))
    footer = meta + index
    footer += b"\x00" * (FOOTER_LEN - 8 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(body) + footer)
