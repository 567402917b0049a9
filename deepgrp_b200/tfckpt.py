"""Reader (and a minimal writer, for tests) of TensorFlow's checkpoint format -- the "tensor bundle" that
``keras.Model.save_weights`` / ``ModelCheckpoint(save_weights_only=True)`` write when the path has no
``.h5`` suffix, which is what the reference's training does (``deepgrp/training.py:53-59``) and what
``deepgrp.prediction.setup_prediction_from_options_checkpoint`` restores (``deepgrp/prediction.py:68-86``).

A bundle is ``<prefix>.index`` -- an SSTable in the LevelDB table format (blocks of prefix-compressed
key/value entries with restart points, an index block, a 48-byte footer with the magic
0xdb4775248b80fb57; blocks may be Snappy-compressed) whose values are ``BundleEntryProto`` messages
(dtype, shape, shard_id, offset, size, crc32c) -- plus ``<prefix>.data-XXXXX-of-YYYYY`` shards holding
the raw little-endian tensor bytes.  The directory's ``checkpoint`` text file names the latest prefix.

PARITY UNPINNED: neither TensorFlow nor a checkpoint written by it is available in this environment
(the reference ships none); the format is restated from the TensorFlow / LevelDB sources
(tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table, tensor_bundle.proto) and is
exercised only against bundles produced by :func:`write_bundle` below and hand-made Snappy streams;
what IS pinned by published values: CRC-32C (RFC 3720 check vectors), its LevelDB mask, the table magic.
Block trailers and per-tensor checksums are verified on read, so a misparsed offset fails loudly.
No GPU work happens here: this is host-side file parsing, like ``hdf5.py``.
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 19: np.float16}
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---- varints / protobuf wire format --------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


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


def _proto_fields(buf: bytes):
    """Yield (field number, wire type, value) of one protobuf message (no groups)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf: bytes) -> List[int]:
    dims = []
    for field, wt, v in _proto_fields(buf):
        if field == 2 and wt == 2:                      # repeated Dim dim = 2
            size = 0
            for f2, w2, v2 in _proto_fields(v):
                if f2 == 1 and w2 == 0:                 # int64 size = 1
                    size = _signed64(v2)
            dims.append(size)
    return dims


def _parse_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for field, wt, v in _proto_fields(buf):
        if field == 1 and wt == 0:
            e["dtype"] = v
        elif field == 2 and wt == 2:
            e["shape"] = _parse_shape(v)
        elif field == 3 and wt == 0:
            e["shard_id"] = v
        elif field == 4 and wt == 0:
            e["offset"] = v
        elif field == 5 and wt == 0:
            e["size"] = v
        elif field == 6 and wt == 5:
            e["crc32c"] = struct.unpack("<I", v)[0]
        elif field == 7:
            e["sliced"] = True
    return e


# ---- crc32c ---------------------------------------------------------------------------------------
def _crc32c_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_CRC_TABLE = _crc32c_table()


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli, reflected polynomial 0x82F63B78); crc32c(b"123456789") == 0xE3069283."""
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked_crc(data: bytes) -> int:
    """The masked form both LevelDB block trailers and BundleEntryProto.crc32c store:
    rotate right by 15, add 0xa282ead8."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- Snappy (raw format) -------------------------------------------------------------------------
def snappy_decompress(buf: bytes) -> bytes:
    """Raw Snappy block: varint uncompressed length, then literal (tag & 3 == 0) and copy elements."""
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
            ln = 4 + ((tag >> 2) & 7)
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("malformed Snappy copy")
        for _ in range(ln):                              # may overlap its own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("Snappy length mismatch: %d != %d" % (len(out), n))
    return bytes(out)


# ---- LevelDB table -------------------------------------------------------------------------------
def _read_block(data: bytes, offset: int, size: int) -> bytes:
    block = data[offset:offset + size]
    ctype = data[offset + size]                          # 1-byte type, then a 4-byte masked crc32c
    stored = struct.unpack("<I", data[offset + size + 1:offset + size + 5])[0]
    if stored != _masked_crc(data[offset:offset + size + 1]):
        raise ValueError("table block at %d: checksum mismatch" % offset)
    if ctype == 0:
        return block
    if ctype == 1:
        return snappy_decompress(block)
    raise ValueError("unknown block compression type %d" % ctype)


def _block_entries(block: bytes):
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path: str) -> Dict[bytes, bytes]:
    """All key/value pairs of an SSTable file."""
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise ValueError("%s is not a LevelDB-format table (bad magic)" % path)
    footer = data[-48:]
    _, pos = _varint(footer, 0)                          # metaindex handle (unused)
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    out: Dict[bytes, bytes] = {}
    for _, handle in _block_entries(_read_block(data, ioff, isize)):
        boff, p2 = _varint(handle, 0)
        bsize, _ = _varint(handle, p2)
        for k, v in _block_entries(_read_block(data, boff, bsize)):
            out[k] = v
    return out


# ---- bundle --------------------------------------------------------------------------------------
def read_bundle(prefix: str, verify: bool = True, select=None) -> Dict[str, np.ndarray]:
    """name -> array for every numeric, unsliced tensor of the bundle ``prefix``(.index / .data-*).
    ``verify`` checks each tensor's stored crc32c (as TensorFlow's BundleReader does); ``select`` is an
    optional predicate on the key (str) to skip tensors, e.g. optimizer slots."""
    table = read_table(prefix + ".index")
    num_shards = 1
    if b"" in table:
        for field, wt, v in _proto_fields(table[b""]):   # BundleHeaderProto
            if field == 1 and wt == 0:
                num_shards = v
            elif field == 2 and wt == 0 and v != 0:
                raise ValueError("big-endian bundles are not supported")
    shards: Dict[int, bytes] = {}
    out: Dict[str, np.ndarray] = {}
    for key, val in table.items():
        if key == b"":
            continue
        if select is not None and not select(key.decode("utf-8", "replace")):
            continue
        e = _parse_entry(val)
        if e["sliced"] or e["dtype"] not in _DTYPES:
            continue                                     # strings (object graph), partitioned variables
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), "rb").read()
        dt = np.dtype(_DTYPES[e["dtype"]])
        count = int(np.prod(e["shape"])) if e["shape"] else 1
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if len(raw) != count * dt.itemsize:
            raise ValueError("tensor %r: %d bytes for shape %s" % (key, len(raw), e["shape"]))
        if verify and e["crc32c"] is not None and e["crc32c"] != _masked_crc(raw):
            raise ValueError("tensor %r: checksum mismatch" % key)
        out[key.decode("utf-8", "replace")] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    return out


def latest_checkpoint(logdir: str) -> Optional[str]:
    """The prefix ``tf.train.CheckpointManager(...).latest_checkpoint`` would return: the ``checkpoint``
    state file's ``model_checkpoint_path``, else the newest ``*.index`` in the directory."""
    state = os.path.join(logdir, "checkpoint")
    if os.path.exists(state):
        m = re.search(r'^model_checkpoint_path:\s*"(.*)"', open(state).read(), re.M)
        if m:
            p = m.group(1)
            p = p if os.path.isabs(p) else os.path.join(logdir, p)
            if os.path.exists(p + ".index"):
                return p
    cands = [os.path.join(logdir, f[:-6]) for f in os.listdir(logdir) if f.endswith(".index")]
    return max(cands, key=lambda p: os.path.getmtime(p + ".index")) if cands else None


def deepgrp_weights(tensors: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Pick the DeepGRP model's variables out of a ``save_weights`` bundle (object-graph keys such as
    ``layer_with_weights-0/cell/kernel/.ATTRIBUTES/VARIABLE_VALUE``; optimizer slots are ignored) and
    return them under the names ``deepgrp_b200.model.ModelWeights`` uses.  Variables are recognised by
    key suffix and confirmed by shape: kernel [5, G U], recurrent_kernel [U, G U], bias [2, 3U] (GRU,
    reset_after) or [4U] (LSTM), attention scale [U], FF kernel [F, C], FF bias [C]."""
    vars_ = {k: v for k, v in tensors.items()
             if "OPTIMIZER_SLOT" not in k and not k.startswith("optimizer") and k.endswith("VARIABLE_VALUE")}

    def pick(pattern, pred):
        hits = [(k, v) for k, v in vars_.items() if re.search(pattern, k) and pred(v)]
        if len(hits) != 1:
            raise ValueError("checkpoint: expected exactly one variable matching %s, found %d (%s)"
                             % (pattern, len(hits), [k for k, _ in hits]))
        return hits[0][1].astype(np.float32)

    rk = pick(r"cell/recurrent_kernel/", lambda a: a.ndim == 2 and a.shape[1] % a.shape[0] == 0)
    units, gates = rk.shape[0], rk.shape[1] // rk.shape[0]
    out = {"recurrent_kernel": rk,
           "kernel": pick(r"cell/kernel/", lambda a: a.shape == (5, gates * units)),
           "bias": pick(r"cell/bias/", lambda a: a.shape in ((2, gates * units), (gates * units,)))}
    att = [v for k, v in vars_.items() if re.search(r"/scale/", k) and v.shape == (units,)]
    if att:
        out["att_scale"] = att[0].astype(np.float32)
    feat = 2 * units if att else units
    out["ff_kernel"] = pick(r"(?<!cell)/kernel/", lambda a: a.ndim == 2 and a.shape[0] == feat)
    out["ff_bias"] = pick(r"(?<!cell)/bias/", lambda a: a.ndim == 1 and a.shape[0] == out["ff_kernel"].shape[1])
    out["rnn"] = "LSTM" if gates == 4 else "GRU"
    return out


# ---- writer (tests only: produces what read_bundle expects) ---------------------------------------
def _build_block(items, restart_interval: int = 16) -> bytes:
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(last), len(k)) and last[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], block_items: int = 4) -> None:
    """Write ``prefix``.index / .data-00000-of-00001 (uncompressed blocks) and a ``checkpoint`` state file
    next to them.  For the tests of this module; TensorFlow is what writes real checkpoints."""
    data, entries = bytearray(), []
    for name in sorted(tensors):
        a = np.asarray(tensors[name])            # (ascontiguousarray would turn a scalar into shape [1])
        raw = a.tobytes(order="C")
        shape = b"".join(b"\x12" + _put_varint(len(d)) + d
                         for d in (b"\x08" + _put_varint(int(s)) for s in a.shape))
        msg = (b"\x08" + _put_varint(_DTYPE_CODES[a.dtype]) + b"\x12" + _put_varint(len(shape)) + shape +
               b"\x20" + _put_varint(len(data)) + b"\x28" + _put_varint(len(raw)) +
               b"\x35" + struct.pack("<I", _masked_crc(raw)))
        entries.append((name.encode(), msg))
        data += raw
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"          # num_shards = 1, version { producer = 1 }
    items = [(b"", header)] + entries
    out, index = bytearray(), []
    for i in range(0, len(items), block_items):
        chunk = items[i:i + block_items]
        block = _build_block(chunk)
        index.append((chunk[-1][0], _put_varint(len(out)) + _put_varint(len(block))))
        out += block + b"\x00" + struct.pack("<I", _masked_crc(block + b"\x00"))
    meta = _build_block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta))
    out += meta + b"\x00" + struct.pack("<I", _masked_crc(meta + b"\x00"))
    iblock = _build_block(index, restart_interval=1)
    index_handle = _put_varint(len(out)) + _put_varint(len(iblock))
    out += iblock + b"\x00" + struct.pack("<I", _masked_crc(iblock + b"\x00"))
    footer = meta_handle + index_handle
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    open(prefix + ".index", "wb").write(bytes(out))
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with open(os.path.join(os.path.dirname(prefix) or ".", "checkpoint"), "w") as fh:
        base = os.path.basename(prefix)
        fh.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
