"""Minimal pure-Python HDF5 reader/writer for Keras model files (no h5py / libhdf5 in this image).

The reference loads its model with ``tf.keras.models.load_model(args.model)``
(``deepgrp/__main__.py:264-269``); the deployable artefact is a Keras HDF5 file written by h5py 3.1 with
``libver='earliest'``.  This module reads the subset of HDF5 such files use:

* superblock version 0/1, 8-byte offsets and lengths;
* version-1 object headers with continuation blocks;
* "old style" groups: symbol-table message -> B-tree v1 (node type 0) -> symbol-table nodes + local heap;
* datasets: dataspace v1/v2, little-endian IEEE float / fixed-point / fixed-length string datatypes,
  contiguous or compact layout (no chunking, no filters -- Keras weights are plain contiguous arrays);
* attributes (message versions 1-3) including variable-length strings through the global heap
  (``model_config``) and arrays of fixed-length strings (``layer_names`` / ``weight_names``).

Stated from the published HDF5 file-format specification; the reference ships no ``.h5``/``.hdf5`` file,
so reading a file written by real Keras is **unpinned** (SURVEY.md section 7 H6).  The writer emits the same
subset and is what the tests round-trip against; ``.npz`` is offered as a side format
(``ModelWeights.save_npz``).
"""
from __future__ import annotations

import json
import struct
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5Error(ValueError):
    pass


# =================================================================================================
# Reader
# =================================================================================================
class _Datatype:
    def __init__(self, cls: int, size: int, np_dtype=None, vlen_string: bool = False):
        self.cls, self.size, self.np_dtype, self.vlen_string = cls, size, np_dtype, vlen_string


class H5Object:
    """A group or a dataset: ``attrs`` dict, ``keys()``/``[]`` for groups, ``value`` for datasets."""

    def __init__(self, f: "H5File", addr: int):
        self._f = f
        self.addr = addr
        self.attrs: Dict[str, Any] = {}
        self._links: Optional[Dict[str, int]] = None
        self._shape: Optional[Tuple[int, ...]] = None
        self._dtype: Optional[_Datatype] = None
        self._layout: Optional[tuple] = None
        self._btree = self._heap = None
        self._parse_header()

    # ---- object header ---------------------------------------------------------------------------
    def _parse_header(self) -> None:
        b = self._f.buf
        ver = b[self.addr]
        if ver != 1:
            raise HDF5Error("only version-1 object headers are supported (found %d)" % ver)
        n_msgs, = struct.unpack_from("<H", b, self.addr + 2)
        hdr_size, = struct.unpack_from("<I", b, self.addr + 8)
        blocks = [(self.addr + 16, hdr_size)]
        seen = 0
        while blocks and seen < n_msgs:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and seen < n_msgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                data = pos + 8
                self._message(mtype, data, msize, blocks)
                pos = data + msize
                seen += 1

    def _message(self, mtype: int, p: int, size: int, blocks: list) -> None:
        b = self._f.buf
        if mtype == 0x0010:                                   # continuation
            off, length = struct.unpack_from("<QQ", b, p)
            blocks.append((off, length))
        elif mtype == 0x0011:                                 # symbol table (old-style group)
            self._btree, self._heap = struct.unpack_from("<QQ", b, p)
        elif mtype == 0x0001:
            self._shape = self._f._dec_dataspace(p)
        elif mtype == 0x0003:
            self._dtype = self._f._dec_datatype(p)
        elif mtype == 0x0008:
            self._layout = self._f._dec_layout(p)
        elif mtype == 0x000B:
            raise HDF5Error("filtered (compressed) datasets are not supported")
        elif mtype == 0x000C:
            name, value = self._f._dec_attribute(p)
            self.attrs[name] = value

    # ---- groups -----------------------------------------------------------------------------------
    @property
    def is_group(self) -> bool:
        return self._btree is not None

    def _load_links(self) -> Dict[str, int]:
        if self._links is None:
            self._links = {}
            if self._btree is not None:
                heap_data = self._f._local_heap(self._heap)
                self._f._walk_group_btree(self._btree, heap_data, self._links)
        return self._links

    def keys(self) -> List[str]:
        return list(self._load_links())

    def __contains__(self, name: str) -> bool:
        return name.split("/")[0] in self._load_links()

    def __getitem__(self, path: str) -> "H5Object":
        obj = self
        for part in [p for p in path.split("/") if p]:
            links = obj._load_links()
            if part not in links:
                raise KeyError(path)
            obj = H5Object(self._f, links[part])
        return obj

    # ---- datasets ---------------------------------------------------------------------------------
    @property
    def shape(self):
        return self._shape

    @property
    def value(self) -> np.ndarray:
        if self._dtype is None or self._layout is None or self._shape is None:
            raise HDF5Error("object is not a dataset")
        return self._f._read_data(self._dtype, self._shape, self._layout)


class H5File(H5Object):
    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            self.buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as fh:
                self.buf = fh.read()
        base = self.buf.find(SIGNATURE)
        if base != 0:
            raise HDF5Error("not an HDF5 file (signature missing at offset 0)")
        ver = self.buf[8]
        if ver not in (0, 1):
            raise HDF5Error("superblock version %d is not supported (Keras/h5py 'earliest' writes 0)" % ver)
        if self.buf[13] != 8 or self.buf[14] != 8:
            raise HDF5Error("only 8-byte offsets/lengths are supported")
        p = 24 + (4 if ver == 1 else 0)
        p += 32                                   # base, free-space, end-of-file, driver addresses
        _name_off, root_addr, cache_type = struct.unpack_from("<QQI", self.buf, p)
        super().__init__(self, root_addr)
        if self._btree is None and cache_type == 1:
            self._btree, self._heap = struct.unpack_from("<QQ", self.buf, p + 24)

    # ---- low-level decoders ------------------------------------------------------------------------
    def _local_heap(self, addr: int) -> bytes:
        b = self.buf
        if b[addr:addr + 4] != b"HEAP":
            raise HDF5Error("bad local heap signature")
        size, _free, data_addr = struct.unpack_from("<QQQ", b, addr + 8)
        return b[data_addr:data_addr + size]

    def _walk_group_btree(self, addr: int, heap: bytes, out: Dict[str, int]) -> None:
        b = self.buf
        if b[addr:addr + 4] == b"SNOD":
            n, = struct.unpack_from("<H", b, addr + 6)
            for i in range(n):
                e = addr + 8 + i * 40
                name_off, obj_addr = struct.unpack_from("<QQ", b, e)
                end = heap.index(b"\x00", name_off)
                out[heap[name_off:end].decode("utf-8")] = obj_addr
            return
        if b[addr:addr + 4] != b"TREE":
            raise HDF5Error("bad B-tree signature")
        _ntype, _level, used = struct.unpack_from("<BBH", b, addr + 4)
        p = addr + 8 + 16 + 8                      # skip siblings and key 0
        for _ in range(used):
            child, = struct.unpack_from("<Q", b, p)
            self._walk_group_btree(child, heap, out)
            p += 16                                # child + next key

    def _dec_dataspace(self, p: int) -> Tuple[int, ...]:
        b = self.buf
        ver, rank, flags = b[p], b[p + 1], b[p + 2]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            if b[p + 3] == 2:                      # null dataspace
                return (0,)
            q = p + 4
        else:
            raise HDF5Error("dataspace version %d" % ver)
        return tuple(struct.unpack_from("<%dQ" % rank, b, q)) if rank else ()

    def _dec_datatype(self, p: int) -> _Datatype:
        b = self.buf
        cls, ver = b[p] & 0x0F, b[p] >> 4
        bits0 = b[p + 1]
        size, = struct.unpack_from("<I", b, p + 4)
        if cls == 1:                               # floating point
            if bits0 & 1:
                raise HDF5Error("big-endian floats are not supported")
            return _Datatype(cls, size, np.dtype("<f%d" % size))
        if cls == 0:                               # fixed point
            signed = bool(bits0 & 0x08)
            return _Datatype(cls, size, np.dtype("<%s%d" % ("i" if signed else "u", size)))
        if cls == 3:                               # fixed-length string
            return _Datatype(cls, size, np.dtype("S%d" % size))
        if cls == 9:                               # variable length
            if (bits0 & 0x0F) != 1:
                raise HDF5Error("variable-length sequences are not supported (only strings)")
            return _Datatype(cls, size, None, vlen_string=True)
        raise HDF5Error("datatype class %d (version %d) is not supported" % (cls, ver))

    def _dec_layout(self, p: int) -> tuple:
        b = self.buf
        ver, cls = b[p], b[p + 1]
        if ver != 3:
            raise HDF5Error("data layout message version %d is not supported" % ver)
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", b, p + 2)
            return ("contiguous", addr, size)
        if cls == 0:
            size, = struct.unpack_from("<H", b, p + 2)
            return ("compact", p + 4, size)
        raise HDF5Error("chunked datasets are not supported (Keras writes contiguous weights)")

    def _global_heap_object(self, addr: int, index: int) -> bytes:
        b = self.buf
        if b[addr:addr + 4] != b"GCOL":
            raise HDF5Error("bad global heap signature")
        size, = struct.unpack_from("<Q", b, addr + 8)
        p, end = addr + 16, addr + size
        while p + 16 <= end:
            idx, _ref, _res, osize = struct.unpack_from("<HHIQ", b, p)
            if idx == 0:
                break
            if idx == index:
                return b[p + 16:p + 16 + osize]
            p += 16 + ((osize + 7) // 8) * 8
        raise HDF5Error("global heap object %d not found" % index)

    def _read_raw(self, dt: _Datatype, shape: Tuple[int, ...], raw: bytes):
        n = int(np.prod(shape)) if shape else 1
        if dt.vlen_string:
            vals = []
            for i in range(n):
                _length, gaddr, gidx = struct.unpack_from("<IQI", raw, i * 16)
                vals.append(self._global_heap_object(gaddr, gidx).decode("utf-8"))
            return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        arr = np.frombuffer(raw[:n * dt.size], dtype=dt.np_dtype)
        if dt.cls == 3:
            arr = np.char.rstrip(arr, b"\x00")
        return arr.reshape(shape) if shape else arr[0]

    def _read_data(self, dt: _Datatype, shape, layout) -> np.ndarray:
        _kind, addr, size = layout
        if addr == UNDEF:
            return np.zeros(shape, dtype=dt.np_dtype)
        return self._read_raw(dt, shape, self.buf[addr:addr + size])

    def _dec_attribute(self, p: int):
        b = self.buf
        ver = b[p]
        name_size, dt_size, ds_size = struct.unpack_from("<HHH", b, p + 2)
        q = p + 8
        if ver == 3:
            q += 1
        pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
        name = b[q:q + name_size].split(b"\x00")[0].decode("utf-8")
        q += pad(name_size)
        dt = self._dec_datatype(q)
        q += pad(dt_size)
        shape = self._dec_dataspace(q)
        q += pad(ds_size)
        n = int(np.prod(shape)) if shape else 1
        return name, self._read_raw(dt, shape, b[q:q + n * dt.size])


# =================================================================================================
# Writer (same subset)
# =================================================================================================
class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def tell(self) -> int:
        return len(self.buf)

    def align(self, n: int = 8) -> None:
        while len(self.buf) % n:
            self.buf.append(0)

    def put(self, data: bytes) -> int:
        self.align()
        at = len(self.buf)
        self.buf += data
        return at


def _dt_float32() -> bytes:
    return bytes([0x11, 0x20, 0x1F, 0x00]) + struct.pack("<I", 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def _dt_string(size: int) -> bytes:
    return bytes([0x13, 0x00, 0x00, 0x00]) + struct.pack("<I", size)


def _dt_vlen_string() -> bytes:
    base = _dt_string(1)
    return bytes([0x19, 0x01, 0x01, 0x00]) + struct.pack("<I", 16) + base     # type=string, charset utf-8


def _dataspace(shape: Tuple[int, ...]) -> bytes:
    return bytes([1, len(shape), 0, 0, 0, 0, 0, 0]) + b"".join(struct.pack("<Q", d) for d in shape)


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _attr_message(name: str, dt: bytes, ds: bytes, data: bytes) -> bytes:
    nm = name.encode("utf-8") + b"\x00"
    body = bytes([1, 0]) + struct.pack("<HHH", len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + data
    return body


def _message(mtype: int, body: bytes) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _object_header(messages: List[bytes]) -> bytes:
    payload = b"".join(messages)
    return bytes([1, 0]) + struct.pack("<HII", len(messages), 1, len(payload)) + b"\x00" * 4 + payload


class H5Builder:
    """Builds a file of nested old-style groups, contiguous float32 datasets and string attributes."""

    def __init__(self):
        self.w = _Writer()
        self.w.buf += b"\x00" * 96                     # superblock placeholder
        self._gcol: List[bytes] = []

    def _vlen_refs(self, strings: List[str]) -> Tuple[int, List[int]]:
        """Write one global heap collection holding `strings`; returns (address, object indices)."""
        objs = b""
        idxs = []
        for i, s in enumerate(strings, start=1):
            data = s.encode("utf-8")
            objs += struct.pack("<HHIQ", i, 1, 0, len(data)) + _pad8(data)
            idxs.append(i)
        free = struct.pack("<HHIQ", 0, 0, 0, 0)
        size = 16 + len(objs) + len(free)
        size = max(4096, (size + 7) // 8 * 8)
        body = b"GCOL" + bytes([1, 0, 0, 0]) + struct.pack("<Q", size) + objs
        body += struct.pack("<HHIQ", 0, 0, 0, size - len(body) - 16)
        body += b"\x00" * (size - len(body))
        return self.w.put(body), idxs

    def attr_messages(self, attrs: Dict[str, Any]) -> List[bytes]:
        msgs = []
        for name, value in attrs.items():
            if isinstance(value, str):                              # scalar variable-length string
                addr, idxs = self._vlen_refs([value])
                data = struct.pack("<IQI", len(value.encode("utf-8")), addr, idxs[0])
                msgs.append(_message(0x000C, _attr_message(name, _dt_vlen_string(), _dataspace(()), data)))
            elif isinstance(value, (list, tuple)) and all(isinstance(v, (bytes, str)) for v in value):
                items = [v.encode("utf-8") if isinstance(v, str) else v for v in value]
                width = max([len(v) for v in items] + [1])
                data = b"".join(v.ljust(width, b"\x00") for v in items)
                msgs.append(_message(0x000C, _attr_message(name, _dt_string(width), _dataspace((len(items),)), data)))
            else:
                arr = np.ascontiguousarray(value, dtype="<f4")
                msgs.append(_message(0x000C, _attr_message(name, _dt_float32(), _dataspace(arr.shape), arr.tobytes())))
        return msgs

    def dataset(self, array: np.ndarray) -> int:
        arr = np.ascontiguousarray(array, dtype="<f4")
        data_addr = self.w.put(arr.tobytes()) if arr.size else UNDEF
        layout = bytes([3, 1]) + struct.pack("<QQ", data_addr, arr.nbytes)
        msgs = [_message(0x0001, _dataspace(arr.shape)), _message(0x0003, _dt_float32()),
                _message(0x0008, layout)]
        return self.w.put(_object_header(msgs))

    def group(self, children: Dict[str, int], attrs: Optional[Dict[str, Any]] = None) -> Tuple[int, int, int]:
        """children: name -> object header address.  Returns (header address, btree, heap)."""
        names = sorted(children)
        heap = bytearray(b"\x00" * 8)
        offsets = {}
        for n in names:
            offsets[n] = len(heap)
            heap += n.encode("utf-8") + b"\x00"
            while len(heap) % 8:
                heap.append(0)
        heap_size = max(len(heap) + 16, 64)
        heap_data = bytes(heap) + b"\x00" * (heap_size - len(heap))
        # free block at the end of the data segment: next (1 = none) and size
        free_off = len(heap)
        heap_data = heap_data[:free_off] + struct.pack("<QQ", 1, heap_size - free_off) + heap_data[free_off + 16:]
        data_addr = self.w.put(heap_data)
        heap_addr = self.w.put(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", heap_size, free_off, data_addr))
        # symbol-table nodes of at most 8 entries (group leaf K = 4 -> 2K entries)
        per_node = 8
        leaves = []
        for i in range(0, max(len(names), 1), per_node):
            chunk = names[i:i + per_node]
            body = b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(chunk))
            for n in chunk:
                body += struct.pack("<QQII16x", offsets[n], children[n], 0, 0)
            body += b"\x00" * ((per_node - len(chunk)) * 40)
            leaves.append((self.w.put(body), offsets[chunk[-1]] if chunk else 0))
        if len(leaves) > 32:
            raise HDF5Error("too many children for one B-tree node in this minimal writer")
        node = b"TREE" + bytes([0, 0]) + struct.pack("<H", len(leaves)) + struct.pack("<QQ", UNDEF, UNDEF)
        node += struct.pack("<Q", 0)
        for addr, last_key in leaves:
            node += struct.pack("<QQ", addr, last_key)
        node += b"\x00" * ((32 - len(leaves)) * 16)
        btree_addr = self.w.put(node)
        msgs = [_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] + self.attr_messages(attrs or {})
        return self.w.put(_object_header(msgs)), btree_addr, heap_addr

    def finish(self, root: Tuple[int, int, int]) -> bytes:
        root_addr, btree, heap = root
        eof = len(self.w.buf)
        sb = SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96, len(sb)
        self.w.buf[:96] = sb
        return bytes(self.w.buf)


# =================================================================================================
# Keras model files
# =================================================================================================
def _node(*inbound):
    """One Keras ``inbound_nodes`` entry: [[layer, node_index, tensor_index, {}], ...]."""
    return [[name, node, tensor, {}] for name, node, tensor in inbound]


def keras_model_config(vecsize: int, units: int, attention: bool, n_classes: int, rnn: str = "GRU") -> dict:
    """The ``model.get_config()`` of ``create_model`` (reference ``deepgrp/model.py:293-336``): the layer graph
    the reference's ``tests/test_model.json`` pins (class names, layer names, inbound nodes, the fields that
    decide the arithmetic), for the GRU with attention, the GRU without and the LSTM."""
    rname = "BGRU" if rnn == "GRU" else "BLSTM"
    rcfg = {"name": rname, "trainable": True, "dtype": "float32", "return_sequences": True,
            "return_state": bool(attention), "go_backwards": False, "stateful": False, "unroll": False,
            "time_major": False, "units": units, "activation": "tanh", "recurrent_activation": "sigmoid",
            "use_bias": True, "dropout": 0.25, "recurrent_dropout": 0.0, "implementation": 2}
    if rnn == "GRU":
        rcfg["reset_after"] = True
    else:
        rcfg["unit_forget_bias"] = True

    def layer(cls, name, inbound, **cfg):
        return {"class_name": cls, "name": name, "inbound_nodes": inbound,
                "config": dict({"name": name, "trainable": True, "dtype": "float32"}, **cfg)}
    layers = [
        {"class_name": "InputLayer", "name": "input_1", "inbound_nodes": [],
         "config": {"batch_input_shape": [None, vecsize, 5], "dtype": "float32", "sparse": False,
                    "ragged": False, "name": "input_1"}},
        layer("Custom>ReverseComplement", "reverse_complement", [_node(("input_1", 0, 0))],
              complements=[3, 2, 1, 0, 4]),
        {"class_name": rnn, "name": rname, "config": rcfg,
         "inbound_nodes": [_node(("input_1", 0, 0)), _node(("reverse_complement", 0, 0))]},
    ]
    if attention:
        layers += [
            layer("Average", "average", [_node((rname, 0, 1), (rname, 1, 1))]),                 # the two last states
            layer("Reshape", "reshape", [_node(("average", 0, 0))], batch_input_shape=[None, units],
                  target_shape=[1, units]),
            layer("Average", "average_1", [_node((rname, 0, 0), (rname, 1, 0))]),               # the two sequences
            layer("AdditiveAttention", "additive_attention", [_node(("reshape", 0, 0), ("average_1", 0, 0))],
                  causal=False, dropout=0.0, use_scale=True),
            layer("Flatten", "flatten", [_node(("additive_attention", 0, 0))], data_format="channels_last"),
            layer("RepeatVector", "repeat_vector", [_node(("flatten", 0, 0))], n=vecsize),
            layer("Concatenate", "concatenate", [_node(("repeat_vector", 0, 0), ("average_1", 0, 0))], axis=-1),
        ]
        last = "concatenate"
    else:
        layers.append(layer("Average", "average", [_node((rname, 0, 0), (rname, 1, 0))]))
        last = "average"
    layers.append(layer("Dense", "FF", [_node((last, 0, 0))], units=n_classes, activation="linear", use_bias=True))
    layers.append(layer("Softmax", "softmax", [_node(("FF", 0, 0))], axis=2))
    return {"class_name": "Functional",
            "config": {"name": "model", "layers": layers, "input_layers": [["input_1", 0, 0]],
                       "output_layers": [["softmax", 0, 0]]}}


def check_keras_topology(config: dict) -> Tuple[str, int, int, bool]:
    """Refuse a ``model_config`` whose arithmetic the CUDA path does not implement, instead of computing
    something else silently: -> (rnn, vecsize, units, attention) of a DeepGRP graph (reference
    ``deepgrp/model.py:293-336`` / ``tests/test_model.json``), :class:`HDF5Error` otherwise."""
    layers = config["config"]["layers"] if "config" in config and "layers" in config["config"] else config["layers"]
    by_class: Dict[str, list] = {}
    for entry in layers:
        by_class.setdefault(entry["class_name"].split(">")[-1], []).append(entry)

    def need(cond, what):
        if not cond:
            raise HDF5Error("unsupported model: " + what)
    need("InputLayer" in by_class, "no InputLayer")
    shape = by_class["InputLayer"][0]["config"]["batch_input_shape"]
    need(len(shape) == 3 and shape[2] == 5, "input shape %s is not [None, vecsize, 5]" % (shape,))
    rnns = [k for k in ("GRU", "LSTM") if k in by_class]
    need(len(rnns) == 1 and len(by_class[rnns[0]]) == 1, "expected exactly one shared GRU or LSTM layer")
    rnn = rnns[0]
    entry = by_class[rnn][0]
    cfg = entry["config"]
    need(cfg.get("activation", "tanh") == "tanh" and cfg.get("recurrent_activation", "sigmoid") == "sigmoid",
         "%s activations %r / %r (tanh / sigmoid are implemented)" % (rnn, cfg.get("activation"), cfg.get("recurrent_activation")))
    need(cfg.get("use_bias", True) and cfg.get("return_sequences", True) and not cfg.get("go_backwards", False)
         and not cfg.get("stateful", False) and not cfg.get("time_major", False),
         "%s must use a bias, return sequences and run forward, stateless, batch-major" % rnn)
    if rnn == "GRU":
        need(cfg.get("reset_after", True), "GRU with reset_after=False (the reference builds reset_after=True)")
    need(len(entry.get("inbound_nodes", [[], []])) == 2, "the %s layer must be applied twice (input and its reverse complement)" % rnn)
    for rc in by_class.get("ReverseComplement", []):
        comp = rc["config"].get("complements", [3, 2, 1, 0, 4])
        need(list(comp) == [3, 2, 1, 0, 4], "ReverseComplement complements %s" % (comp,))
    attention = "AdditiveAttention" in by_class
    if attention:
        acfg = by_class["AdditiveAttention"][0]["config"]
        need(rnn == "GRU", "attention on an LSTM (the reference builds it for the GRU only)")
        need(acfg.get("use_scale", True) and not acfg.get("causal", False),
             "AdditiveAttention must have use_scale=True and causal=False")
        need(cfg.get("return_state", True), "attention needs the GRU's last state (return_state=True)")
        for cat in by_class.get("Concatenate", []):
            names = [x[0] for x in cat["inbound_nodes"][0]] if cat.get("inbound_nodes") else []
            need(len(names) != 2 or names[0].startswith("repeat_vector"),
                 "Concatenate order %s (attention context first, then the averaged sequence)" % (names,))
    need("Attention" not in by_class, "dot-product Attention layer")
    need("Dense" in by_class and len(by_class["Dense"]) == 1, "expected one Dense layer")
    dcfg = by_class["Dense"][0]["config"]
    need(dcfg.get("activation", "linear") == "linear" and dcfg.get("use_bias", True), "Dense must be linear with a bias")
    need("Softmax" in by_class and by_class["Softmax"][0]["config"].get("axis", -1) in (2, -1), "Softmax over the class axis")
    for cls in by_class:
        need(cls in ("InputLayer", "ReverseComplement", "GRU", "LSTM", "Average", "Reshape", "AdditiveAttention", "Flatten",
                     "RepeatVector", "Concatenate", "Dense", "Softmax", "Dropout"), "layer class %s" % cls)
    return rnn, int(shape[1]), int(cfg["units"]), attention


def save_keras_model(path: str, weights, config: Optional[dict] = None) -> None:
    """Write ``weights`` (a :class:`deepgrp_b200.model.ModelWeights`) as a Keras-layout HDF5 file; ``config``
    replaces the ``model_config`` (a ``get_config()`` dict or a Functional wrapper of one) and its layer names
    name the weight groups."""
    b = H5Builder()
    rnn_name = "BGRU" if weights.rnn == "GRU" else "BLSTM"
    cell = "gru_cell" if weights.rnn == "GRU" else "lstm_cell"
    layer_weights = {
        rnn_name: [("%s/%s/kernel:0" % (rnn_name, cell), weights.kernel),
                   ("%s/%s/recurrent_kernel:0" % (rnn_name, cell), weights.recurrent_kernel),
                   ("%s/%s/bias:0" % (rnn_name, cell), weights.bias)],
        "FF": [("FF/kernel:0", weights.ff_kernel), ("FF/bias:0", weights.ff_bias)],
    }
    if config is None:
        config = keras_model_config(weights.vecsize, weights.units, weights.attention, weights.n_classes,
                                    weights.rnn)
    if weights.attention:
        layer_weights["additive_attention"] = [("additive_attention/scale:0", weights.att_scale)]
    graph = config["config"] if "config" in config else config
    order = [entry["name"] for entry in graph["layers"]]           # every layer gets a group, as Keras writes them
    for entry in graph["layers"]:                                     # weights follow the config's layer names
        cls = entry["class_name"].split(">")[-1]
        canon = {"GRU": rnn_name, "LSTM": rnn_name, "AdditiveAttention": "additive_attention", "Dense": "FF"}.get(cls)
        if canon and canon in layer_weights and entry["name"] != canon:
            layer_weights[entry["name"]] = [(entry["name"] + w[len(canon):], a) for w, a in layer_weights.pop(canon)]

    def build(tree: dict, attrs=None):
        children = {}
        for name, node in tree.items():
            children[name] = build(node)[0] if isinstance(node, dict) else b.dataset(node)
        return b.group(children, attrs)

    layer_groups = {}
    for lname in order:
        tree: dict = {}
        names = []
        for wname, arr in layer_weights.get(lname, []):
            names.append(wname)
            node = tree
            parts = wname.split("/")
            for part in parts[:-1]:
                node = node.setdefault(part, {})
            node[parts[-1]] = arr
        layer_groups[lname] = build(tree, {"weight_names": [n.encode() for n in names]})[0]
    mw = b.group(layer_groups, {"layer_names": [n.encode() for n in order], "backend": "tensorflow",
                                "keras_version": "2.5.0"})
    if "class_name" not in config:
        config = {"class_name": "Functional", "config": config}
    root = b.group({"model_weights": mw[0]}, {"model_config": json.dumps(config), "backend": "tensorflow",
                                              "keras_version": "2.5.0"})
    with open(path, "wb") as fh:
        fh.write(b.finish(root))


def _as_str(v) -> str:
    if isinstance(v, bytes):
        return v.decode("utf-8")
    if isinstance(v, np.ndarray) and v.shape == ():
        return _as_str(v.item())
    return str(v)


def load_keras_model(path: str):
    """Read a Keras HDF5 model file into a :class:`deepgrp_b200.model.ModelWeights`.  Layer and weight
    names are taken from the ``layer_names`` / ``weight_names`` attributes (they can carry ``_1``
    suffixes), the architecture from ``model_config``."""
    from .model import ModelWeights
    f = H5File(path)
    if "model_config" not in f.attrs:
        raise HDF5Error("%s has no model_config attribute (not a Keras model file)" % path)
    config = json.loads(_as_str(f.attrs["model_config"]))
    rnn, vecsize, units, _ = check_keras_topology(config)
    layers = config["config"]["layers"]
    by_class: Dict[str, dict] = {}
    for layer in layers:
        by_class.setdefault(layer["class_name"].split(">")[-1], layer)
    mw = f["model_weights"]
    found: Dict[str, np.ndarray] = {}
    for lname in [_as_str(n) for n in np.atleast_1d(mw.attrs["layer_names"])]:
        g = mw[lname]
        if "weight_names" not in g.attrs:
            continue
        for wname in [_as_str(n) for n in np.atleast_1d(g.attrs["weight_names"])]:
            found[wname] = np.asarray(g[wname].value, dtype=np.float32)

    def pick(suffix: str, layer_hint: str):
        for k, v in found.items():
            if k.endswith(suffix) and layer_hint in k:
                return v
        return None
    rnn_layer = by_class[rnn]["config"]["name"]
    kernel = pick("/kernel:0", rnn_layer + "/")
    recurrent = pick("recurrent_kernel:0", rnn_layer + "/")
    bias = pick("bias:0", rnn_layer + "/")
    scale = pick("scale:0", by_class["AdditiveAttention"]["config"]["name"] + "/") if "AdditiveAttention" in by_class else None
    dense = by_class["Dense"]["config"]["name"]
    ffk, ffb = pick("kernel:0", dense + "/"), pick("bias:0", dense + "/")
    for name, v in (("kernel", kernel), ("recurrent_kernel", recurrent), ("bias", bias),
                    ("FF kernel", ffk), ("FF bias", ffb)):
        if v is None:
            raise HDF5Error("weight %s not found in %s" % (name, path))
    return ModelWeights(vecsize, units, kernel, recurrent, bias, ffk, ffb, scale, rnn)
