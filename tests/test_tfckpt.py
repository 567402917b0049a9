"""TensorFlow checkpoint ("tensor bundle") reader: table / protobuf / Snappy parsing on bundles written
by the module's own minimal writer (no TensorFlow here -- see the PARITY UNPINNED note in tfckpt.py),
and the route deepgrp.prediction.setup_prediction_from_options_checkpoint takes through it."""
import os

import numpy as np
import pytest

from deepgrp_b200 import tfckpt


def _keras_names(w, with_optimizer=True):
    t = {
        "layer_with_weights-0/cell/kernel/.ATTRIBUTES/VARIABLE_VALUE": w["kernel"],
        "layer_with_weights-0/cell/recurrent_kernel/.ATTRIBUTES/VARIABLE_VALUE": w["recurrent_kernel"],
        "layer_with_weights-0/cell/bias/.ATTRIBUTES/VARIABLE_VALUE": w["bias"],
        "layer_with_weights-2/kernel/.ATTRIBUTES/VARIABLE_VALUE": w["ff_kernel"],
        "layer_with_weights-2/bias/.ATTRIBUTES/VARIABLE_VALUE": w["ff_bias"],
        "save_counter/.ATTRIBUTES/VARIABLE_VALUE": np.array(7, dtype=np.int64),
    }
    if "att_scale" in w:
        t["layer_with_weights-1/scale/.ATTRIBUTES/VARIABLE_VALUE"] = w["att_scale"]
    if with_optimizer:
        t["layer_with_weights-0/cell/kernel/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE"] = w["kernel"] * 0
        t["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"] = np.array(123, dtype=np.int64)
    return t


@pytest.mark.parametrize("block_items", [1, 3, 64])
def test_bundle_round_trip(tmp_path, block_items):
    rng = np.random.default_rng(0)
    tensors = {"a/b/c": rng.normal(size=(3, 4)).astype(np.float32), "a/b/d": rng.integers(0, 9, size=(5,)).astype(np.int64),
               "scalar": np.array(2.5, dtype=np.float64), "z" * 300: rng.normal(size=(2, 2, 2)).astype(np.float32)}
    prefix = str(tmp_path / "07")
    tfckpt.write_bundle(prefix, tensors, block_items=block_items)
    got = tfckpt.read_bundle(prefix)
    assert set(got) == set(tensors)
    for k in tensors:
        assert got[k].dtype == tensors[k].dtype and np.array_equal(got[k], tensors[k])
    assert tfckpt.latest_checkpoint(str(tmp_path)) == prefix


def test_snappy_known_streams():
    # literal only
    assert tfckpt.snappy_decompress(bytes([5, 4 << 2]) + b"hello") == b"hello"
    # literal "ab" + copy (1-byte offset form) of 6 bytes at offset 2 -> "abababab"
    assert tfckpt.snappy_decompress(bytes([8, 1 << 2]) + b"ab" + bytes([((6 - 4) << 2) | 1, 2])) == b"abababab"
    # 2-byte-offset copy: "xyz" then 3 bytes at offset 3
    assert tfckpt.snappy_decompress(bytes([6, 2 << 2]) + b"xyz" + bytes([((3 - 1) << 2) | 2, 3, 0])) == b"xyzxyz"
    # long literal (length in one extra byte)
    payload = bytes(range(100))
    assert tfckpt.snappy_decompress(bytes([100, 60 << 2, 99]) + payload) == payload
    with pytest.raises(ValueError):
        tfckpt.snappy_decompress(bytes([9, 4 << 2]) + b"hello")


def test_compressed_block_is_read(tmp_path):
    """A table whose data block is stored as a Snappy literal (compression type 1)."""
    import struct
    prefix = str(tmp_path / "01")
    tfckpt.write_bundle(prefix, {"w": np.arange(6, dtype=np.float32).reshape(2, 3)}, block_items=64)
    raw = open(prefix + ".index", "rb").read()
    footer = raw[-48:]
    _, p = tfckpt._varint(footer, 0); msz, p = tfckpt._varint(footer, p)
    ioff, p = tfckpt._varint(footer, p); isz, p = tfckpt._varint(footer, p)
    first = next(tfckpt._block_entries(raw[ioff:ioff + isz]))[1]
    boff, q = tfckpt._varint(first, 0); bsz, _ = tfckpt._varint(first, q)
    block = raw[boff:boff + bsz]
    # Snappy stream of the block as one literal (length bsz <= 60 fits the short form, else 1-2 extra bytes)
    if bsz <= 60:
        comp = tfckpt._put_varint(bsz) + bytes([(bsz - 1) << 2]) + block
    elif bsz <= 256:
        comp = tfckpt._put_varint(bsz) + bytes([60 << 2, bsz - 1]) + block
    else:
        comp = tfckpt._put_varint(bsz) + bytes([61 << 2]) + struct.pack("<H", bsz - 1) + block
    # rebuild the file: compressed data block, then a fresh meta block, index block and footer
    crc = lambda b: struct.pack("<I", tfckpt._masked_crc(b))
    out = bytearray(comp + b"\x01" + crc(comp + b"\x01"))
    meta = tfckpt._build_block([])
    mh = tfckpt._put_varint(len(out)) + tfckpt._put_varint(len(meta))
    out += meta + b"\x00" + crc(meta + b"\x00")
    ib = tfckpt._build_block([(b"w", tfckpt._put_varint(0) + tfckpt._put_varint(len(comp)))], restart_interval=1)
    ih = tfckpt._put_varint(len(out)) + tfckpt._put_varint(len(ib))
    out += ib + b"\x00" + crc(ib + b"\x00")
    f = mh + ih
    out += f + b"\x00" * (40 - len(f)) + struct.pack("<Q", tfckpt.TABLE_MAGIC)
    open(prefix + ".index", "wb").write(bytes(out))
    got = tfckpt.read_bundle(prefix)
    assert np.array_equal(got["w"], np.arange(6, dtype=np.float32).reshape(2, 3))


@pytest.mark.parametrize("rnn,attention", [("GRU", True), ("GRU", False), ("LSTM", False)])
def test_setup_prediction_from_tf_checkpoint(tmp_path, rnn, attention):
    from deepgrp_b200 import model, prediction
    ref = model.random_weights(150, 32, attention=attention, seed=5)
    w = ref.as_dict()
    if rnn == "LSTM":
        rng = np.random.default_rng(1)
        w = dict(kernel=rng.normal(size=(5, 128)).astype(np.float32), recurrent_kernel=rng.normal(size=(32, 128)).astype(np.float32),
                 bias=rng.normal(size=(128,)).astype(np.float32), ff_kernel=w["ff_kernel"], ff_bias=w["ff_bias"])
    for epoch in ("01", "02"):
        scaled = {k: v * (2.0 if epoch == "02" else 1.0) for k, v in w.items()}
        tfckpt.write_bundle(str(tmp_path / epoch), _keras_names(scaled))
    opts = model.Options()
    opts.vecsize = 150
    got = prediction.setup_prediction_from_options_checkpoint(opts, tmp_path)
    assert got.rnn == rnn and got.attention == attention and got.vecsize == 150 and got.units == 32
    for k, v in w.items():
        exp = (v * 2.0).reshape(1, -1) if (k == "bias" and v.ndim == 1) else v * 2.0     # epoch 02 is the latest
        assert np.array_equal(getattr(got, k), exp), k


def test_directory_without_checkpoint_falls_back_to_weight_files(tmp_path):
    from deepgrp_b200 import model, prediction
    ref = model.random_weights(150, 32, attention=True, seed=5)
    ref.save_npz(str(tmp_path / "weights.npz"))
    opts = model.Options()
    opts.vecsize = 200
    got = prediction.setup_prediction_from_options_checkpoint(opts, tmp_path)
    assert got.vecsize == 200 and np.array_equal(got.kernel, ref.kernel)
    with pytest.raises(FileNotFoundError):
        prediction.setup_prediction_from_options_checkpoint(opts, tmp_path / "nothing" if (tmp_path / "nothing").mkdir() is None else tmp_path)


def test_crc32c_known_answers():
    """Standard CRC-32C check values (RFC 3720 B.4; the same vectors LevelDB's crc32c_test.cc uses)."""
    assert tfckpt.crc32c(b"123456789") == 0xE3069283
    assert tfckpt.crc32c(bytes(32)) == 0x8A9136AA
    assert tfckpt.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tfckpt.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfckpt.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    c = tfckpt.crc32c(b"foo")
    m = tfckpt._masked_crc(b"foo")
    assert m != c
    u = (m - 0xA282EAD8) & 0xFFFFFFFF                      # LevelDB's Unmask
    assert ((u >> 17) | (u << 15)) & 0xFFFFFFFF == c


def test_corrupt_bundle_is_rejected(tmp_path):
    prefix = str(tmp_path / "01")
    tfckpt.write_bundle(prefix, {"a": np.arange(6, dtype=np.float32).reshape(2, 3)})
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[5] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="checksum"):
        tfckpt.read_bundle(prefix)
    assert tfckpt.read_bundle(prefix, verify=False)["a"].shape == (2, 3)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[3] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError, match="checksum"):
        tfckpt.read_bundle(prefix, verify=False)
