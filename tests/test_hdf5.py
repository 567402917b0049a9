"""Keras-HDF5 model files without h5py: the reader/writer subset of deepgrp_b200.hdf5 (CPU only)."""
import json
import struct

import numpy as np
import pytest

from deepgrp_b200 import hdf5, model


@pytest.mark.parametrize("T,U,att", [(342, 60, True), (150, 32, True), (150, 32, False), (77, 7, True)])
def test_keras_hdf5_roundtrip(tmp_path, T, U, att):
    w = model.random_weights(T, U, attention=att, seed=T)
    path = str(tmp_path / "model.hdf5")
    hdf5.save_keras_model(path, w)
    w2 = model.load_model(path)
    assert (w2.vecsize, w2.units, w2.attention, w2.n_classes, w2.rnn) == (T, U, att, 5, "GRU")
    assert w2.input_shape == (None, T, 5)          # what the CLI reads (reference __main__.py:270)
    for k, v in w.as_dict().items():
        assert np.array_equal(v, w2.as_dict()[k]), k


def test_file_structure_is_keras_layout(tmp_path):
    w = model.random_weights(150, 32, attention=True, seed=1)
    path = str(tmp_path / "m.h5")
    hdf5.save_keras_model(path, w)
    f = hdf5.H5File(path)
    assert f.buf[:8] == hdf5.SIGNATURE and f.buf[8] == 0          # superblock version 0
    cfg = json.loads(f.attrs["model_config"])                     # variable-length string via the global heap
    names = [l["class_name"] for l in cfg["config"]["layers"]]
    assert names == ["InputLayer", "ReverseComplement", "GRU", "Average", "AdditiveAttention", "Dense", "Softmax"]
    gru = cfg["config"]["layers"][2]["config"]
    assert gru["reset_after"] is True and gru["units"] == 32 and gru["recurrent_activation"] == "sigmoid"
    mw = f["model_weights"]
    assert [n.decode() for n in mw.attrs["layer_names"]] == ["input_1", "reverse_complement", "BGRU", "average_1",
                                                             "additive_attention", "FF", "softmax"]
    assert [n.decode() for n in mw["BGRU"].attrs["weight_names"]] == [
        "BGRU/gru_cell/kernel:0", "BGRU/gru_cell/recurrent_kernel:0", "BGRU/gru_cell/bias:0"]
    assert mw["BGRU"]["BGRU/gru_cell/kernel:0"].shape == (5, 96)
    assert mw["FF/FF/kernel:0"].value.shape == (64, 5)
    assert "nope" not in mw
    with pytest.raises(KeyError):
        mw["nope"]


def test_rejects_what_it_cannot_read(tmp_path):
    with pytest.raises(hdf5.HDF5Error):
        hdf5.H5File(b"not an hdf5 file at all")
    w = model.random_weights(30, 8, attention=False, seed=2)
    path = str(tmp_path / "m.h5")
    hdf5.save_keras_model(path, w)
    raw = bytearray(open(path, "rb").read())
    raw[8] = 2                                                    # pretend superblock version 2
    with pytest.raises(hdf5.HDF5Error):
        hdf5.H5File(bytes(raw))


def test_checkpoint_dir_lookup(tmp_path):
    from deepgrp_b200.prediction import setup_prediction_from_options_checkpoint
    from deepgrp_b200.model import Options
    w = model.random_weights(150, 32, attention=True, seed=3)
    hdf5.save_keras_model(str(tmp_path / "a.hdf5"), w)
    m = setup_prediction_from_options_checkpoint(Options(vecsize=150), tmp_path)
    assert np.array_equal(m.kernel, w.kernel)
    (tmp_path / "empty").mkdir()
    with pytest.raises(FileNotFoundError):
        setup_prediction_from_options_checkpoint(Options(), tmp_path / "empty")
