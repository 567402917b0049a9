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
    assert names == ["InputLayer", "Custom>ReverseComplement", "GRU", "Average", "Reshape", "Average", "AdditiveAttention",
                     "Flatten", "RepeatVector", "Concatenate", "Dense", "Softmax"]
    gru = cfg["config"]["layers"][2]["config"]
    assert gru["reset_after"] is True and gru["units"] == 32 and gru["recurrent_activation"] == "sigmoid"
    mw = f["model_weights"]
    assert [n.decode() for n in mw.attrs["layer_names"]] == ["input_1", "reverse_complement", "BGRU", "average", "reshape",
                                                             "average_1", "additive_attention", "flatten", "repeat_vector",
                                                             "concatenate", "FF", "softmax"]
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


# ---- the reference's own model configs (tests/golden/keras_model_configs.json <- reference tests/test_model.json)
def _reference_configs():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keras_model_configs.json")
    return json.load(open(path))


def _graph(cfg):
    """(class, name, inbound (layer, node, tensor) triples, arithmetic-relevant config) per layer."""
    keys = ("units", "activation", "recurrent_activation", "use_bias", "reset_after", "unit_forget_bias",
            "return_sequences", "return_state", "go_backwards", "implementation", "complements", "use_scale",
            "causal", "axis", "n", "target_shape", "batch_input_shape")
    return [(e["class_name"], e["name"], [[tuple(x[:3]) for x in node] for node in e["inbound_nodes"]],
             {k: e["config"][k] for k in keys if k in e["config"]}) for e in cfg["layers"]]


@pytest.mark.parametrize("version", ["2.1", "2.2", "2.3", "2.4", "2.5"])
def test_written_config_is_the_reference_layer_graph(version):
    """keras_model_config reproduces the layer graph create_model has under every TensorFlow version the
    reference pins in tests/test_model.json (reference tests/test_model.py:254-262 compares the same way)."""
    ref = _reference_configs()[version]
    ours = hdf5.keras_model_config(150, 32, True, 5, "GRU")["config"]
    assert _graph(ours) == _graph(ref["GRU"])
    assert ours["input_layers"] == ref["GRU"]["input_layers"] and ours["output_layers"] == ref["GRU"]["output_layers"]
    # the LSTM fixture was built second in the reference's test session: Keras numbered its layers input_2,
    # reverse_complement_1, average_2, softmax_1; same graph after renaming
    lstm = json.loads(json.dumps(ref["LSTM"]).replace("input_2", "input_1").replace("reverse_complement_1", "reverse_complement")
                      .replace("average_2", "average").replace("softmax_1", "softmax"))
    assert _graph(hdf5.keras_model_config(150, 32, False, 5, "LSTM")["config"]) == _graph(lstm)


@pytest.mark.parametrize("version", ["2.1", "2.5"])
@pytest.mark.parametrize("kind", ["GRU", "LSTM"])
def test_loader_reads_files_with_the_reference_config(tmp_path, version, kind):
    """A model file whose model_config and layer names are the reference's own (incl. the _1 / _2 suffixes)."""
    cfg = _reference_configs()[version][kind]
    w = model.random_weights(150, 32, attention=(kind == "GRU"), seed=9)
    if kind == "LSTM":
        rng = np.random.default_rng(1)
        w = model.ModelWeights(150, 32, rng.normal(size=(5, 128)).astype(np.float32), rng.normal(size=(32, 128)).astype(np.float32),
                               rng.normal(size=(1, 128)).astype(np.float32), w.ff_kernel, w.ff_bias, None, "LSTM")
    path = str(tmp_path / "ref.hdf5")
    hdf5.save_keras_model(path, w, config=cfg)
    got = model.load_model(path)
    assert (got.rnn, got.vecsize, got.units, got.attention) == (kind, 150, 32, kind == "GRU")
    for k, v in w.as_dict().items():
        assert np.array_equal(v, got.as_dict()[k]), k
    f = hdf5.H5File(path)
    assert [n.decode() for n in f["model_weights"].attrs["layer_names"]] == [e["name"] for e in cfg["layers"]]


@pytest.mark.parametrize("edit,what", [
    (lambda g: g["layers"][2]["config"].update(reset_after=False), "reset_after"),
    (lambda g: g["layers"][2]["config"].update(recurrent_activation="hard_sigmoid"), "activations"),
    (lambda g: g["layers"][2]["config"].update(go_backwards=True), "forward"),
    (lambda g: g["layers"][1]["config"].update(complements=[0, 1, 2, 3, 4]), "complements"),
    (lambda g: g["layers"][6]["config"].update(use_scale=False), "use_scale"),
    (lambda g: g["layers"][9].update(inbound_nodes=[[["average_1", 0, 0, {}], ["repeat_vector", 0, 0, {}]]]), "Concatenate"),
    (lambda g: g["layers"][11]["config"].update(axis=1), "Softmax"),
    (lambda g: g["layers"][10]["config"].update(activation="relu"), "Dense"),
    (lambda g: g["layers"].insert(3, {"class_name": "Conv1D", "name": "c", "inbound_nodes": [], "config": {"name": "c"}}), "Conv1D"),
])
def test_loader_refuses_graphs_it_does_not_compute(tmp_path, edit, what):
    """A config whose arithmetic differs from the reference's graph is an error, not a silently different result."""
    cfg = hdf5.keras_model_config(150, 32, True, 5, "GRU")
    edit(cfg["config"])
    path = str(tmp_path / "odd.hdf5")
    hdf5.save_keras_model(path, model.random_weights(150, 32, attention=True, seed=1), config=cfg)
    with pytest.raises(hdf5.HDF5Error, match=what):
        model.load_model(path)
