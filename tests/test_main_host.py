"""deepgrp_b200.__main__ without a GPU: the wiring the reference tests by substituting collaborators
(tests/test_main.py:38-46 parser state, :165-193 run -> Options, :196-236 _predict, :47-83 predict) and the rows
the stepwise route writes (deepgrp/__main__.py:275-292).  Kernels are not involved: where a GPU call would sit a
stand-in (or the oracle, as the checker) is patched in."""
import argparse
import gzip

import numpy as np
import pytest

from deepgrp_b200 import __main__ as cli
from deepgrp_b200 import model as dgmodel


def test_parser_initial_state():
    parser = cli.CommandLineParser()
    assert isinstance(parser.parser, argparse.ArgumentParser)
    assert (parser.threads, parser.xla, parser.verbose, parser.args) == (1, False, 0, None)


@pytest.mark.parametrize("command", ["train", "predict"])
@pytest.mark.parametrize("xdrop_len,batch_size,min_mss_length", [(23, 100, 10), (15, 150, 15)])
def test_run_builds_options_and_dispatches(monkeypatch, command, xdrop_len, batch_size, min_mss_length):
    parser = cli.CommandLineParser()
    parser.args = argparse.Namespace(command=command, xdrop_length=xdrop_len, batch_size=batch_size,
                                     min_mss_length=min_mss_length)
    seen = []

    def record(name):
        def call(args, options):
            assert args is parser.args
            assert (options.min_mss_len, options.batch_size, options.xdrop_len) == (min_mss_length, batch_size, xdrop_len)
            seen.append(name)
        return call
    monkeypatch.setattr(parser, "train", record("train"))
    monkeypatch.setattr(parser, "predict", record("predict"))
    parser.run()
    assert seen == [command]


@pytest.mark.parametrize("use_mss", (True, False))
def test_predict_one_sequence_wiring(monkeypatch, oracle, use_mss):
    """_predict: encode -> windows -> predict -> apply_mss | softmax -> argmax, with the trimmed length and the
    model's class count as the output shape (reference deepgrp/__main__.py:46-83)."""
    opt = dgmodel.Options(batch_size=10, vecsize=10)
    weights = dgmodel.random_weights(10, 4, attention=True, seed=0)
    rng = np.random.default_rng(0)
    seq = "NN" + "".join(rng.choice(list("NACGT"), size=100)) + "A" + "NNN"
    calls = []
    monkeypatch.setattr(cli.dgsequence, "one_hot_encode_dna_sequence", oracle.one_hot_encode_dna_sequence)

    def fake_predict(mdl, data_iterator, output_shape, step_size):
        assert mdl is weights and step_size == 3
        assert (data_iterator.step_size, data_iterator.batch_size, data_iterator.vecsize) == (3, 10, 10)
        assert output_shape == (data_iterator.length, 5)
        calls.append("predict")
        return np.tile(np.array([0.1, 0.6, 0.1, 0.1, 0.1], np.float32), (output_shape[0], 1))

    def fake_mss(prediction, options):
        assert use_mss and options is opt and isinstance(prediction, np.ndarray)
        calls.append("mss")
        return np.eye(5)[np.full(prediction.shape[0], 2)]

    def fake_softmax(prediction):
        assert not use_mss
        calls.append("softmax")
        return prediction
    monkeypatch.setattr(cli.dgpred, "predict", fake_predict)
    monkeypatch.setattr(cli.dgpred, "apply_mss", fake_mss)
    monkeypatch.setattr(cli.dgpred, "softmax", fake_softmax)
    labels, startpos = cli._predict(dnasequence=seq, model=weights, options=opt, step_size=3, use_mss=use_mss)
    exp_start, exp_fwd = oracle.one_hot_encode_dna_sequence(seq)
    assert startpos == exp_start and labels.shape == (exp_fwd.shape[1],)
    assert calls == ["predict", "mss" if use_mss else "softmax"]
    assert (labels == (2 if use_mss else 1)).all()


@pytest.mark.parametrize("compressed", (False, True))
def test_stepwise_predict_writes_the_reference_rows(monkeypatch, oracle, tmp_path, compressed):
    """`deepgrp predict model a.fa b.fa --stepwise --output out`: per record, label > 0 segments as
    file <TAB> header <TAB> start+startpos <TAB> end+startpos <TAB> label, one stream for all files."""
    records = {"a": [("chr1 first", "ACGT" * 5), ("chr2", "GG" * 10)], "b": [("chrX", "T" * 20)]}
    labels = {"ACGT" * 5: (np.array([0] * 4 + [2] * 6 + [0] * 3 + [1] * 7), 5),
              "GG" * 10: (np.zeros(20, dtype=np.int64), 0),
              "T" * 20: (np.array([3] * 20), 100)}
    paths = []
    for stem, recs in records.items():
        path = tmp_path / (stem + (".fa.gz" if compressed else ".fa"))
        text = "".join(">%s\n%s\n%s\n" % (h, s[:7].lower(), s[7:]) for h, s in recs)
        (gzip.open(path, "wt") if compressed else open(path, "w")).write(text)
        paths.append(str(path))

    class Model:
        input_shape = (None, 77, 5)
    loaded = []
    monkeypatch.setattr(cli.dgmodel, "load_model", lambda path: loaded.append(path) or Model())

    def fake_predict(dnasequence, model, options, step_size, use_mss):
        assert isinstance(model, Model) and options.vecsize == 77 and step_size == 25 and use_mss is True
        lab, start = labels[dnasequence]                      # the reader upper-cases the sequence
        return lab.astype(np.int64), start
    monkeypatch.setattr(cli, "_predict", fake_predict)
    monkeypatch.setattr(cli.dgsequence, "yield_segments", oracle.yield_segments)
    out = tmp_path / "out.tsv"
    cli.CommandLineParser().parse_args(["-s", "25", "predict", "model.hdf5", *paths, "--stepwise", "--output",
                                        str(out)]).set_logging().setup_tensorflow().run()
    assert loaded == ["model.hdf5"]
    a, b = paths
    # a run that reaches the last element is split at size - 1 (reference sequence.pyx:43-51)
    assert out.read_text().splitlines() == [
        "%s\tchr1 first\t9\t15\t2" % a, "%s\tchr1 first\t18\t24\t1" % a, "%s\tchr1 first\t24\t25\t1" % a,
        "%s\tchrX\t100\t119\t3" % b, "%s\tchrX\t119\t120\t3" % b]


def test_train_is_refused():
    with pytest.raises(SystemExit):
        cli.CommandLineParser.train(argparse.Namespace(), dgmodel.Options())


def test_stepwise_route_with_the_oracle_in_place_of_the_kernels(monkeypatch, oracle, tmp_path):
    """The whole stepwise command line on the CPU: real files, real Options / window descriptor / row writer, and
    the oracle standing in for the five GPU calls.  The text must equal the oracle's own end-to-end restatement
    of the reference (deepgrp/__main__.py:46-83, 275-292), i.e. the host code between the calls adds nothing."""
    from conftest import random_dna, write_fasta
    T, U = 150, 32
    weights = dgmodel.random_weights(T, U, attention=True, seed=0).scaled(4.0)
    wd = weights.as_dict()
    fasta = tmp_path / "in.fa"
    write_fasta(str(fasta), [("r1 x", "NN" + random_dna(2500, 1, "ACGTacgtN") + "N"), ("r2", random_dna(120, 2)),
                             ("r3", random_dna(1800, 3))])
    monkeypatch.setattr(cli.dgmodel, "load_model", lambda path: weights)
    monkeypatch.setattr(cli.dgsequence, "one_hot_encode_dna_sequence", oracle.one_hot_encode_dna_sequence)
    monkeypatch.setattr(cli.dgsequence, "yield_segments", oracle.yield_segments)

    def predict(model, data, results_shape, step_size):
        assert model is weights
        return oracle.predict(lambda b: oracle.model_forward(b, wd), iter(data), results_shape, step_size)
    monkeypatch.setattr(cli.dgpred, "predict", predict)
    monkeypatch.setattr(cli.dgpred, "apply_mss", lambda p, o: oracle.apply_mss(p, o.min_mss_len, o.xdrop_len))
    monkeypatch.setattr(cli.dgpred, "softmax", oracle.softmax)
    for extra, use_mss in ((["--stepwise"], True), (["--stepwise", "--no_use_mss"], False)):
        out = tmp_path / ("out%d.tsv" % use_mss)
        cli.CommandLineParser().parse_args(["-b", "16", "-s", "50", "predict", "m.hdf5", str(fasta), "--output",
                                            str(out)] + extra).run()
        expected = oracle.predict_fasta_tsv(str(fasta), wd, T, 16, 50, use_mss, 50, 50)
        assert out.read_text() == expected and expected.count("\n") > 3
