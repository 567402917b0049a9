"""deepgrp_b200.preprocessing: the reference's known answers for the label matrix, the edge-N cut and the
.npz one-hot file (reference tests/test_preprocessing.py:11-52, tests/test_preprocess_sequence.py:11-25)."""
import gzip
import io

import numpy as np
import pytest

from deepgrp_b200 import preprocessing as dgpreprocess


def test_label_matrix_known_answer(tmp_path):
    """The vector of the reference's test_preprocess_y: only chr1 lines whose repeat number is searched for
    count (the chr2 line and repeat 7 do not), row 0 is the complement."""
    path = tmp_path / "annotation.txt"
    path.write_text("chr1 5 10 2 X\nchr2 6 11 5 X\nchr1 13 15 4 X\nchr1 16 18 7 X\n")
    y = dgpreprocess.preprocess_y(str(path), "chr1", 20, [1, 2, 3, 4])
    assert y.dtype == np.int8 and y.shape == (5, 20)
    want = {2: set(range(5, 10)), 4: set(range(13, 15)), 1: set(), 3: set()}
    for row, cols in want.items():
        assert set(np.flatnonzero(y[row]).tolist()) == cols
    assert set(np.flatnonzero(y[0]).tolist()) == set(range(20)) - want[2] - want[4]
    assert (y.sum(axis=0) == 1).all()


@pytest.mark.parametrize("lead,trail", [(0, 0), (0, 10), (10, 0), (10, 10), (20, 10), (10, 20), (20, 20), (0, 20), (20, 0)])
def test_edge_n_cut_known_answer(lead, trail):
    """The reference's test_drop_start_end_n: 100 columns, `lead` / `trail` N columns around C columns; the cut
    keeps [lead, 100 - trail - 1) -- the last base column is lost with the N run (reference behaviour)."""
    n = 100
    fwd = np.zeros((5, n))
    fwd[4, :lead] = 1
    fwd[1, lead:n - trail] = 1
    fwd[4, n - trail:] = 1
    positions = np.arange(n)[None, :]
    x, y = dgpreprocess.drop_start_end_n(fwd, positions)
    kept = n - lead - trail - 1
    assert x.shape == (5, kept) and y.shape == (1, kept)
    assert y[0].tolist() == list(range(lead, n - trail - 1))
    assert x[1].all() and not x[[0, 2, 3, 4]].any()


def test_fastaparser_hash_known_answer():
    header, md5, seq = dgpreprocess.fastaparser(io.BytesIO(b">test\nACGTNACGTN\n"))
    assert (header, md5, seq) == ("test", "ff8ed7aaa145d49602bf5fdf5e5b8338", "ACGTNACGTN")
    # the hash covers the lines as they are in the file (case included), the sequence is upper-cased
    header, md5_lower, seq = dgpreprocess.fastaparser(io.BytesIO(b">t x\nacgtn\nACGTN\n"))
    assert header == "t x" and seq == "ACGTNACGTN" and md5_lower != md5


def test_load_onehot_npz_round_trip(tmp_path):
    fwd = np.eye(5, dtype=np.int8)[:, [0, 1, 2, 3, 4, 4, 0]]
    np.savez_compressed(str(tmp_path / "x.fa.gz"), fwd=fwd, hash=np.array(["abc"]))
    got, md5 = dgpreprocess.load_onehot_npz(str(tmp_path / "x.fa.gz.npz"))
    assert md5 == "abc" and got.dtype == np.int8 and np.array_equal(got, fwd)


@pytest.mark.gpu
def test_preprocess_sequence_known_answer(tmp_path, gpu_ctx):
    inputs = tmp_path / "inputs.fa.gz"
    with gzip.open(inputs, "w") as fh:
        fh.write(b">test\nACGTNACGTN\n")
    out = tmp_path / "inputs.fa.gz.npz"
    assert not out.exists()
    dgpreprocess.main([str(inputs)])
    assert out.exists()
    got = np.load(out)
    expected = [[1, 0, 0, 0, 0], [0, 1, 0, 0, 0], [0, 0, 1, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 1]] * 2
    np.testing.assert_array_equal(got["fwd"], np.array(expected).T)
    assert got["hash"][0] == "ff8ed7aaa145d49602bf5fdf5e5b8338"
    # unchanged input: not rewritten; edge N runs are kept (nothing is trimmed in this format); bad characters
    assert dgpreprocess.preprocess_sequence(str(inputs)) is False
    assert dgpreprocess.preprocess_sequence(str(inputs), force=True) is True
    with gzip.open(inputs, "w") as fh:
        fh.write(b">t\nNNNacgtNN\nNN\n")
    assert dgpreprocess.preprocess_sequence(str(inputs)) is True
    fwd, _ = dgpreprocess.load_onehot_npz(str(out))
    assert fwd.shape == (5, 11) and fwd[4].tolist() == [1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1]
    assert fwd[:4, 3:7].tolist() == np.eye(4, dtype=int).tolist()
    with pytest.raises(KeyError):
        dgpreprocess.one_hot_untrimmed("ACGTRN")
    assert dgpreprocess.one_hot_untrimmed("NNNN")[4].tolist() == [1, 1, 1, 1]
