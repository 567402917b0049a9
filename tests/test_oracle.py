"""CPU tests of the oracle itself (the checker must be right before it checks anything):
* against golden vectors generated from the reference's own compiled natives
  (tests/golden/native_vectors.npz, made by tests/golden/make_golden.py);
* against the reference's known-answer tests, restated (reference tests/test_sequence.py,
  tests/test_mss.py, tests/test_prediction.py, tests/test_model.py:215-229);
* directly against oracle/_ref when it is present (the build container);
* forward numerics (parity unpinned in the reference): numpy float32 vs torch.nn.GRU vs float64.
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "native_vectors.npz"))


def test_oracle_encode_golden(oracle, golden):
    for i in range(int(golden["enc_n"])):
        text = golden["enc_text_%d" % i].tobytes()
        st, fwd = oracle.one_hot_encode_bytes(text)
        assert st == int(golden["enc_start_%d" % i])
        assert np.array_equal(fwd, golden["enc_fwd_%d" % i])


def test_oracle_get_max_golden(oracle, golden):
    for i in range(int(golden["max_n"])):
        out = oracle.get_max(golden["max_base_%d" % i].copy(), golden["max_in_%d" % i],
                             int(golden["max_stride_%d" % i]))
        assert np.array_equal(out, golden["max_out_%d" % i])


def test_oracle_segments_golden(oracle, golden):
    for i in range(int(golden["seg_n"])):
        got = oracle.yield_segments_array(golden["seg_lab_%d" % i], 11)
        assert np.array_equal(got, golden["seg_out_%d" % i])


def test_oracle_mss_golden(oracle, golden):
    for k in range(int(golden["mss_n"])):
        ml, xd = (int(v) for v in golden["mss_par_%d" % k])
        got = oracle.find_mss_relabel(golden["mss_s_%d" % k], golden["mss_lab_%d" % k], 5, ml, xd)
        assert np.array_equal(got, golden["mss_out_%d" % k]), k
        full = oracle.find_mss_labels(golden["mss_s_%d" % k], golden["mss_lab_%d" % k], 5, ml, xd)
        assert np.array_equal(full.argmax(axis=1), golden["mss_out_%d" % k])


def test_oracle_all_n_and_case(oracle):
    with pytest.raises(ValueError):
        oracle.one_hot_encode_dna_sequence("NNNN")
    st, fwd = oracle.one_hot_encode_dna_sequence("nnACGTnn")     # lowercase n is NOT trimmed
    assert st == 0 and fwd.shape == (5, 8) and fwd[4].tolist() == [1, 1, 0, 0, 0, 0, 1, 1]
    st, fwd = oracle.one_hot_encode_dna_sequence("")
    assert st == 0 and fwd.shape == (5, 0)


def test_oracle_segment_last_element_quirk(oracle):
    # a run reaching the last element is split at size-1
    lab = np.array([0, 2, 2, 2, 2], np.int64)
    assert list(oracle.yield_segments(lab, 100)) == [(101, 104, 2), (104, 105, 2)]
    assert list(oracle.yield_segments(np.array([1, 0, 0], np.int64), 0)) == [(0, 1, 1), (2, 3, 0)]


def test_oracle_min_sc_truncation(oracle):
    # mss.c:35: `int min_sc` -- a 4.3-score segment passes min_mss_len=1 (threshold trunc(4.595)=4)
    s = np.array([-1.0, 4.3, -1.0])
    out = oracle.mss_find_all(s, np.log(99.0) * 1, -1.0)
    assert out.size == 1 and (out["st"][0], out["en"][0]) == (1, 2)
    out = oracle.mss_find_all(np.array([-1.0, 3.9, -1.0]), np.log(99.0) * 1, -1.0)
    assert out.size == 0


def test_oracle_score_transform_known_answer(oracle):
    # reference tests/test_prediction.py:39-64
    probs = np.array([[0.9, 0.1, 0, 0, 0], [0.1, 0.9, 0, 0, 0], [0.995, 0.005, 0, 0, 0],
                      [0, 0, 0, 0, 0]], dtype=np.float32)
    sc, cl = oracle.apply_mss_scores(probs)
    assert cl.tolist() == [0, 1, 0, 0]
    m = np.float32(0.9) + np.float32(1e-6)
    t = np.log(m / (1 - m))
    assert sc[0] == pytest.approx(-10 * t, rel=1e-6) and sc[1] == pytest.approx(t, rel=1e-6)
    assert sc[2] == pytest.approx(-10 * np.log(0.99 / 0.01), rel=1e-5)
    assert sc[3] == pytest.approx(138.155, abs=1e-2)            # never-covered row scores POSITIVE


def test_oracle_softmax_vs_scipy(oracle):
    from scipy.special import softmax as sp_softmax
    rng = np.random.default_rng(0)
    x = rng.random((100, 5)).astype(np.float32)
    assert np.allclose(oracle.softmax(x), sp_softmax(x, axis=1), atol=1e-6)


def test_oracle_windowing_and_partial_batch(oracle):
    # reference tests/test_prediction.py:16-36: ceil((L-T)/step) windows, short final batch
    L, T, step, B = 5000, 150, 50, 16
    fwd = np.zeros((5, L), np.int8)
    fwd[0] = 1
    batches = list(oracle.fetch_validation_batch(fwd, step, B, T))
    n_win = -(-(L - T) // step)
    assert sum(b.shape[0] for b in batches) == n_win == 97
    assert batches[-1].shape[0] == n_win % B == 1
    # prediction.py:105: the short batch lands at i * b_last * step
    pred = oracle.predict(lambda b: np.ones(b.shape[:2] + (5,), np.float32), iter(batches), (L, 5), step)
    covered = pred[:, 0] > 0
    assert covered[:(96 - 1) * step + T].all()
    assert not covered[(96 - 1) * step + T:].any()            # true tail window never placed


def test_oracle_reverse_complement_golden(oracle):
    # reference tests/test_model.py:215-229: reverse + channels [3,2,1,0,4]
    x = np.eye(5, dtype=np.float32)[[0, 1, 2, 3, 4, 0]][None]
    rc = oracle.reverse_complement(x)[0].argmax(axis=1)
    assert rc.tolist() == [3, 4, 0, 1, 2, 3]


def test_oracle_forward_engines_agree(oracle):
    from deepgrp_b200.model import random_weights
    for T, U, att in ((150, 32, True), (60, 60, True), (40, 16, False)):
        w = random_weights(T, U, attention=att, seed=1).as_dict()
        rng = np.random.default_rng(T)
        batch = np.eye(5, dtype=np.float32)[rng.integers(0, 5, size=(6, T))]
        a = oracle.model_forward(batch, w, dtype=np.float32, engine="numpy")
        b = oracle.model_forward(batch, w, dtype=np.float32, engine="torch")
        c = oracle.model_forward(batch, w, dtype=np.float64, engine="numpy")
        assert np.abs(a - c).max() < 5e-6 and np.abs(b - c).max() < 5e-6
        assert np.allclose(a.sum(axis=2), 1, atol=1e-5)


def test_oracle_fasta_reader(oracle, tmp_path):
    # reference tests/test_main.py:238-247 + the quirks of SURVEY.md section 8a row a1
    import io
    text = "ACGT\n>h1 desc\nacgt\nNNAC\n>\nGGGG\n>h3\nTT"
    recs = list(oracle.read_multi_fasta(io.StringIO(text)))
    assert recs == [("h1 desc", "ACGTNNAC"), ("h3", "TT")]
    with pytest.raises(IndexError):
        list(oracle.read_multi_fasta(io.StringIO(">a\nAC\n\nGT\n")))


def test_oracle_against_compiled_reference(oracle):
    from oracle import build_ref
    mods = build_ref.load()
    if mods is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    ref_seq, ref_mss = mods
    rng = np.random.default_rng(9)
    for trial in range(40):
        n = int(rng.integers(1, 3000))
        s = rng.normal(size=n) * 3 - rng.random()
        lab = rng.integers(0, 5, size=n)
        for ml, xd in ((0, -1), (3, 2), (50, 50)):
            assert np.array_equal(ref_mss.find_mss_labels(s, lab, 5, ml, xd),
                                  oracle.find_mss_labels(s, lab, 5, ml, xd))
        text = "".join(np.array(list("ACGTNacgtnX"))[rng.integers(0, 11, size=n)])
        a, b = ref_seq.one_hot_encode_dna_sequence(text) if set(text) != {"N"} else (None, None), None
        if a[0] is not None:
            st, fwd = oracle.one_hot_encode_dna_sequence(text)
            assert st == a[0] and np.array_equal(fwd, a[1])
        assert list(ref_seq.yield_segments(lab, 5)) == list(oracle.yield_segments(lab, 5))


def test_oracle_lstm_engines_agree(oracle):
    from deepgrp_b200.model import random_weights
    w = random_weights(60, 24, attention=True, seed=1, rnn="LSTM").as_dict()
    rng = np.random.default_rng(0)
    batch = np.eye(5, dtype=np.float32)[rng.integers(0, 5, size=(4, 60))]
    a = oracle.model_forward(batch, w, engine="numpy")
    b = oracle.model_forward(batch, w, engine="torch")
    c = oracle.model_forward(batch, w, dtype=np.float64)
    assert np.abs(a - c).max() < 5e-6 and np.abs(b - c).max() < 5e-6


@pytest.mark.parametrize("min_len", (10, 20))
def test_oracle_filter_segments_reference_vector(oracle, min_len):
    """The known-answer case of the reference's tests/test_prediction.py:183-195."""
    segment_length = min_len * 2
    data = np.zeros(1000)
    data[110:110 + segment_length] = 1
    data[210 + segment_length:210 + 2 * segment_length] = 1
    expected = data.copy()
    data[0:min_len - 1] = 1
    data[120 + segment_length:120 + segment_length + min_len - 1] = 1
    data[(-min_len) + 1:] = 1
    oracle.filter_segments(data, min_len=min_len)
    np.testing.assert_equal(data, expected)


def test_oracle_confusion_matrix_vs_histogram(oracle):
    """tests/test_prediction.py:175-181 checks against pycm (absent here): same counts via bincount."""
    rng = np.random.default_rng(3)
    t = rng.integers(0, 4, size=100)
    p = rng.integers(0, 4, size=100)
    got = oracle.confusion_matrix(t, p)
    np.testing.assert_equal(got, np.bincount(t * 4 + p, minlength=16).reshape(4, 4))
