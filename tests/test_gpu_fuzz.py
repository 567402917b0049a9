"""Seeded sweep over model shapes and window parameters: every forward kernel (two-tile tcgen05, wide tcgen05 with one
CTA or a CTA pair, fp32 FFMA) against the oracle through `predict`, including the reference's placement of the short
last batch (deepgrp/prediction.py:105), steps larger than the window, windows longer than the record, 1..128 units."""
import numpy as np
import pytest

from conftest import random_dna

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(2024)
    out = []
    for k in range(28):
        U = int(rng.choice([1, 3, 8, 17, 32, 33, 60, 64, 65, 90, 128]))
        T = int(rng.choice([5, 16, 31, 64, 150, 257, 342, 512]))
        step = int(rng.choice([1, 7, 50, T, 2 * T + 3])) if T <= 64 else int(rng.choice([13, 50, 97, T, T + 11]))
        batch = int(rng.choice([1, 5, 64, 256]))
        L = int(rng.integers(max(T // 2, 3), 6 * T + 900))
        rnn = "LSTM" if (k % 5 == 4 and U <= 64) else "GRU"
        att = bool(rng.integers(0, 2))
        out.append((k, T, U, step, batch, L, rnn, att))
    return out


@pytest.mark.parametrize("k,T,U,step,batch,L,rnn,att", _cases())
def test_predict_matches_oracle_on_random_shapes(gpu_ctx, oracle, k, T, U, step, batch, L, rnn, att):
    import deepgrp_b200.model as model
    import deepgrp_b200.prediction as pred
    import deepgrp_b200.sequence as seq
    w = model.random_weights(T, U, attention=att, seed=100 + k, rnn=rnn).scaled(2.0)
    rng = np.random.default_rng(k)
    if rnn == "GRU":
        w.bias[:] = rng.normal(scale=0.2, size=w.bias.shape).astype(np.float32)      # non-zero biases
    w.ff_bias[:] = rng.normal(scale=0.2, size=w.ff_bias.shape).astype(np.float32)
    text = random_dna(L, 1000 + k, "ACGTN" if k % 3 == 0 else "ACGT").strip("N") or "A"
    st, fwd = seq.one_hot_encode_dna_sequence(text)
    n = fwd.shape[1]
    ds = pred.fetch_validation_batch(fwd, step, batch, T)
    got = pred.predict(w, ds, (n, 5), step)
    wd = w.as_dict()
    exp = oracle.predict(lambda b: oracle.model_forward(b, wd), oracle.fetch_validation_batch(fwd, step, batch, T),
                         (n, 5), step)
    assert got.shape == exp.shape
    assert np.abs(got - exp).max() < 3e-5, (gpu_ctx.get_int("forward_used_tc"), float(np.abs(got - exp).max()))
    assert ((got == 0).all(axis=1) == (exp == 0).all(axis=1)).all()          # the same never-covered rows
    # and end to end (fused vote + score, MSS, rows) against the oracle's record driver
    labels, startpos, rows = pred.predict_sequence(w, text.encode(), step, batch, True, 50, 50)
    lab_o, st_o = oracle.predict_record(text, wd, T, batch, step, True)
    assert startpos == st_o and labels.shape == lab_o.shape
    assert (labels == lab_o).mean() >= 0.999
