"""deepgrp_b200.prediction, the parts that need no GPU: the wiring of predict_complete (the reference tests it by
substituting every collaborator, tests/test_prediction.py:95-152) and the metrics arithmetic
(tests/test_prediction.py:155-172 checks it against pycm; scikit-learn is what this image has)."""
import numpy as np
import pytest

from deepgrp_b200 import model, prediction, preprocessing


@pytest.mark.parametrize("step_size", (10, 20))
@pytest.mark.parametrize("use_mss", (True, False))
def test_predict_complete_wiring(monkeypatch, tmp_path, step_size, use_mss):
    opt = model.Options()
    fwd = np.zeros((100, 10))
    data = preprocessing.Data(fwd, np.random.rand(100, 4))
    calls = []

    def fake_setup(options, logdir):
        assert options is opt and logdir == tmp_path
        calls.append("setup")
        return "MODEL"

    def fake_fetch(array, steps, batch_size, vecsize):
        assert array is fwd and steps == step_size
        assert (batch_size, vecsize) == (opt.batch_size, opt.vecsize)
        calls.append("fetch")
        return "DATA"

    def fake_predict(mdl, iterator, output_shape, steps):
        assert (mdl, iterator, output_shape, steps) == ("MODEL", "DATA", (4, 100), step_size)   # truelbl.shape[::-1]
        calls.append("predict")
        return "PREDICTIONS"

    def fake_mss(pred, options):
        assert pred == "PREDICTIONS" and options is opt
        return "MSS"

    def fake_softmax(pred):
        assert pred == "PREDICTIONS"
        return "SOFTMAX"

    monkeypatch.setattr(prediction, "setup_prediction_from_options_checkpoint", fake_setup)
    monkeypatch.setattr(prediction, "fetch_validation_batch", fake_fetch)
    monkeypatch.setattr(prediction, "predict", fake_predict)
    monkeypatch.setattr(prediction, "apply_mss", fake_mss)
    monkeypatch.setattr(prediction, "softmax", fake_softmax)
    got = prediction.predict_complete(step_size=step_size, options=opt, logdir=tmp_path, data=data, use_mss=use_mss)
    assert got == ("MSS" if use_mss else "SOFTMAX")
    assert calls == ["setup", "fetch", "predict"]


@pytest.mark.parametrize("seed", range(5))
def test_metrics_against_scikit_learn(oracle, seed):
    from sklearn import metrics as skm
    rng = np.random.default_rng(seed)
    truth = rng.choice([0, 1, 2, 3], size=400, p=[0.55, 0.2, 0.15, 0.1])
    pred = np.where(rng.random(400) < 0.7, truth, rng.choice([0, 1, 2, 3], size=400))
    cnf = oracle.confusion_matrix(truth, pred)                   # rows = truth, columns = prediction
    assert np.array_equal(cnf, skm.confusion_matrix(truth, pred, labels=[0, 1, 2, 3]))
    m = prediction._calculate_metrics(cnf)
    prec, rec, f1, _ = skm.precision_recall_fscore_support(truth, pred, labels=[0, 1, 2, 3], zero_division=0)
    assert np.allclose(m["TPR"], rec) and np.allclose(m["PPV"], prec) and np.allclose(m["F1"], f1)
    assert np.allclose(m["FNR"], 1 - rec) and np.allclose(m["FDR"], 1 - prec)
    for k in range(4):                                           # one-vs-rest rates
        t, p = truth == k, pred == k
        tn, fp, fn, tp = skm.confusion_matrix(t, p, labels=[False, True]).ravel()
        assert np.isclose(m["TNR"][k], tn / (tn + fp)) and np.isclose(m["FPR"][k], fp / (fp + tn))
        assert np.isclose(m["NPV"][k], tn / (tn + fn)) and np.isclose(m["ACC"][k], (tp + tn) / 400)
    assert np.isclose(m["MCC"], skm.matthews_corrcoef(truth, pred))
    assert np.isclose(prediction.calculate_multiclass_matthews_cc(cnf), skm.matthews_corrcoef(truth, pred))


def test_matthews_cc_limits():
    assert np.isclose(prediction.calculate_multiclass_matthews_cc(np.diag([5, 7, 3])), 1.0)
    assert np.isclose(prediction.calculate_multiclass_matthews_cc(np.array([[0, 5], [5, 0]])), -1.0)
    assert abs(prediction.calculate_multiclass_matthews_cc(np.array([[25, 25], [25, 25]]))) < 1e-12


@pytest.mark.parametrize("step_size", (1, 2))
def test_predict_generic_route_known_answer(monkeypatch, oracle, step_size):
    """The vector of the reference's test_predict (tests/test_prediction.py:74-93): a model that always answers
    class 1 at the first position of each of its 4 windows, three equal batches -> twelve rows [0, 1, 0] at
    multiples of the step.  The generic route of predict (any object with predict_on_batch, any iterable of
    batches) is the reference's loop; the max-merge it calls is the oracle's here, the kernel's on a GPU box."""
    answer = np.zeros((4, 10, 3), dtype=np.float32)
    answer[:, 0, 1] = 1

    class ConstantModel:
        def predict_on_batch(self, batch):
            assert batch.shape == (4, 10, 5)
            return answer
    monkeypatch.setattr(prediction.dgsequence, "get_max", oracle.get_max)
    batches = (np.random.rand(4, 10, 5) for _ in range(3))
    got = prediction.predict(model=ConstantModel(), data=batches, results_shape=(50, 3), step_size=step_size)
    assert got.dtype == np.float32 and got.shape == (50, 3)
    assert got.sum(axis=0).tolist() == [0, 12, 0]
    assert all(got[i * step_size].tolist() == [0, 1, 0] for i in range(12))


def test_predict_takes_the_literal_loop_when_the_steps_differ(monkeypatch, oracle):
    """The reference enumerates windows with the dataset's step and places them with predict's (deepgrp/prediction.py:
    28-37, 103-111).  The fused GPU route uses one step for both, so a mismatch -- or a non-integer matrix -- must go
    through the reference's literal loop (here with the oracle's vote standing in for the GPU call)."""
    import deepgrp_b200.sequence as dgsequence
    T, U = 20, 4
    w = model.random_weights(T, U, attention=True, seed=0)
    rng = np.random.default_rng(0)
    fwd = np.eye(5, dtype=np.int8)[rng.integers(0, 5, size=300)].T.copy()

    def no_gpu(*a, **k):
        raise AssertionError("the fused route must not be taken")
    monkeypatch.setattr(prediction._lib, "context", no_gpu)
    monkeypatch.setattr(dgsequence, "get_max", oracle.get_max)
    monkeypatch.setattr(model.ModelWeights, "predict_on_batch",
                        lambda self, b: oracle.model_forward(b, self.as_dict()))
    ds = prediction.fetch_validation_batch(fwd, 10, 8, T)
    got = prediction.predict(w, ds, (300, 5), 5)                   # placed with step 5, enumerated with 10
    exp = oracle.predict(lambda b: oracle.model_forward(b, w.as_dict()),
                         oracle.fetch_validation_batch(fwd, 10, 8, T), (300, 5), 5)
    assert np.array_equal(got, exp)
    ds_f = prediction.fetch_validation_batch(fwd.astype(np.float32) * 0.5, 10, 8, T)
    got_f = prediction.predict(w, ds_f, (300, 5), 10)              # a float matrix that is not one-hot
    assert got_f.shape == (300, 5) and np.isfinite(got_f).all()
