"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of DeepGRP's prediction path (`deepgrp predict`), used only as the checker by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing
under deepgrp_b200/ imports this module.

Parity status
-------------
* Integer / byte / double parts (encode, max-vote, segments, MSS + gap fill): PINNED against the
  reference's own compiled natives (oracle/_ref, see oracle/build_ref.py), against the
  reference's known-answer tests (tests/test_sequence.py, tests/test_mss.py) and against golden
  vectors generated from those natives (tests/golden/, made by tests/golden/make_golden.py).
* GRU / attention / Dense / softmax numerics: **parity unpinned**.  The arithmetic lives in
  TensorFlow 2.5.0 / keras-nightly 2.5.0.dev2021032900 (poetry.lock:1005-1006, 501-502), which is
  not in /root/reference and not installable here; the reference ships no weights and no test
  that checks an output value.  The restatement below follows the published Keras 2.5 equations
  (`standard_gru` step with reset_after=True; `AdditiveAttention._calculate_scores`) anchored on
  the call sites deepgrp/model.py:293-336 and the layer graph in tests/test_model.json, and is
  cross-checked between two independent engines (numpy step loop vs torch.nn.GRU) and against a
  float64 run in tests/test_oracle.py.

All `file:line` citations are relative to /root/reference.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[ctypes.CDLL] = None


# --------------------------------------------------------------------------------------------
# C restatement loader (oracle/oracle_c.c)
# --------------------------------------------------------------------------------------------
def build_c(force: bool = False) -> str:
    """Compile oracle_c.c into oracle/liboracle.so with gcc (no reference sources involved)."""
    src = os.path.join(HERE, "oracle_c.c")
    out = os.path.join(HERE, "liboracle.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", out, src, "-lm"], check=True)
    return out


class _Seg(ctypes.Structure):
    _fields_ = [("st", ctypes.c_int), ("en", ctypes.c_int), ("sc", ctypes.c_double)]


def clib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c())
        lib.orc_one_hot_encode.restype = ctypes.c_long
        lib.orc_one_hot_encode.argtypes = [ctypes.c_void_p, ctypes.c_long,
                                           ctypes.POINTER(ctypes.c_long), ctypes.c_void_p]
        lib.orc_get_max.restype = None
        lib.orc_get_max.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_size_t] * 4
        lib.orc_get_segments.restype = None
        lib.orc_get_segments.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long,
                                         ctypes.c_void_p]
        lib.orc_yield_segments.restype = ctypes.c_long
        lib.orc_yield_segments.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long,
                                           ctypes.c_void_p]
        lib.orc_mss_find_all.restype = ctypes.POINTER(_Seg)
        lib.orc_mss_find_all.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_double,
                                         ctypes.c_double, ctypes.POINTER(ctypes.c_int)]
        lib.orc_free.restype = None
        lib.orc_free.argtypes = [ctypes.c_void_p]
        lib.orc_find_mss_labels.restype = None
        lib.orc_find_mss_labels.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_int]
        lib.orc_find_mss_relabel.restype = None
        lib.orc_find_mss_relabel.argtypes = lib.orc_find_mss_labels.argtypes
        _LIB = lib
    return _LIB


def _ptr(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


# --------------------------------------------------------------------------------------------
# deepgrp.sequence
# --------------------------------------------------------------------------------------------
def one_hot_encode_bytes(raw: bytes) -> Tuple[int, np.ndarray]:
    """deepgrp/sequence.pyx:21-36 on raw bytes."""
    buf = np.frombuffer(raw, dtype=np.uint8) if len(raw) else np.zeros(0, np.uint8)
    start = ctypes.c_long(0)
    out_len = clib().orc_one_hot_encode(_ptr(buf), len(raw), ctypes.byref(start), None)
    if out_len < 0:
        # np.zeros((5, negative)) in the reference (sequence.pyx:32)
        raise ValueError("negative dimensions are not allowed")
    fwd = np.zeros((5, out_len), dtype=np.int8)
    if out_len:
        clib().orc_one_hot_encode(_ptr(buf), len(raw), ctypes.byref(start), _ptr(fwd))
    return int(start.value), fwd


def one_hot_encode_dna_sequence(sequence: str) -> Tuple[int, np.ndarray]:
    """deepgrp/sequence.pyx:55-58."""
    return one_hot_encode_bytes(sequence.encode("utf-8"))


def get_max(output: np.ndarray, inputs: np.ndarray, stride: int) -> np.ndarray:
    """deepgrp/sequence.pyx:67-76 -> deepgrp/maxcalc.c:10-24 (in place, returns `output`)."""
    assert output.dtype == np.float32 and inputs.dtype == np.float32
    assert output.flags.c_contiguous and inputs.flags.c_contiguous
    clib().orc_get_max(_ptr(output), _ptr(inputs), inputs.shape[1], inputs.shape[2], stride,
                       inputs.shape[0])
    return output


def get_segments(classes: np.ndarray, startpos: int) -> Tuple[int, int, int]:
    """deepgrp/sequence.pyx:40-53."""
    classes = np.ascontiguousarray(classes, dtype=np.int64)
    out = np.zeros(3, dtype=np.int64)
    clib().orc_get_segments(_ptr(classes), classes.size, startpos, _ptr(out))
    return int(out[0]), int(out[1]), int(out[2])


def yield_segments_array(classes: np.ndarray, start_offset: int) -> np.ndarray:
    """deepgrp/sequence.pyx:79-85, materialised as int64[n,3] (all labels, including 0)."""
    classes = np.ascontiguousarray(classes, dtype=np.int64)
    out = np.zeros((max(classes.size, 1), 3), dtype=np.int64)
    n = clib().orc_yield_segments(_ptr(classes), classes.size, start_offset, _ptr(out))
    return out[:n].copy()


def yield_segments(classes: np.ndarray, start_offset: int) -> Iterator[Tuple[int, int, int]]:
    for s, e, l in yield_segments_array(classes, start_offset):
        yield int(s), int(e), int(l)


# --------------------------------------------------------------------------------------------
# deepgrp.mss
# --------------------------------------------------------------------------------------------
def mss_find_all(scores: np.ndarray, min_sc: float, xdrop: float) -> np.ndarray:
    """deepgrp/_mss/mss.c:50-101. Returns a structured array (st, en, sc)."""
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    n = ctypes.c_int(0)
    p = clib().orc_mss_find_all(scores.size, _ptr(scores), min_sc, xdrop, ctypes.byref(n))
    out = np.zeros(n.value, dtype=[("st", np.int32), ("en", np.int32), ("sc", np.float64)])
    for i in range(n.value):
        out[i] = (p[i].st, p[i].en, p[i].sc)
    clib().orc_free(p)
    return out


def find_mss_labels(inputs: np.ndarray, label: np.ndarray, nof_labels: int, min_mss_len: int,
                    xdrop_len: int) -> np.ndarray:
    """deepgrp/_mss/pymss.pyx:16-27."""
    inputs = np.ascontiguousarray(inputs, dtype=np.float64)
    label = np.ascontiguousarray(label, dtype=np.int64)
    out = np.zeros((inputs.size, nof_labels), dtype=np.float64)
    clib().orc_find_mss_labels(_ptr(inputs), _ptr(label), nof_labels, min_mss_len, xdrop_len,
                               _ptr(out), inputs.size)
    return out


def find_mss_relabel(inputs: np.ndarray, label: np.ndarray, nof_labels: int, min_mss_len: int,
                     xdrop_len: int) -> np.ndarray:
    """argmax(find_mss_labels(...), axis=1) as uint8 without the [n, nof] matrix."""
    inputs = np.ascontiguousarray(inputs, dtype=np.float64)
    label = np.ascontiguousarray(label, dtype=np.int64)
    out = np.zeros(inputs.size, dtype=np.uint8)
    clib().orc_find_mss_relabel(_ptr(inputs), _ptr(label), nof_labels, min_mss_len, xdrop_len,
                                _ptr(out), inputs.size)
    return out


# --------------------------------------------------------------------------------------------
# Model forward (Keras 2.5 equations; call sites deepgrp/model.py:293-336)
# --------------------------------------------------------------------------------------------
COMPLEMENT = [3, 2, 1, 0, 4]                       # deepgrp/model.py:233-237


def reverse_complement(x: np.ndarray) -> np.ndarray:
    """deepgrp/model.py:277-279: gather(reverse(x, axis=1), [3,2,1,0,4], axis=2)."""
    return x[:, ::-1, :][:, :, COMPLEMENT]


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gru_sequence(x: np.ndarray, w: Dict[str, np.ndarray], dtype=np.float32):
    """Keras GRU(reset_after=True) over x[B,T,5] from a zero state (deepgrp/model.py:225-229;
    gate order z, r, h; bias[0] = input bias, bias[1] = recurrent bias).
    Returns (sequence [B,T,U], last state [B,U])."""
    kern = w["kernel"].astype(dtype)
    rec = w["recurrent_kernel"].astype(dtype)
    b_in = w["bias"][0].astype(dtype)
    b_rec = w["bias"][1].astype(dtype)
    units = rec.shape[0]
    nb, nt, _ = x.shape
    h = np.zeros((nb, units), dtype=dtype)
    seq = np.empty((nb, nt, units), dtype=dtype)
    xs = x.astype(dtype)
    for t in range(nt):
        mx = xs[:, t, :] @ kern + b_in
        mh = h @ rec + b_rec
        z = _sigmoid(mx[:, :units] + mh[:, :units])
        r = _sigmoid(mx[:, units:2 * units] + mh[:, units:2 * units])
        hh = np.tanh(mx[:, 2 * units:] + r * mh[:, 2 * units:])
        h = z * h + (1 - z) * hh
        seq[:, t, :] = h
    return seq, h


def gru_sequence_torch(x: np.ndarray, w: Dict[str, np.ndarray], threads: int = 0):
    """Second, independent engine: torch.nn.GRU on CPU (gate order r,z,n; h=(1-z)n+zh, the same
    reset_after form).  Keras column blocks [z,r,h] are permuted to [r,z,n] and transposed."""
    import torch
    if threads > 0:
        torch.set_num_threads(threads)
    units = w["recurrent_kernel"].shape[0]

    def perm(m):                       # [*, 3U] keras z,r,h -> torch r,z,n  (rows after .T)
        z, r, hh = m[..., :units], m[..., units:2 * units], m[..., 2 * units:]
        return np.concatenate([r, z, hh], axis=-1)

    gru = torch.nn.GRU(5, units, batch_first=True)
    with torch.no_grad():
        gru.weight_ih_l0.copy_(torch.from_numpy(perm(w["kernel"]).T.copy()))
        gru.weight_hh_l0.copy_(torch.from_numpy(perm(w["recurrent_kernel"]).T.copy()))
        gru.bias_ih_l0.copy_(torch.from_numpy(perm(w["bias"][0]).copy()))
        gru.bias_hh_l0.copy_(torch.from_numpy(perm(w["bias"][1]).copy()))
        seq, last = gru(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)))
    return seq.numpy(), last[0].numpy()


def lstm_sequence(x: np.ndarray, w: Dict[str, np.ndarray], dtype=np.float32):
    """Keras LSTM (implementation 2, gate order i, f, c, o; one bias) over x[B,T,5] from zero states
    (deepgrp/model.py:219-223): z = x.W + h.R + b; c' = sig(z_f)*c + sig(z_i)*tanh(z_c);
    h' = sig(z_o)*tanh(c').  Returns (sequence [B,T,U], last h)."""
    kern = w["kernel"].astype(dtype)
    rec = w["recurrent_kernel"].astype(dtype)
    bias = w["bias"].reshape(-1).astype(dtype)
    units = rec.shape[0]
    nb, nt, _ = x.shape
    h = np.zeros((nb, units), dtype=dtype)
    c = np.zeros((nb, units), dtype=dtype)
    seq = np.empty((nb, nt, units), dtype=dtype)
    xs = x.astype(dtype)
    for t in range(nt):
        z = xs[:, t, :] @ kern + h @ rec + bias
        i = _sigmoid(z[:, :units])
        f = _sigmoid(z[:, units:2 * units])
        c = f * c + i * np.tanh(z[:, 2 * units:3 * units])
        o = _sigmoid(z[:, 3 * units:])
        h = o * np.tanh(c)
        seq[:, t, :] = h
    return seq, h


def lstm_sequence_torch(x: np.ndarray, w: Dict[str, np.ndarray]):
    """Second engine: torch.nn.LSTM on CPU (gate order i, f, g, o = Keras' i, f, c, o)."""
    import torch
    units = w["recurrent_kernel"].shape[0]
    lstm = torch.nn.LSTM(5, units, batch_first=True)
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(torch.from_numpy(w["kernel"].T.copy()))
        lstm.weight_hh_l0.copy_(torch.from_numpy(w["recurrent_kernel"].T.copy()))
        lstm.bias_ih_l0.copy_(torch.from_numpy(w["bias"].reshape(-1).copy()))
        lstm.bias_hh_l0.zero_()
        seq, (last, _) = lstm(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)))
    return seq.numpy(), last[0].numpy()


def is_lstm(w: Dict[str, np.ndarray]) -> bool:
    return w["kernel"].shape[1] == 4 * w["recurrent_kernel"].shape[0]


def model_forward(batch: np.ndarray, w: Dict[str, np.ndarray], dtype=np.float32,
                  engine: str = "numpy") -> np.ndarray:
    """The functional graph of deepgrp/model.py:293-336 on one batch float[B,T,5] -> [B,T,C].

    `w` keys: kernel[5,3U], recurrent_kernel[U,3U], bias[2,3U], ff_kernel[F,C], ff_bias[C] and,
    when the model has attention, att_scale[U] (F = 2U with attention, U without)."""
    x = batch.astype(dtype)
    x_rc = np.ascontiguousarray(reverse_complement(x))
    if is_lstm(w):                                         # rnn="LSTM": attention is ignored (model.py:308)
        run = lstm_sequence_torch if engine == "torch" else (lambda a, b: lstm_sequence(a, b, dtype))
        fwd, _ = run(x, w)
        rev, _ = run(x_rc, w)
        hf = hr = None
    elif engine == "torch":
        fwd, hf = gru_sequence_torch(x, w)
        rev, hr = gru_sequence_torch(x_rc, w)
    else:
        fwd, hf = gru_sequence(x, w, dtype)
        rev, hr = gru_sequence(x_rc, w, dtype)
    half = dtype(0.5) if dtype is not np.float64 else 0.5
    avg = (fwd + rev) * half                               # Average (model.py:312), not re-reversed
    if not is_lstm(w) and "att_scale" in w and w["att_scale"] is not None:
        hidden = (hf + hr) * half                          # model.py:311
        scale = w["att_scale"].astype(dtype)
        # AdditiveAttention(use_scale=True): reduce_sum(scale * tanh(q + k), -1)
        scores = np.einsum("btu,u->bt", np.tanh(hidden[:, None, :] + avg), scale)
        scores = scores - scores.max(axis=1, keepdims=True)
        e = np.exp(scores)
        att = e / e.sum(axis=1, keepdims=True)             # softmax over the value axis
        ctx = np.einsum("bt,btu->bu", att, avg)            # matmul(weights, value)
        feat = np.concatenate([np.repeat(ctx[:, None, :], avg.shape[1], axis=1), avg], axis=2)
    else:
        feat = avg                                         # model.py:321-323
    logits = feat @ w["ff_kernel"].astype(dtype) + w["ff_bias"].astype(dtype)   # model.py:325
    logits = logits - logits.max(axis=2, keepdims=True)
    e = np.exp(logits)
    return (e / e.sum(axis=2, keepdims=True)).astype(dtype)                     # model.py:329


# --------------------------------------------------------------------------------------------
# deepgrp.prediction
# --------------------------------------------------------------------------------------------
def window_starts(length: int, vecsize: int, step_size: int) -> range:
    """deepgrp/prediction.py:31: range(0, L - T, step) -- end EXCLUSIVE."""
    return range(0, length - vecsize, step_size)


def fetch_validation_batch(data: np.ndarray, step_size: int, batch_size: int,
                           vecsize: int) -> Iterator[np.ndarray]:
    """deepgrp/prediction.py:14-37 as a plain generator of float32[<=B, T, C] batches
    (tf.data's batch() emits a short final batch)."""
    data_t = data.T
    buf: List[np.ndarray] = []
    for index in window_starts(data_t.shape[0], vecsize, step_size):
        buf.append(data_t[index:index + vecsize].astype("float32"))
        if len(buf) == batch_size:
            yield np.stack(buf)
            buf = []
    if buf:
        yield np.stack(buf)


def predict(forward, data: Iterator[np.ndarray], results_shape: Tuple[int, int],
            step_size: int) -> np.ndarray:
    """deepgrp/prediction.py:89-111, INCLUDING line 105: index = i * batch.shape[0] * step_size
    uses the current batch's size, so a short final batch is max-merged at the wrong offset."""
    predictions = np.zeros(results_shape, dtype=np.float32)
    for i, batch in enumerate(data):
        index = i * batch.shape[0] * step_size
        probas = np.ascontiguousarray(forward(batch), dtype=np.float32)
        get_max(predictions[index:], probas, step_size)
    return predictions


def apply_mss_scores(probs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """deepgrp/prediction.py:51-57: (float64 scores, int64 classes)."""
    results_classes = probs.argmax(axis=1)
    mins = probs.max(axis=1) + 1e-6
    mins[mins > 0.99] = 0.99
    t_scores = np.log(mins / (1 - mins))
    scores = np.where(results_classes > 0, t_scores, -10 * t_scores).astype(float)
    return scores, results_classes


def apply_mss(probs: np.ndarray, min_mss_len: int, xdrop_len: int) -> np.ndarray:
    """deepgrp/prediction.py:40-59."""
    scores, classes = apply_mss_scores(probs)
    return find_mss_labels(scores, classes, probs.shape[1], min_mss_len, xdrop_len)


def softmax(array: np.ndarray) -> np.ndarray:
    """deepgrp/prediction.py:62-65 (global max, row sums)."""
    e_x = np.exp(array - np.max(array))
    return e_x / e_x.sum(axis=1, keepdims=True)


# --------------------------------------------------------------------------------------------
# deepgrp.__main__
# --------------------------------------------------------------------------------------------
def read_multi_fasta(filestream) -> Iterator[Tuple[str, str]]:
    """deepgrp/__main__.py:20-43: strip each line; '>' starts a record (header = rest of line);
    other lines are upper-cased and joined; records with an empty header are dropped; a blank
    line raises IndexError (line[0] on '')."""
    header = ""
    parts: List[str] = []
    for line in filestream:
        line = line.strip()
        if line[0] == ">":
            if header:
                yield header, "".join(parts)
            header = line[1:]
            parts = []
        else:
            parts.append(line.upper())
    if header:
        yield header, "".join(parts)


def predict_record(dnasequence: str, w: Dict[str, np.ndarray], vecsize: int, batch_size: int,
                   step_size: int, use_mss: bool, min_mss_len: int = 50, xdrop_len: int = 50,
                   engine: str = "numpy", dtype=np.float32, return_probs: bool = False):
    """deepgrp/__main__.py:46-83 (_predict): returns (labels int64[L], startpos)."""
    start_pos, inputs = one_hot_encode_dna_sequence(dnasequence)
    it = fetch_validation_batch(inputs, step_size, batch_size, vecsize)
    n_classes = w["ff_kernel"].shape[1]
    prediction = predict(lambda b: model_forward(b, w, dtype=dtype, engine=engine), it,
                         (inputs.shape[1], n_classes), step_size)
    probs = prediction
    if use_mss:
        prediction = apply_mss(prediction, min_mss_len, xdrop_len)
    else:
        prediction = softmax(prediction)
    labels = np.asanyarray(prediction.argmax(axis=1))
    if return_probs:
        return labels, start_pos, probs
    return labels, start_pos


def predict_fasta_tsv(filename: str, w: Dict[str, np.ndarray], vecsize: int, batch_size: int = 256,
                      step_size: int = 50, use_mss: bool = True, min_mss_len: int = 50,
                      xdrop_len: int = 50, engine: str = "numpy") -> str:
    """deepgrp/__main__.py:275-292: the TSV text `deepgrp predict` writes for one FASTA file."""
    rows: List[str] = []
    with open(filename, "r") as fh:
        for header, seq in read_multi_fasta(fh):
            labels, startpos = predict_record(seq, w, vecsize, batch_size, step_size, use_mss,
                                              min_mss_len, xdrop_len, engine=engine)
            for s, e, l in yield_segments(labels, startpos):
                if l > 0:
                    rows.append("{}\t{}\t{}\t{}\t{}\n".format(filename, header, s, e, l))
    return "".join(rows)


# ---- evaluation helpers (deepgrp/prediction.py:200-260) ---------------------------------------------

def confusion_matrix(truelbl, predictedlbl):
    """deepgrp/prediction.py:200-218, the literal loop."""
    truelbl, predictedlbl = np.asarray(truelbl), np.asarray(predictedlbl)
    assert truelbl.size == predictedlbl.size
    n_classes = max(truelbl.max(), predictedlbl.max()) - min(truelbl.min(), predictedlbl.min()) + 1
    cnf = np.zeros((n_classes, n_classes), dtype=int)
    for i, j in zip(truelbl, predictedlbl):
        cnf[i, j] += 1
    return cnf


def filter_segments(array, min_len=50):
    """deepgrp/prediction.py:242-260, the literal loop (in place)."""
    indices = np.where(array > 0)[0]
    next_idx = 0
    for idx in indices:
        if next_idx > idx:
            continue
        next_idx = idx + 1
        found = 1
        while next_idx < array.size and array[next_idx] == array[idx]:
            found += 1
            next_idx += 1
        if found < min_len:
            array[idx:next_idx] = 0

