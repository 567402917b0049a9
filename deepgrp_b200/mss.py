"""deepgrp_b200.mss -- drop-in for the reference's Cython module ``deepgrp.mss``
(``deepgrp/_mss/pymss.pyx`` + ``deepgrp/_mss/mss.c``; stub ``deepgrp/mss.pyi:4-6``)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .sequence import _typed


def find_mss_labels(inputs: np.ndarray, label: np.ndarray, nof_labels: int, min_mss_len: int,
                    xdrop_len: int) -> np.ndarray:
    """Maximal scoring segments of ``inputs`` with the majority non-zero label filled into the
    label-0 positions of every segment; one-hot ``float64[n, nof_labels]``
    (reference ``deepgrp/_mss/pymss.pyx:16-27``)."""
    _typed("inputs", inputs, np.float64, 1)
    _typed("label", label, np.int64, 1)
    inputs = np.ascontiguousarray(inputs)
    label = np.ascontiguousarray(label)
    n = int(inputs.shape[0])
    one_hot = np.zeros((n, int(nof_labels)), dtype=np.float64)
    if n == 0:
        return one_hot
    ctx = _lib.context()
    _lib.check(_lib.lib().dgrp_find_mss_labels(ctx.handle, _lib.ptr(inputs), _lib.ptr(label), n,
                                               int(nof_labels), int(min_mss_len), int(xdrop_len),
                                               _lib.ptr(one_hot)))
    return one_hot


def mss_find_all(scores: np.ndarray, min_sc: float, xdrop: float) -> np.ndarray:
    """``mss_find_all`` (reference ``deepgrp/_mss/mss.c:50-101``): structured array (st, en, sc)."""
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    n = int(scores.size)
    if n == 0:
        return np.zeros(0, dtype=_lib.SEG_DTYPE)
    ctx = _lib.context()
    cap = max(16, n // 2 + 1)
    out = np.zeros(cap, dtype=_lib.SEG_DTYPE)
    n_seg = ctypes.c_int(0)
    _lib.check(_lib.lib().dgrp_mss_find_all(ctx.handle, n, _lib.ptr(scores), float(min_sc),
                                            float(xdrop), _lib.ptr(out), cap, ctypes.byref(n_seg)))
    return out[:n_seg.value]
