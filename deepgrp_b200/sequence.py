"""deepgrp_b200.sequence -- drop-in for the reference's Cython module ``deepgrp.sequence``
(``deepgrp/sequence.pyx`` + ``deepgrp/maxcalc.c``; stubs ``deepgrp/sequence.pyi:6-20``).

Same names, argument meaning, dtypes and error behaviour; the work runs in CUDA kernels behind
the C ABI (``csrc/encode.cu``, ``csrc/vote.cu``, ``csrc/segments.cu``).
"""
from __future__ import annotations

import ctypes
from typing import Iterator, Tuple

import numpy as np

from . import _lib


def _one_hot_encode_bytes(raw: bytes, fold_case: bool = False) -> Tuple[int, np.ndarray]:
    ctx = _lib.context()
    n = len(raw)
    buf = np.frombuffer(raw, dtype=np.uint8) if n else np.zeros(0, np.uint8)
    start, out_len = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = _lib.lib().dgrp_one_hot_stage(ctx.handle, _lib.ptr(buf), n, int(fold_case),
                                       ctypes.byref(start), ctypes.byref(out_len))
    if rc == _lib.E_ALLN:
        # np.zeros((5, length - startpos)) with a negative size, deepgrp/sequence.pyx:32
        raise ValueError("negative dimensions are not allowed")
    _lib.check(rc)
    fwd = np.zeros((5, out_len.value), dtype=np.int8)
    if out_len.value:
        _lib.check(_lib.lib().dgrp_one_hot_fetch(ctx.handle, _lib.ptr(fwd)))
    return int(start.value), fwd


def one_hot_encode_dna_sequence(sequence: str) -> Tuple[int, np.ndarray]:
    """One hot encodes sequence, drops leading and trailing N's
    (reference ``deepgrp/sequence.pyx:55-58``).  Returns ``(startpos, int8[5, L])``."""
    if not isinstance(sequence, str):
        raise AttributeError("'%s' object has no attribute 'encode'" % type(sequence).__name__)
    return _one_hot_encode_bytes(sequence.encode("utf-8"))


def _typed(name: str, a, dtype, ndim: int) -> np.ndarray:
    """The checks Cython's typed buffer arguments perform (TypeError for None / non-arrays,
    ValueError for dtype, ndim or contiguity mismatches)."""
    if a is None:
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got NoneType)" % name)
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)"
                        % (name, type(a).__name__))
    if a.dtype != np.dtype(dtype):
        raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'"
                         % (np.dtype(dtype).name, a.dtype.name))
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a.ndim))
    return a


def get_max(output: np.ndarray, inputs: np.ndarray, stride: int) -> np.ndarray:
    """Window ``b`` of ``inputs[B, T, C]`` is merged by elementwise maximum into ``output`` starting
    at row ``b*stride``; in place, returns ``output`` (reference ``deepgrp/sequence.pyx:67-76`` ->
    ``deepgrp/maxcalc.c:10-24``)."""
    _typed("output", output, np.float32, 2)
    _typed("inputs", inputs, np.float32, 3)
    if not output.flags.c_contiguous or not inputs.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")
    b, d0, d1 = inputs.shape
    ctx = _lib.context()
    _lib.check(_lib.lib().dgrp_get_max(ctx.handle, _lib.ptr(output), output.shape[0],
                                       _lib.ptr(inputs), b, d0, d1, int(stride)))
    return output


def get_segments(classes: np.ndarray, startpos: int) -> Tuple[int, int, int]:
    """Start, end and label of the next segment at or after ``startpos``
    (reference ``deepgrp/sequence.pyx:40-53``, including its ``size - 1`` loop bounds)."""
    _typed("classes", classes, np.int64, 1)
    classes = np.ascontiguousarray(classes)
    out = np.zeros(3, dtype=np.int64)
    ctx = _lib.context()
    _lib.check(_lib.lib().dgrp_get_segments(ctx.handle, _lib.ptr(classes), classes.size,
                                            int(startpos), _lib.ptr(out)))
    return int(out[0]), int(out[1]), int(out[2])


def segments_array(classes: np.ndarray, start_offset: int) -> np.ndarray:
    """All ``(start, end, label)`` triples ``yield_segments`` produces, as ``int64[n, 3]``."""
    classes = np.ascontiguousarray(classes, dtype=np.int64)
    if classes.size == 0:
        return np.zeros((0, 3), np.int64)
    ctx = _lib.context()
    cap = 1024
    while True:
        out = np.zeros((cap, 3), dtype=np.int64)
        n_out = ctypes.c_int64(0)
        rc = _lib.lib().dgrp_yield_segments(ctx.handle, _lib.ptr(classes), classes.size,
                                            int(start_offset), _lib.ptr(out), cap,
                                            ctypes.byref(n_out))
        if rc == _lib.E_CAPACITY:
            cap = int(n_out.value)
            continue
        _lib.check(rc)
        return out[:n_out.value]


def yield_segments(classes: np.ndarray, start_offset: int) -> Iterator[Tuple[int, int, int]]:
    """Converts an array of classes to an iterator over continuous segments
    (reference ``deepgrp/sequence.pyx:79-85``)."""
    for start, end, label in segments_array(np.asarray(classes), start_offset):
        yield int(start), int(end), int(label)
