#!/usr/bin/env python
"""CPU model of the forward kernel's split-precision recurrent product (forward_tc.cu, NP = 2): state x 2^8 and
weights x 2^shift as two fp16 pieces each, the three products hi.hi + hi.lo + lo.hi accumulated in float32 -- run
through the whole GRU recurrence next to a float64 run and a plain float32 run of the same weights.  A design
input for shapes the tcgen05 kernel does not cover yet (units > 64: K = 128 + 16, T = 512): does the 3-product
form still sit at the float32 kernel's distance from float64?

    python tools/split_precision_model.py [--units 128 --vecsize 512 --windows 8 --scale 1]

No GPU code is exercised: this is numpy arithmetic on the oracle's equations (oracle/oracle.py: gru_sequence)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def split16(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def gru(x, w, mode):
    """mode: 'f64', 'f32' or 'fp16x2' (the recurrent product only; gates and state update in float32)."""
    dt = np.float64 if mode == "f64" else np.float32
    kern, rec = w["kernel"].astype(dt), w["recurrent_kernel"].astype(dt)
    b_in, b_rec = w["bias"][0].astype(dt), w["bias"][1].astype(dt)
    U = rec.shape[0]
    if mode == "fp16x2":
        e = np.frexp(np.abs(rec).max())[1]
        shift = 14 - int(e)
        r_hi, r_lo = split16(np.ldexp(rec, shift).astype(np.float32))
        unscale = np.float32(np.ldexp(1.0, -(8 + shift)))
    h = np.zeros((x.shape[0], U), dtype=dt)
    seq = np.empty((x.shape[0], x.shape[1], U), dtype=dt)
    for t in range(x.shape[1]):
        mx = x[:, t, :].astype(dt) @ kern + b_in
        if mode == "fp16x2":
            h_hi, h_lo = split16(np.ldexp(h, 8).astype(np.float32))
            acc = (h_lo @ r_hi) + (h_hi @ r_lo)                  # smallest terms first, float32 accumulate
            acc = acc + (h_hi @ r_hi)
            mh = acc * unscale + b_rec
        else:
            mh = h @ rec + b_rec
        z = 1 / (1 + np.exp(-(mx[:, :U] + mh[:, :U])))
        r = 1 / (1 + np.exp(-(mx[:, U:2 * U] + mh[:, U:2 * U])))
        hh = np.tanh(mx[:, 2 * U:] + r * mh[:, 2 * U:])
        h = (z * h + (1 - z) * hh).astype(dt)
        seq[:, t, :] = h
    return seq


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=128)
    ap.add_argument("--vecsize", type=int, default=512)
    ap.add_argument("--windows", type=int, default=8)
    ap.add_argument("--scale", type=float, default=1.0, help="weight scale (4 = the sharp-margin regime of the tests)")
    a = ap.parse_args()
    from deepgrp_b200 import model
    w = model.random_weights(a.vecsize, a.units, attention=True, seed=0).scaled(a.scale).as_dict()
    rng = np.random.default_rng(1)
    x = np.eye(5, dtype=np.float32)[rng.integers(0, 4, size=(a.windows, a.vecsize))]
    ref = gru(x, w, "f64")
    for mode in ("f32", "fp16x2"):
        got = gru(x, w, mode)
        err = np.abs(got - ref)
        print("%-7s max |dh| %.3e  mean |dh| %.3e  (last step: max %.3e)"
              % (mode, err.max(), err.mean(), err[:, -1].max()))


if __name__ == "__main__":
    main()
