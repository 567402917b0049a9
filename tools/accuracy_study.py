"""Accuracy of the two forward kernels against the oracle (float32 and float64 restatements):
max |dp| on the max-voted probabilities and label agreement, for random-init and x4-scaled weights.
Run on a GPU box:  python tools/accuracy_study.py [bases]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402
from deepgrp_b200 import _lib, model, prediction, sequence  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    ctx = _lib.context()
    rng = np.random.default_rng(1)
    seq = "".join(np.array(list("ACGT"))[rng.integers(0, 4, size=n)])
    st, fwd = sequence.one_hot_encode_dna_sequence(seq)
    for (T, U) in ((150, 32), (342, 60)):
        for scale in (1.0, 4.0):
            w = model.random_weights(T, U, attention=True, seed=0).scaled(scale)
            ds = prediction.fetch_validation_batch(fwd, 50, 256, T)
            out = {}
            for name, flag in (("tc", 1), ("ffma", 0)):
                ctx.set_int("forward_tc", flag)
                t0 = time.perf_counter()
                out[name] = prediction.predict(w, ds, (n, 5), 50)
                out[name + "_s"] = time.perf_counter() - t0
                out[name + "_used_tc"] = ctx.get_int("forward_used_tc")
            ctx.set_int("forward_tc", 1)
            wd = w.as_dict()
            ref32 = orc.predict(lambda b: orc.model_forward(b, wd, dtype=np.float32, engine="torch"),
                                orc.fetch_validation_batch(fwd, 50, 256, T), (n, 5), 50)
            ref64 = orc.predict(lambda b: orc.model_forward(b, wd, dtype=np.float64).astype(np.float32),
                                orc.fetch_validation_batch(fwd, 50, 256, T), (n, 5), 50)
            lab = {k: v.argmax(axis=1) for k, v in (("tc", out["tc"]), ("ffma", out["ffma"]),
                                                    ("ref32", ref32), ("ref64", ref64))}
            print("T=%d U=%d scale=%g  used_tc=%d/%d" % (T, U, scale, out["tc_used_tc"], out["ffma_used_tc"]))
            for a, b in (("tc", "ref64"), ("ffma", "ref64"), ("ref32", "ref64"), ("tc", "ref32"),
                         ("ffma", "ref32"), ("tc", "ffma")):
                pa = out[a] if a in out else (ref32 if a == "ref32" else ref64)
                pb = out[b] if b in out else (ref32 if b == "ref32" else ref64)
                print("   %-5s vs %-5s  max|dp| %.3e  mean|dp| %.3e  labels differ %d of %d (%.5f%%)"
                      % (a, b, np.abs(pa - pb).max(), np.abs(pa - pb).mean(),
                         int((lab[a] != lab[b]).sum()), n, 100.0 * (lab[a] != lab[b]).mean()))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
