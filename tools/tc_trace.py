"""Where the cycles of the tcgen05 forward go: builds the library with -DDGRP_TC_TRACE (clock64 around
the segments of the gate warps' step loop, CTA 0 only), runs one forward and prints cycles per
tile-step and segment.  `python tools/tc_trace.py build` here (no GPU needed), then on a GPU box
`python tools/tc_trace.py run [bases]`."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "deepgrp_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_trace")
LIB = os.path.join(OUT, "libdeepgrp_b200_trace.so")
SOURCES = ["api.cu", "encode.cu", "vote.cu", "forward.cu", "forward_tc.cu", "forward_tcw.cu", "mss.cu", "segments.cu", "fasta.cu", "tsv.cu", "evaluate.cu"]


def build():
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for src in SOURCES:
        o = os.path.join(OUT, src[:-3] + ".o")
        subprocess.run(["nvcc", "-O3", "-std=c++17", "-lineinfo", "-DDGRP_TC_TRACE", "-gencode",
                        "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-c",
                        os.path.join(CSRC, src), "-o", o], check=True)
        objs.append(o)
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lcudart"],
                   check=True)


def run(bases):
    os.environ["DEEPGRP_B200_LIB"] = LIB
    sys.path.insert(0, ROOT)
    import numpy as np
    from deepgrp_b200 import _lib, model, prediction, sequence
    ctx = _lib.context()
    T, U = 342, 60
    w = model.random_weights(T, U, attention=True, seed=0)
    rng = np.random.default_rng(1)
    codes = rng.integers(0, 4, size=bases, dtype=np.uint8)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].tobytes().decode()
    st, fwd = sequence.one_hot_encode_dna_sequence(seq)
    ds = prediction.fetch_validation_batch(fwd, 50, 256, T)
    lib = _lib.lib()
    prediction.predict(w, ds, (bases, 5), 50)          # warm-up
    lib.dgrp_debug_tc_trace(None, 1)
    lib.dgrp_debug_tc_trace2(None, 1)
    prediction.predict(w, ds, (bases, 5), 50)
    buf = (ctypes.c_ulonglong * (17 * 2 * 8))()
    lib.dgrp_debug_tc_trace(buf, 0)
    tr = np.array(buf[:], dtype=np.float64).reshape(17, 2, 8)
    n_windows = len(range(0, bases - T, 50))
    n_tiles = (n_windows + 63) // 64
    tiles0 = n_tiles * 1 // 148 - 0            # CTA 0's tile range is [0, n_tiles/148)
    steps = tiles0 * T / 1.0                   # tile-steps of CTA 0 (each slot does about half)
    print("windows %d, tiles of CTA 0: %d" % (n_windows, tiles0))
    names = ["wait done", "tmem loads", "gates+A stores", "fence+arrive", "loop top (prefetch)", "proj shfl+store", "sum shfl+store"]
    g = tr[:16]
    per = g[:, :, :7].sum(axis=1) / steps      # cycles per tile-step, per warp
    print("cycles per tile-step (mean over the 16 gate warps; min..max):")
    for k, name in enumerate(names):
        print("  %-18s %7.0f   (%5.0f .. %5.0f)" % (name, per[:, k].mean(), per[:, k].min(), per[:, k].max()))
    print("  %-18s %7.0f" % ("sum", per.sum(axis=1).mean()))
    buf2 = (ctypes.c_ulonglong * (16 * 8))()
    lib.dgrp_debug_tc_trace2(buf2, 0)
    tr2 = np.array(buf2[:], dtype=np.float64).reshape(16, 8) / tiles0
    print("second phase, cycles per tile (mean over the 16 warps; the recurrence of a tile is %d x the tile-step):" % T)
    for k, name in enumerate(["scores", "barrier", "softmax over t", "logits + vote", "barrier"]):
        print("  %-18s %8.0f   (%6.0f .. %6.0f)" % (name, tr2[:, k].mean(), tr2[:, k].min(), tr2[:, k].max()))
    iss = tr[16, :, 5:7].sum(axis=0) / steps
    print("issuer per tile-step: waiting for ready %.0f, issuing %.0f" % (iss[0], iss[1]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 23_350_000)
