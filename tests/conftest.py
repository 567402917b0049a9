"""Shared fixtures.  `-m "not gpu"` tests run on the CPU box (oracle vs golden vectors, host logic,
library symbols); `-m gpu` tests are the parity tests proper and call through the C ABI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # the shared library and the checker's natives are build products (git-ignored): build them when
    # a fresh checkout runs the tests before __graft_entry__.build() has been called
    lib = os.path.join(ROOT, "deepgrp_b200", "libdeepgrp_b200.so")
    if not os.path.exists(lib) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build_c()
    return o


@pytest.fixture(scope="session")
def gpu_ctx():
    from deepgrp_b200 import _lib
    if _lib.device_count() == 0:
        pytest.fail("no CUDA device visible: -m gpu tests need a B200 (there is no CPU fallback)")
    return _lib.context()


def random_dna(n, seed, alphabet="ACGT"):
    rng = np.random.default_rng(seed)
    return "".join(np.array(list(alphabet))[rng.integers(0, len(alphabet), size=n)])


def write_fasta(path, records, width=60):
    with open(path, "w") as fh:
        for header, seq in records:
            fh.write(">" + header + "\n")
            for i in range(0, len(seq), width):
                fh.write(seq[i:i + width] + "\n")


def messy_fasta(seed):
    """FASTA text with mixed line ends (LF, CRLF, lone CR), whitespace around and inside lines, headers with
    blanks, ragged line widths, text before the first '>', an empty-header record, an optional missing final
    line end and (sometimes) a blank line."""
    import random
    r = random.Random(seed)
    parts = []
    if r.random() < 0.3:
        parts.append("ACGT" * r.randint(1, 50) + r.choice(["\n", "\r\n"]))
    for i in range(r.randint(1, 4)):
        nl = r.choice(["\n", "\r\n", "\r", "\n"])
        hdr = r.choice(["chr%d desc" % i, "x", "", " spaced\t", "h>h"])
        parts.append(r.choice(["", " ", "\t"]) + ">" + hdr + r.choice(["", " ", "\t "]) + nl)
        for _ in range(r.randint(20, 400)):
            w = r.choice([60, 60, 60, 1, 7, 80, 16, 15, 17])
            line = "".join(r.choice("ACGTacgtNn") for _ in range(w))
            if r.random() < 0.05:
                line = line[:w // 2] + r.choice([" ", "\t", "  "]) + line[w // 2:]
            if r.random() < 0.05:
                line = r.choice([" ", "\t"]) + line
            if r.random() < 0.05:
                line = line + r.choice([" ", "\t", " \t "])
            parts.append(line + nl)
    text = "".join(parts)
    if r.random() < 0.3:
        text = text.rstrip("\r\n")
    if r.random() < 0.1:
        text += r.choice(["\n\n", "\n \n", " ", "\n\t"])
    return text
