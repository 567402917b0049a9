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
