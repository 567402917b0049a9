"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares
(no compute without a GPU), the chunked MSS kernel logic (compiled for the host from the product
header) is bit-exact against the oracle, Options / weights / CLI parsing / window descriptors."""
import ctypes
import io
import os
import re
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_library_exports_every_declared_symbol():
    from deepgrp_b200 import _lib
    header = open(os.path.join(ROOT, "include", "deepgrp_b200.h")).read()
    declared = set(re.findall(r"\b(dgrp_[a-z0-9_]+)\s*\(", header))
    declared -= {"dgrp_ctx", "dgrp_model"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name)
    assert handle.dgrp_version() >= 100


def test_no_gpu_is_an_error_not_a_fallback():
    from deepgrp_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.DeepgrpError) as e:
        _lib.Context(0)
    assert e.value.code == _lib.E_NOGPU
    import deepgrp_b200.sequence as seq
    with pytest.raises(_lib.DeepgrpError):
        seq.one_hot_encode_dna_sequence("ACGT")


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "deepgrp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text and "oracle_c" not in text, f


# ---- chunked MSS kernel logic on the host ---------------------------------------------------------
class _Seg(ctypes.Structure):
    _fields_ = [("st", ctypes.c_int), ("en", ctypes.c_int), ("sc", ctypes.c_double)]


@pytest.fixture(scope="module")
def host_mss():
    so = os.path.join(HERE, "host", "libmss_host.so")
    src = os.path.join(HERE, "host", "mss_host.cpp")
    core = os.path.join(ROOT, "deepgrp_b200", "csrc", "mss_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)

    def run(S, min_sc, xdrop, CH, max_rounds=0):
        S = np.ascontiguousarray(S)
        cap = S.size // 2 + 2
        out = (_Seg * cap)()
        rounds = ctypes.c_int(0)
        fn = lib.host_mss_f64 if S.dtype == np.float64 else lib.host_mss_f32
        fn.restype = ctypes.c_int
        m = fn(S.size, ctypes.c_void_p(S.ctypes.data), ctypes.c_double(min_sc), ctypes.c_double(xdrop),
               CH, out, cap, ctypes.byref(rounds), max_rounds)
        return [(out[i].st, out[i].en, out[i].sc) for i in range(m)], rounds.value

    def run_grouped(S, min_sc, xdrop, CH, group, max_rounds=0):
        S = np.ascontiguousarray(S)
        cap = S.size // 2 + 2
        out = (_Seg * cap)()
        rounds, mism = ctypes.c_int(0), ctypes.c_int(0)
        fn = lib.host_mss_grouped_f64 if S.dtype == np.float64 else lib.host_mss_grouped_f32
        fn.restype = ctypes.c_int
        m = fn(S.size, ctypes.c_void_p(S.ctypes.data), ctypes.c_double(min_sc), ctypes.c_double(xdrop),
               CH, group, out, cap, ctypes.byref(rounds), max_rounds, ctypes.byref(mism))
        return [(out[i].st, out[i].en, out[i].sc) for i in range(m)], rounds.value, mism.value
    def run_open(S, min_sc, xdrop, CH, group, L0, open_end, max_rounds=0):
        """Prefix of a record: (segments, rounds, restart index or -1, running sum before the restart run)."""
        S = np.ascontiguousarray(S)
        cap = S.size // 2 + 2
        out = (_Seg * cap)()
        rounds, restart, restart_l = ctypes.c_int(0), ctypes.c_int(-1), ctypes.c_double(0.0)
        fn = lib.host_mss_open_f64 if S.dtype == np.float64 else lib.host_mss_open_f32
        fn.restype = ctypes.c_int
        m = fn(S.size, ctypes.c_void_p(S.ctypes.data), ctypes.c_double(min_sc), ctypes.c_double(xdrop),
               CH, group, ctypes.c_double(L0), int(open_end), out, cap, ctypes.byref(rounds), max_rounds,
               ctypes.byref(restart), ctypes.byref(restart_l))
        return [(out[i].st, out[i].en, out[i].sc) for i in range(m)], rounds.value, restart.value, restart_l.value
    run.grouped = run_grouped
    run.open = run_open
    return run


def test_chunked_mss_logic_bit_exact(host_mss, oracle):
    rng = np.random.default_rng(5)
    for trial in range(120):
        n = int(rng.integers(1, 400))
        kind = trial % 6
        if kind == 0:
            S = rng.normal(size=n)
        elif kind == 1:
            S = rng.normal(size=n) - 0.5
        elif kind == 2:
            S = rng.normal(size=n) + 0.3
        elif kind == 3:
            S = rng.integers(-3, 4, size=n).astype(float)
        elif kind == 4:
            S = np.where(rng.random(n) < 0.05, 13.2, -1.32) * (1 + 1e-3 * rng.normal(size=n))
        else:
            S = rng.choice([-4.59, 4.59, -45.9, 0.0, 1e-7, -1e-7], size=n)
        if trial % 2:
            S = S.astype(np.float32)
        for xdrop in (-1.0, 0.0, 3.0, 20.0):
            for min_sc in (0.0, 2.7, 10.0):
                ref = [(int(a), int(b), float(c)) for a, b, c in
                       oracle.mss_find_all(S.astype(np.float64), min_sc, xdrop)]
                for CH in (1, 3, 16, 64, 1000):
                    for max_rounds in (0, 2):      # 2: forces the sequential completion path
                        got, _ = host_mss(S, min_sc, xdrop, CH, max_rounds)
                        assert got == ref, (trial, xdrop, min_sc, CH, max_rounds)


def _mss_in_slabs(host_mss, S, min_sc, xdrop, cuts, CH, group, max_rounds=0):
    """What dgrp_fasta_stream does with a long record (api.cu, "early rows"): after every slab of scores the
    scan runs over [restart, slab end) from the carried running sum, keeps what the last FLUSH run made final
    and resumes at that run; the last pass runs to the true end."""
    r, L0, out = 0, 0.0, []
    for p in list(cuts) + [None]:
        sub = S[r:p] if p is not None else S[r:]
        segs, _, restart, restart_l = host_mss.open(sub, min_sc, xdrop, CH, group, L0, p is not None, max_rounds)
        out += [(a + r, b + r, sc) for a, b, sc in segs]
        if p is not None and restart >= 0:
            r, L0 = r + restart, restart_l
    return out


def test_resumable_mss_equals_the_whole_record(host_mss, oracle):
    """mss.c:50-101 has no entry point that resumes; the claim checked here is that restarting at the first
    element of a run that found no candidate with a smaller L (mss.c:78-81: the stack is flushed, the run becomes
    the bottom) with the running sum carried over reproduces every later segment bit for bit -- in the exact
    regime, with x-drop resets, with upward drift (no flush: nothing is final before the end) and with sums
    that round."""
    rng = np.random.default_rng(11)
    for trial in range(200):
        n = int(rng.integers(20, 3000))
        kind = trial % 6
        base = rng.normal(size=n)
        if kind == 0:
            S = base - 0.3
        elif kind == 1:
            S = base + 0.3
        elif kind == 2:
            S = base * 5 - 1
        elif kind == 3:
            S = (base - 0.1) * 1e3 + 1e9 * (rng.random(n) < 0.002)
        elif kind == 4:
            S = np.where(rng.random(n) < 0.5, base, -np.abs(base) * 3)
        else:
            S = np.where(rng.random(n) < 0.1, 138.0, -4.6) * (1 + 1e-3 * base)
        S = S.astype(np.float32)
        if trial % 3 == 0:
            S = S.astype(np.float64) * (1 + 1e-9 * rng.normal(size=n))     # true doubles: sums round
        min_sc = float(rng.choice([0.0, 1.0, 5.0, 229.7]))
        xdrop = float(rng.choice([-1.0, 3.0, 20.0, 100.0, 2297.0]))
        ref = [(int(a), int(b), float(c)) for a, b, c in oracle.mss_find_all(S.astype(np.float64), min_sc, xdrop)]
        cuts = sorted(set(int(x) for x in rng.integers(1, n, size=int(rng.integers(1, 6)))))
        for CH, group, max_rounds in ((16, 0, 0), (64, 4, 0), (1000, 0, 0), (32, 0, 2)):
            got = _mss_in_slabs(host_mss, S, min_sc, xdrop, cuts, CH, group, max_rounds)
            assert got == ref, (trial, kind, cuts, CH, group, max_rounds)


def test_grouped_summary_chain(host_mss, oracle):
    """The parallel (grouped) summary chain: composing chunk effects (mss_core.cuh compose) must predict
    exactly the start states the sequential chain predicts whenever the additions are exact (integer
    and few-bit scores), and the final segments are the oracle's for ANY input (a wrong prediction is
    caught by the bitwise verification and only costs another round)."""
    rng = np.random.default_rng(11)
    for trial in range(160):
        n = int(rng.integers(1, 600))
        kind = trial % 5
        if kind == 0:
            S = rng.integers(-3, 4, size=n).astype(float)                 # exact arithmetic, many ties
        elif kind == 1:
            S = rng.choice([-4.5, 4.5, -45.5, 0.0, 0.25, -0.25], size=n)  # exact, reset regimes
        elif kind == 2:
            S = np.where(rng.random(n) < 0.3, 13.25, -1.25)               # exact, positive drift, no reset
        elif kind == 3:
            S = rng.normal(size=n)                                        # inexact: predictions may miss
        else:
            S = np.where(rng.random(n) < 0.05, 13.2, -1.32) * (1 + 1e-3 * rng.normal(size=n))
        if trial % 2:
            S = S.astype(np.float32)
        exact = kind <= 2
        for xdrop in (-1.0, 3.0, 40.0):
            for min_sc in (0.0, 2.7):
                ref = [(int(a), int(b), float(c)) for a, b, c in
                       oracle.mss_find_all(S.astype(np.float64), min_sc, xdrop)]
                for CH in (1, 2, 5, 16, 64):
                    for group in (1, 2, 3, 8, 32):
                        got, rounds, mism = host_mss.grouped(S, min_sc, xdrop, CH, group, 12)
                        assert got == ref, (trial, xdrop, min_sc, CH, group)
                        if exact:
                            assert mism == 0, (trial, xdrop, min_sc, CH, group, mism)


def test_chunked_mss_converges_in_few_rounds(host_mss, oracle):
    """float32-valued scores (what the fused path produces) add exactly in double, so the predicted
    chunk start states are right and the scan needs O(1) parallel rounds in every drift regime --
    including the benchmark's random-weight regime where no x-drop reset ever fires."""
    rng = np.random.default_rng(1)
    n = 300_000
    xdrop, min_sc = np.log(99) * 500, np.log(99) * 50
    cases = {
        "trained-like": np.where(np.repeat(rng.random(n // 100) < 0.4, 100), 4.59, -45.9) * rng.random(n),
        "negative drift": np.where(rng.random(n) < 0.02, 13.2, -1.32) * (1 + 1e-3 * rng.normal(size=n)),
        "positive drift": np.where(rng.random(n) < 0.3, 13.2, -1.32),
    }
    for name, S in cases.items():
        S = S.astype(np.float32)
        got, rounds = host_mss(S, min_sc, xdrop, 1024, 12)
        ref = [(int(a), int(b), float(c)) for a, b, c in oracle.mss_find_all(S.astype(np.float64), min_sc, xdrop)]
        assert got == ref, name
        assert 0 < rounds <= 4, (name, rounds)
        # the grouped (parallel) chain predicts the same start states on float32-valued scores
        got_g, rounds_g, mism = host_mss.grouped(S, min_sc, xdrop, 1024, 32, 12)
        assert got_g == ref and rounds_g == rounds and mism == 0, (name, rounds_g, mism)


def test_chunked_mss_in_the_rounding_regime(host_mss, oracle):
    """Scores that drift upwards (a confident random-weight network: 70-90 % of the positions score positive, no
    x-drop reset ever fires) push the running sum to 2^20..2^30, where the double additions round.  The shadow
    trajectory (mss_core.cuh, ChunkSummary) keeps the predicted chunk start states exact up to binade crossings:
    a few rounds, never the sequential completion -- and bit-identical segments."""
    rng = np.random.default_rng(7)
    n = 4_000_000
    xdrop, min_sc = np.log(99) * 500, np.log(99) * 50
    pos = rng.random(n) < 0.8
    S = np.where(pos, rng.random(n) * 9.0 + 0.01, -(rng.random(n) * 2.0)).astype(np.float32)
    ref = [(int(a), int(b), float(c)) for a, b, c in oracle.mss_find_all(S.astype(np.float64), min_sc, xdrop)]
    assert len(ref) == 1 and ref[0][2] > 2 ** 23          # one maximal segment, sums far beyond exactness
    for ch in (512, 1024):
        got, rounds, _ = host_mss.grouped(S, min_sc, xdrop, ch, 32, 16)
        assert got == ref and 0 < rounds <= 10, (ch, rounds)
    got, rounds = host_mss(S, min_sc, xdrop, 1024, 16)    # the sequential chain (float64 route) as well
    assert got == ref and 0 < rounds <= 10, rounds


# ---- Options / weights / CLI ---------------------------------------------------------------------
def test_options_defaults_aliases_and_toml_roundtrip():
    from deepgrp_b200.model import Options
    o = Options()
    assert (o.vecsize, o.units, o.batch_size, o.min_mss_len, o.xdrop_len, o.attention) == (150, 32, 256, 50, 50, False)
    o = Options(gru_units=60, gru_dropout=0.1, attention=True)
    assert o.units == 60 and o.dropout == 0.1 and o["gru_units"] == 60
    buf = io.StringIO()
    o.to_toml(buf)
    o2 = Options.from_toml(io.StringIO(buf.getvalue()))
    assert o2.todict() == o.todict()
    with pytest.raises(TypeError):
        Options.from_toml("not a file")


def test_reference_defaults_toml_parses():
    from deepgrp_b200.model import Options
    text = ('vecsize = 342\nunits = 60\nattention = true\nrnn = "GRU"\nbatch_size = 256\n'
            'min_mss_len = 50\nxdrop_len = 50\nrepeats_to_search = [ 1, 2, 3, 4,]\n')
    o = Options.from_toml(io.StringIO(text))
    assert (o.vecsize, o.units, o.attention) == (342, 60, True)


def test_random_weights_shapes_and_npz_roundtrip(tmp_path):
    from deepgrp_b200.model import ModelWeights, create_model, Options
    w = create_model(Options(vecsize=342, units=60, attention=True))
    assert w.kernel.shape == (5, 180) and w.recurrent_kernel.shape == (60, 180)
    assert w.bias.shape == (2, 180) and w.att_scale.shape == (60,) and w.ff_kernel.shape == (120, 5)
    assert w.input_shape == (None, 342, 5) and w.output_shape == (None, 342, 5)
    r = w.recurrent_kernel
    assert np.allclose(r @ r.T, np.eye(60), atol=1e-5)          # Keras Orthogonal
    p = str(tmp_path / "w.npz")
    w.save_npz(p)
    w2 = ModelWeights.load_npz(p)
    for k, v in w.as_dict().items():
        assert np.array_equal(v, w2.as_dict()[k])
    w3 = create_model(Options(vecsize=150, units=32, attention=False))
    assert w3.att_scale is None and w3.ff_kernel.shape == (32, 5)


def test_window_dataset_matches_reference_enumeration(oracle):
    from deepgrp_b200.prediction import fetch_validation_batch
    rng = np.random.default_rng(0)
    for L, T, step, B in ((1000, 150, 50, 16), (150, 150, 50, 4), (151, 150, 50, 4), (777, 100, 33, 5)):
        fwd = np.eye(5, dtype=np.int8)[rng.integers(0, 5, size=L)].T.copy()
        ours = list(fetch_validation_batch(fwd, step, B, T))
        ref = list(oracle.fetch_validation_batch(fwd, step, B, T))
        assert len(ours) == len(ref)
        for a, b in zip(ours, ref):
            assert a.dtype == np.float32 and np.array_equal(a, b)


def test_cli_argument_mapping(monkeypatch):
    from deepgrp_b200 import __main__ as cli
    seen = {}
    monkeypatch.setattr(cli.CommandLineParser, "predict",
                        staticmethod(lambda args, options: seen.update(args=args, options=options)))
    p = cli.CommandLineParser().parse_args(["-b", "64", "-s", "25", "-x", "7", "-l", "9", "-vv",
                                            "predict", "m.hdf5", "a.fa", "b.fa", "--no_use_mss",
                                            "--output", "o.tsv"])
    p.set_logging().setup_tensorflow().run()
    a, o = seen["args"], seen["options"]
    assert (a.model, a.FASTA, a.output, a.no_use_mss, a.step_size) == ("m.hdf5", ["a.fa", "b.fa"], "o.tsv", True, 25)
    assert (o.batch_size, o.min_mss_len, o.xdrop_len) == (64, 9, 7)
    # README form without the sub-command (reference README.rst:92-98)
    p = cli.CommandLineParser().parse_args(["-s", "10", "model.npz", "x.fa"])
    p.run()
    assert seen["args"].command == "predict" and seen["args"].model == "model.npz" and seen["args"].FASTA == ["x.fa"]


def test_cli_fasta_reader_matches_oracle(oracle):
    from deepgrp_b200.__main__ import _read_multi_fasta
    text = "ACGT\n>h1 desc\nacgt\n NNAC \n>\nGGGG\n>h3\nTT"
    assert list(_read_multi_fasta(io.StringIO(text))) == list(oracle.read_multi_fasta(io.StringIO(text)))
    with pytest.raises(IndexError):
        list(_read_multi_fasta(io.StringIO(">a\nAC\n\nGT\n")))


# ---- GPU FASTA decode logic on the host ---------------------------------------------------------------
@pytest.fixture(scope="module")
def host_fasta():
    so = os.path.join(HERE, "host", "libfasta_host.so")
    src = os.path.join(HERE, "host", "fasta_host.cpp")
    core = os.path.join(ROOT, "deepgrp_b200", "csrc", "fasta_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.fasta_host_decode.restype = ctypes.c_int

    def decode(raw: bytes):
        """-> list of (header, sequence) as api.cu assembles them from the decode's tables, or IndexError."""
        n = len(raw)
        buf = np.frombuffer(raw, dtype=np.uint8) if n else np.zeros(1, np.uint8)
        seq = np.zeros(max(n, 1), np.uint8)
        hpos, hseq = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.int64)
        n_seq, n_hdr = ctypes.c_int64(0), ctypes.c_int64(0)
        rc = lib.fasta_host_decode(ctypes.c_void_p(buf.ctypes.data), ctypes.c_int64(n), ctypes.c_void_p(seq.ctypes.data),
                                   ctypes.byref(n_seq), ctypes.c_void_p(hpos.ctypes.data),
                                   ctypes.c_void_p(hseq.ctypes.data), ctypes.byref(n_hdr))
        if rc:
            raise IndexError("string index out of range")
        bounds = hseq[:n_hdr.value].tolist() + [n_seq.value]
        records = []
        for k in range(n_hdr.value):
            e = b = int(hpos[k]) + 1                        # header = rest of the line, right-stripped (api.cu)
            while e < n and raw[e] != 10 and not (raw[e] == 13 and not (e + 1 < n and raw[e + 1] == 10)):
                e += 1
            header = raw[b:e].decode("latin-1").rstrip("\t\n\x0b\x0c\r\x1c\x1d\x1e\x1f ")
            if header:
                records.append((header, seq[bounds[k]:bounds[k + 1]].tobytes().decode("latin-1").upper()))
        return records

    decode.lib = lib
    return decode


def _python_reader(text):
    import io
    from deepgrp_b200.__main__ import _read_multi_fasta
    return list(_read_multi_fasta(io.StringIO(text, newline=None)))


def test_fasta_push_byte_equals_table_form(host_fasta):
    assert host_fasta.lib.fasta_host_push_byte_mismatches() == 0


@pytest.mark.parametrize("seed", range(40))
def test_fasta_kernel_logic_on_hostile_text(host_fasta, seed):
    """The kernel logic of csrc/fasta.cu (compiled for the host) against `_read_multi_fasta` (reference
    deepgrp/__main__.py:20-43): mixed line ends, whitespace around and inside lines, ragged widths, dropped
    records, blank lines."""
    from conftest import messy_fasta
    text = messy_fasta(seed)
    try:
        exp = _python_reader(text)
    except IndexError:
        with pytest.raises(IndexError):
            host_fasta(text.encode("latin-1"))
        return
    assert host_fasta(text.encode("latin-1")) == exp


@pytest.mark.parametrize("text", [
    "", ">a\nACGT\n", ">a\nACGT", ">a\r\nAC\r\nGT\r\n", ">a\rAC\rGT\r", "ACGT\n>a\nAC\n", ">\nACGT\n>b\nGG\n",
    ">a\n\nAC\n", ">a\nAC\n\n", ">a\nAC\n \n", ">a\nAC\n ", ">a\n A C \n", " \t>a b \t\nAC\n", ">a\nAC\r", "\n", " ",
    ">a\n" + "ACGT" * 5000 + "\n>b\n" + "N" * 4095 + "\n" + "T" * 4097 + "\n",     # lines across tile borders
    ">a\n" + ("ACGTACGTACGTACG\r\n" * 600),                                          # CR LF split between threads
])
def test_fasta_kernel_logic_edge_cases(host_fasta, text):
    try:
        exp = _python_reader(text)
    except IndexError:
        with pytest.raises(IndexError):
            host_fasta(text.encode("latin-1"))
        return
    assert host_fasta(text.encode("latin-1")) == exp


# ---- segment kernels and TSV number formatting on the host -------------------------------------------
@pytest.fixture(scope="module")
def host_seg():
    so = os.path.join(HERE, "host", "libseg_host.so")
    src = os.path.join(HERE, "host", "seg_host.cpp")
    core = os.path.join(ROOT, "deepgrp_b200", "csrc", "seg_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.seg_host_rows.restype = ctypes.c_int64
    lib.seg_host_fmt.restype = ctypes.c_int

    def rows(lab, mis=0, open_end=False, offset=0):
        lab = np.ascontiguousarray(lab, dtype=np.uint8)
        tri = np.zeros((lab.size + 2, 3), np.int64)
        m = lib.seg_host_rows(ctypes.c_void_p(lab.ctypes.data), ctypes.c_int64(lab.size), int(mis), int(open_end),
                              ctypes.c_int64(offset), ctypes.c_void_p(tri.ctypes.data), ctypes.c_int64(lab.size + 2))
        assert m >= 0, "incomplete row or unpaired start / end"
        return tri[:m]
    rows.lib = lib
    return rows


def test_segment_kernel_logic_equals_yield_segments(host_seg, oracle):
    """seg_core.cuh (four labels per 32-bit operation, tile lists, rows that straddle tiles) against
    sequence.pyx:40-53, 79-85 restated by the oracle: every alignment of the label pointer, labels with the high
    bit set, runs longer than a tile, the `size - 1` special case, and the open-ended form of the early rows."""
    rng = np.random.default_rng(0)
    for trial in range(160):
        n = int(rng.integers(1, 20000))
        kind = trial % 4
        if kind == 0:
            lab = rng.choice([0, 1, 2, 3, 4, 127, 128, 255], n)
        elif kind == 1:
            lab = np.repeat(rng.integers(0, 3, n // 50 + 1), 50)[:n]
        elif kind == 2:
            lab = np.repeat(rng.integers(0, 4, n // 5000 + 1), 5000)[:n]
        else:
            lab = np.where(rng.random(n) < 0.02, 0, 1)
        lab = lab.astype(np.int64)
        mis, off = int(rng.integers(0, 16)), int(rng.integers(0, 1000))
        exp = np.array([(s, e, l) for s, e, l in oracle.yield_segments(lab, off) if l > 0], np.int64).reshape(-1, 3)
        got = host_seg(lab, mis, False, off)
        assert got.shape == exp.shape and (got == exp).all(), (trial, n, mis)
        # open end: the rows of a longer record, except that a run reaching the end of the prefix is closed there
        longer = np.concatenate([lab, [0, 0, 0]])
        exp_o = np.array([(s, e, l) for s, e, l in oracle.yield_segments(longer, off) if l > 0], np.int64).reshape(-1, 3)
        got_o = host_seg(lab, mis, True, off)
        assert got_o.shape == exp_o.shape, (trial, "open")
        if lab[-1] != 0:
            assert (got_o[:-1] == exp_o[:-1]).all() and got_o[-1, 0] == exp_o[-1, 0] and got_o[-1, 1] == n + off
        else:
            assert (got_o == exp_o).all()
    for lab in ([7], [0], [3, 3], [3, 0], [0, 3], [5] * 4096, [5] * 4097, [1, 2] * 2049):
        lab = np.array(lab, np.int64)
        exp = np.array([(s, e, l) for s, e, l in oracle.yield_segments(lab, 0) if l > 0], np.int64).reshape(-1, 3)
        for mis in (0, 1, 15):
            got = host_seg(lab, mis)
            assert got.shape == exp.shape and (got == exp).all(), (lab[:4], lab.size, mis)


def test_tsv_number_text_equals_python(host_seg):
    """tsv.cu formats a row's numbers with 32-bit arithmetic below 2^32 and a 64-bit loop above; both must be
    Python's "{}".format (deepgrp/__main__.py:288-292)."""
    buf = (ctypes.c_ubyte * 32)()
    rng = np.random.default_rng(1)
    values = [0, 1, 9, 10, 11, 99, 100, 999_999_999, 1_000_000_000, 2_147_483_647, 4_294_967_295, 4_294_967_296,
              9_999_999_999, 10**18, 2**63 - 1, -1, -10, -4_294_967_296, -(2**63)]
    values += [10**k + d for k in range(1, 19) for d in (-1, 0, 1)]
    values += [int(x) for x in rng.integers(0, 2**31, 300)] + [int(x) for x in rng.integers(-2**62, 2**62, 100)]
    for v in values:
        n = host_seg.lib.seg_host_fmt(ctypes.c_longlong(v), buf)
        assert n > 0 and bytes(buf[:n]) == str(v).encode(), v


def _gap_fill(lab2, labels, base, segs, nof):
    """pymss.pyx:59-77 on the segments of one pass (local coordinates + base)."""
    for st, en, _ in segs:
        part = labels[base + st:base + en]
        counts = np.bincount(part, minlength=nof)
        best, bestv = 1, counts[1]
        for k in range(2, nof):
            if bestv < counts[k]:
                best, bestv = k, counts[k]
        lab2[base + st:base + en] = np.where(part == 0, best, part)


def test_early_rows_scheme_on_the_host(host_mss, host_seg, oracle):
    """The early rows of dgrp_fasta_stream (api.cu) with the kernels' own scalar logic on the CPU: after every slab
    the resumable MSS (mss_core.cuh) over [restart, slab end), gap fill of what became final, the open-ended segment
    pass (seg_core.cuh) over [emitted, restart) with the run that reaches the end held back; the last slab closes
    both.  The rows must be those of the reference's whole-record sequence find_mss_labels -> argmax ->
    yield_segments (prediction.py:58, __main__.py:83, 288-292), whatever the cuts."""
    rng = np.random.default_rng(21)
    s0 = float(np.log(0.99 / (1.0 - 0.99)))
    nof = 5
    early_rows = 0
    for trial in range(60):
        n = int(rng.integers(300, 6000))
        kind = trial % 4
        # labels with runs, scores as apply_mss makes them: positive for repeats, -10 x for class 0 (prediction.py:51-57)
        labels = np.repeat(rng.choice([0, 0, 0, 1, 2, 3, 4], n // 7 + 1), rng.integers(1, 15, n // 7 + 1))[:n]
        if labels.size < n:
            labels = np.concatenate([labels, np.zeros(n - labels.size, np.int64)])
        labels = labels.astype(np.int64)
        conf = {0: 0.9, 1: 0.6, 2: 0.3, 3: 0.52}[kind]
        m = np.clip(conf + 0.25 * rng.normal(size=n), 0.05, 0.99).astype(np.float32)
        t = np.log(m / (1 - m)).astype(np.float32)
        scores = np.where(labels > 0, t, -10 * t).astype(np.float32)
        min_len, xdrop_len = int(rng.choice([1, 5, 50])), int(rng.choice([0, 5, 50]))
        min_sc, xdrop = s0 * min_len, (s0 * xdrop_len * 10.0 if xdrop_len > 0 else -1.0)
        start = int(rng.integers(0, 50))
        final = oracle.find_mss_relabel(scores.astype(np.float64), labels, nof, min_len, xdrop_len).astype(np.int64)
        exp = [(s, e, l) for s, e, l in oracle.yield_segments(final, start) if l > 0]
        cuts = sorted(set(int(x) for x in rng.integers(1, n, size=int(rng.integers(1, 6)))))
        r = e = 0
        L0, rows, lab2 = 0.0, [], labels.copy()
        for p in cuts:
            segs, _, restart, restart_l = host_mss.open(scores[r:p], min_sc, xdrop, 64, 4, L0, True)
            if restart > 0:
                _gap_fill(lab2, labels, r, [sg for sg in segs], nof)
                r, L0 = r + restart, restart_l
                tri = host_seg(lab2[e:r].astype(np.uint8), int(rng.integers(0, 16)), True, start + e)
                tail = r - e
                if len(tri) and lab2[r - 1] != 0:          # the run that reaches the end of the prefix is held back
                    tail = int(tri[-1, 0]) - (start + e)
                    tri = tri[:-1]
                rows += [tuple(int(x) for x in row) for row in tri]
                e += tail
        early_rows += len(rows)
        segs, _, _, _ = host_mss.open(scores[r:], min_sc, xdrop, 64, 4, L0, False)
        _gap_fill(lab2, labels, r, segs, nof)
        rows += [tuple(int(x) for x in row) for row in host_seg(lab2[e:].astype(np.uint8), 3, False, start + e)]
        assert np.array_equal(lab2, final), (trial, kind)
        assert rows == exp, (trial, kind, cuts)
    assert early_rows > 1000          # the scheme was exercised: rows did leave before the last slab
