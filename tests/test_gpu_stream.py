"""Early rows of dgrp_fasta_stream_* (api.cu): a long record is computed in position slabs and the rows that are
final after a slab -- everything before the last run at which the reference flushes its candidate stack
(deepgrp/_mss/mss.c:78-81) -- are formatted and copied back while the next slab's forward runs.  The text must be
byte for byte what the one-shot call returns (which runs MSS once over the whole record), in every score regime:
downward drift (flushes everywhere), upward drift (no flush: nothing leaves early), long runs of one label, N runs,
both placement modes.  The resumed scan itself is pinned against the reference algorithm without a GPU in
tests/test_host.py::test_resumable_mss_equals_the_whole_record."""
import io

import numpy as np
import pytest

from conftest import random_dna, write_fasta

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dg(gpu_ctx):
    import deepgrp_b200.model as model
    import deepgrp_b200.prediction as pred

    class NS:
        pass
    ns = NS()
    ns.model, ns.pred, ns.ctx = model, pred, gpu_ctx
    return ns


class early:
    """Context options for the early-rows route: rows = 2 every long record (the default, 1, only the stream's last
    record), 0 off; `unit` = windows per slab unit (the default is one wave of the forward kernel, far above a test
    record)."""

    def __init__(self, ctx, rows=2, unit=64, slabs=0):
        self.ctx, self.opts = ctx, {"stream_early_rows": rows, "stream_early_unit": unit, "stream_early_slabs": slabs}

    def __enter__(self):
        for k, v in self.opts.items():
            self.ctx.set_int(k, v)
        return self

    def __exit__(self, *exc):
        self.ctx.set_int("stream_early_rows", 1)
        self.ctx.set_int("stream_early_unit", 0)
        self.ctx.set_int("stream_early_slabs", 0)


def stream_text(dg, w, raw, name, compat="reference", step=50):
    out = io.BytesIO()
    stats = dg.pred.predict_fasta_tsv_stream(w, raw, name, out, step, 256, True, 50, 50, compat)
    return out.getvalue(), stats


def one_record(n, seed, n_runs=False):
    seq = random_dna(n, seed)
    if n_runs:
        rng = np.random.default_rng(seed)
        s = list(seq)
        for _ in range(6):
            a, k = int(rng.integers(0, n - 500)), int(rng.integers(5, 400))
            s[a:a + k] = ["N"] * k
        seq = "NNNN" + "".join(s) + "NN"
    return (">chrT test record %d\n" % n + "\n".join(seq[i:i + 70] for i in range(0, len(seq), 70)) + "\n").encode()


@pytest.mark.parametrize("T,U,scale", [(150, 32, 1.0), (150, 32, 4.0), (342, 60, 1.0), (100, 20, 2.0)])
def test_early_rows_equal_the_one_shot_text(dg, T, U, scale):
    w = dg.model.random_weights(T, U, attention=True, seed=0)
    if scale != 1.0:
        w = w.scaled(scale)
    raw = one_record(60_000, 7)
    ref = bytes(dg.pred.predict_fasta_tsv_view(w, raw, "e.fa", 50, 256, True, 50, 50))
    assert ref.count(b"\n") > 10
    parts_seen = []
    for slabs in (0, 2, 3, 6):
        with early(dg.ctx, 2, 64, slabs):
            got, stats = stream_text(dg, w, raw, "e.fa")
            parts_seen.append(dg.ctx.get_int("stream_early_parts"))
        assert got == ref, (T, U, scale, slabs)
        assert stats["rows"] == ref.count(b"\n") and stats["records"] == 1
    with early(dg.ctx, 0, 64, 0):
        got, _ = stream_text(dg, w, raw, "e.fa")
        assert got == ref and dg.ctx.get_int("stream_early_parts") == 0
    if (T, U, scale) == (342, 60, 1.0):
        # the benchmark's regime (defaults.toml shape, random-init weights): the scores drift downwards, the stack
        # is flushed every few hundred positions and every slab leaves rows behind
        assert min(parts_seen) >= 1 and max(parts_seen) >= 3, parts_seen


def test_early_rows_pieces_and_record_marks(dg):
    """Parts of a record travel as pieces without the end-of-record mark; the last piece of the last part carries
    it, and the per-piece row counts add up."""
    w = dg.model.random_weights(342, 60, attention=True, seed=0)
    raw = one_record(50_000, 11) + one_record(4_000, 12).replace(b"chrT", b"chrS") + one_record(45_000, 13).replace(b"chrT", b"chrU")
    ref = bytes(dg.pred.predict_fasta_tsv_view(w, raw, "p.fa", 50, 256, True, 50, 50))
    with early(dg.ctx, 2, 64, 4):
        text, marks, rows = {}, {}, 0
        with dg.pred.FastaTsvStream(w, raw, "p.fa", 50, 256, True, 50, 50) as st:
            for sl, od, nr, last, view in st:
                text.setdefault((sl, od), []).append(bytes(view))
                marks.setdefault((sl, od), []).append(last)
                rows += nr
    assert sorted(text) == [(0, 0), (0, 1), (0, 2)]
    assert b"".join(b"".join(text[k]) for k in sorted(text)) == ref
    assert rows == ref.count(b"\n")
    for k, m in marks.items():
        assert m[-1] and not any(m[:-1]), (k, m)
    assert len(marks[(0, 0)]) >= 3          # the long record left in several parts
    # the default (1): only the stream's last record is computed in slabs -- the text of the others crosses PCIe
    # under the next record's forward anyway
    with early(dg.ctx, 1, 64, 4):
        marks1, text1 = {}, {}
        with dg.pred.FastaTsvStream(w, raw, "p.fa", 50, 256, True, 50, 50) as st:
            for sl, od, nr, last, view in st:
                marks1.setdefault((sl, od), []).append(last)
                text1.setdefault((sl, od), []).append(bytes(view))
    assert b"".join(b"".join(text1[k]) for k in sorted(text1)) == ref
    assert len(marks1[(0, 0)]) == 1 and len(marks1[(0, 1)]) == 1 and len(marks1[(0, 2)]) >= 3


@pytest.mark.parametrize("compat", ("reference", "fixed"))
def test_early_rows_with_n_runs_steps_and_placement(dg, compat):
    """N runs inside the record (codes 4), edge N trim (startpos > 0), a step that does not divide the window, both
    placements of the short last batch."""
    w = dg.model.random_weights(150, 32, attention=True, seed=3)
    for seed, step, n in ((21, 50, 41_234), (22, 37, 30_011), (23, 150, 52_000)):
        raw = one_record(n, seed, n_runs=True)
        ref = bytes(dg.pred.predict_fasta_tsv_view(w, raw, "n.fa", step, 256, True, 50, 50, compat))
        with early(dg.ctx, 2, 48, 5):
            got, _ = stream_text(dg, w, raw, "n.fa", compat, step)
        assert got == ref, (seed, step, compat)


def test_early_rows_long_label_runs(dg):
    """Confident outputs: rows are long runs of one label that cross the slab cuts and the resume points, so the
    open-ended segment pass has to hold the last run back every time."""
    w = dg.model.random_weights(150, 32, attention=True, seed=5).scaled(8.0)
    raw = one_record(80_000, 31)
    ref = bytes(dg.pred.predict_fasta_tsv_view(w, raw, "l.fa", 50, 256, True, 50, 50))
    for unit in (32, 64, 200):
        with early(dg.ctx, 2, unit, 5):
            got, _ = stream_text(dg, w, raw, "l.fa")
        assert got == ref, unit


def test_early_rows_default_unit_on_a_long_record(dg):
    """The default slab plan (one forward wave per unit) on a record long enough to use it: 12 Mbp of the
    benchmark's shape (12.7 units -> 4 slabs)."""
    w = dg.model.random_weights(342, 60, attention=True, seed=0)
    raw = (">big\n" + random_dna(12_000_000, 77) + "\n").encode()
    ref = bytes(dg.pred.predict_fasta_tsv_view(w, raw, "b.fa", 50, 256, True, 50, 50))
    got, stats = stream_text(dg, w, raw, "b.fa")
    assert dg.ctx.get_int("stream_early_parts") >= 2
    assert got == ref and stats["rows"] == ref.count(b"\n")
