#!/usr/bin/env python
"""Per-kernel table of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches, mean
duration, share of the serialised total and -- for the HBM-bound kernels of the path -- algorithmic
bytes per launch / duration against the measured HBM peak of MEASURED_PEAKS.json.

    python tools/launch_table.py profiles/r01j_launches.csv [--bases 46700000 --text 47478362 \
        --vecsize 342 --step 50 --rows 19352370 --tsv 1151928519]

The byte counts are the ALGORITHMIC ones (what the kernel must read and write once), stated per kernel
below; ncu's per-launch times are cold-cache and serialised, so the GB/s are lower bounds of what the
kernels reach inside a step."""
import argparse
import collections
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path, until=None):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].replace("void ", "").split("(")[0].split("<")[0].replace("dgrp::", "")
        if until and name == until:      # e.g. the first FASTA decode: only the device-resident steps before it
            break
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")) * scale[r[ui]])
    return agg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--bases", type=int, default=46_700_000)
    ap.add_argument("--text", type=int, default=47_478_362, help="FASTA bytes per step (e2e leg)")
    ap.add_argument("--vecsize", type=int, default=342)
    ap.add_argument("--step", type=int, default=50)
    ap.add_argument("--classes", type=int, default=5)
    ap.add_argument("--rows", type=int, default=19_352_370, help="TSV rows per step")
    ap.add_argument("--tsv", type=int, default=1_151_928_519, help="TSV bytes per step")
    ap.add_argument("--until", default=None, help="stop at the first launch of this kernel (fa_tile_fn_kernel: only "
                                                  "the device-resident steps, whose launches cover the whole record)")
    a = ap.parse_args()
    L, T, C = a.bases, a.vecsize, a.classes
    W = len(range(0, L - T, a.step))
    # algorithmic bytes per launch: (bytes, how they are counted)
    algo = {
        "fa_tile_fn_kernel": (a.text, "text read"),
        "fa_count_kernel": (a.text, "text read"),
        "fa_scatter_kernel": (a.text + L, "text read + sequence bytes written"),
        "codes_kernel": (2 * L, "sequence byte read + code written"),
        "trim_kernel": (L, "sequence bytes read"),
        "vote_gather_kernel": (4 * C * (W * T + 2 * L), "f32 window probabilities [W,T,C] read + predictions [L,C] read and written"),
        "score_kernel": (4 * C * L + 5 * L, "predictions read + label (1 B) and score (4 B) written"),
        "mss_scan_kernel": (5 * L, "score (4 B) + label (1 B) read, one round"),
        "mss_count_runs_kernel": (4 * L, "score read"),
        "copy_labels_kernel": (2 * L, "label read + written"),
        "seg_count_kernel": (L, "labels read"),
        "seg_scatter_kernel": (L + 20 * a.rows, "labels read + (start, end, label) per run written"),
        "mss_count_runs4_kernel": (4 * L, "score read"),
        "copy_labels16_kernel": (2 * L, "label read + written"),
        "segv_count_kernel": (L, "labels read"),
        "segv_scatter_kernel": (L + 24 * a.rows, "labels read + (start, end, label) int64 triples written"),
        "tsv_len_kernel": (20 * a.rows + 4 * a.rows, "triples read + row length written"),
        "tsv_write_kernel": (20 * a.rows + 8 * a.rows + a.tsv, "triples + row offsets read, text written"),
    }
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        hbm = 6650.0
    agg = load(a.csv, a.until)
    total = sum(sum(v) for v in agg.values())
    print("| kernel | launches | mean us | share | algorithmic MB / launch | GB/s | of %.0f GB/s | bytes counted |" % hbm)
    print("|---|---|---|---|---|---|---|---|")
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        mean = sum(v) / len(v)
        if name in algo:
            b, how = algo[name]
            gbs = b / mean / 1e3
            print("| `%s` | %d | %.1f | %.2f %% | %.1f | %.0f | %.1f %% | %s |"
                  % (name, len(v), mean, 100 * sum(v) / total, b / 1e6, gbs, 100 * gbs / hbm, how))
        else:
            print("| `%s` | %d | %.1f | %.2f %% | | | | |" % (name, len(v), mean, 100 * sum(v) / total))


if __name__ == "__main__":
    main()
