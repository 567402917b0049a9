"""Row-level diff of two `deepgrp predict` TSV outputs (reference deepgrp/__main__.py:288-292:
"{file}\t{header}\t{start}\t{end}\t{label}\n", label > 0, 0-based half-open).

    python tools/tsv_diff.py a.tsv b.tsv [--json out.json]

Reports, per record and in total: rows in both, rows only in A / only in B, and the number of BASES whose
label differs (rows painted back onto the positions) -- the figure `north_star` calls "BED output diffed".
As a module: diff_tsv(text_a, text_b) -> dict."""
import argparse
import json
import sys
from collections import defaultdict

import numpy as np


def parse(text):
    recs = defaultdict(list)
    for line in text.splitlines():
        if not line:
            continue
        prefix, start, end, label = line.rsplit("\t", 3)      # a header may itself contain tabs
        f, _, header = prefix.partition("\t")
        recs[(f, header)].append((int(start), int(end), int(label)))
    return recs


def paint(rows, lo, hi):
    lab = np.zeros(hi - lo, np.uint8)
    for s, e, l in rows:
        lab[s - lo:e - lo] = l
    return lab


def diff_tsv(text_a, text_b):
    a, b = parse(text_a), parse(text_b)
    total = {"rows_a": 0, "rows_b": 0, "rows_both": 0, "rows_only_a": 0, "rows_only_b": 0, "bases_differ": 0,
             "bases_labelled_a": 0, "bases_labelled_b": 0}
    records = []
    for key in sorted(set(a) | set(b)):
        ra, rb = a.get(key, []), b.get(key, [])
        sa, sb = set(ra), set(rb)
        both = len(sa & sb)
        span = [r for r in ra + rb]
        lo, hi = min(r[0] for r in span), max(r[1] for r in span)
        la, lb = paint(ra, lo, hi), paint(rb, lo, hi)
        d = {"file": key[0], "header": key[1], "rows_a": len(ra), "rows_b": len(rb), "rows_both": both,
             "rows_only_a": len(sa) - both, "rows_only_b": len(sb) - both,
             "bases_differ": int((la != lb).sum()), "bases_labelled_a": int((la > 0).sum()),
             "bases_labelled_b": int((lb > 0).sum())}
        records.append(d)
        for k in total:
            total[k] += d[k]
    return {"total": total, "records": records}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("a")
    ap.add_argument("b")
    ap.add_argument("--json")
    args = ap.parse_args()
    out = diff_tsv(open(args.a).read(), open(args.b).read())
    text = json.dumps(out, indent=1)
    if args.json:
        open(args.json, "w").write(text)
    print(json.dumps(out["total"]))
    return 0 if out["total"]["bases_differ"] == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
