#!/bin/bash
# round 2, state R: direction-per-slot form of the two-tile kernel against the interleaved form
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02r_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02r_tests.log
for rep in 1 2; do for v in 1 0; do
DEEPGRP_KNOBS="forward_dirslots=$v" timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sections "" > gpurun_out/r02r_ds${v}_$rep.json 2> gpurun_out/r02r_ds${v}_$rep.err; echo "dirslots=$v rep=$rep rc=$?"
done; done
python - <<'PY'
import json
for rep in (1, 2):
    for v in (1, 0):
        try:
            d = json.load(open("gpurun_out/r02r_ds%d_%d.json" % (v, rep)))
            print("dirslots", v, "rep", rep, "fwd ms", round(d["stages_ms"]["forward_ms"], 2), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "TF", round(d["roofline"]["achieved"], 1), "sm MHz", d["clocks"]["sm_mhz"])
        except Exception as e:
            print(v, rep, "failed", e)
PY
