#!/bin/bash
# round 2, state K: two-tile kernel variant 1 (per-warp arrivals, wait hint) against variant 0, alternating on one box
set -u
mkdir -p gpurun_out
for rep in 1 2; do for v in 0 1; do
DEEPGRP_KNOBS="forward_variant=$v" timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sections "" > gpurun_out/r02k_v${v}_$rep.json 2> gpurun_out/r02k_v${v}_$rep.err; echo "v=$v rep=$rep rc=$?"
done; done
python - <<'PY'
import json
for rep in (1, 2):
    for v in (0, 1):
        try:
            d = json.load(open("gpurun_out/r02k_v%d_%d.json" % (v, rep)))
            print("variant", v, "rep", rep, "fwd ms", round(d["stages_ms"]["forward_ms"], 2), "value", round(d["value"], 1), "sm MHz", d["clocks"]["sm_mhz"])
        except Exception as e:
            print(v, rep, "failed", e)
PY
DEEPGRP_KNOBS="forward_variant=1" timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02k_tests_v1.log 2>&1; echo "tests v1 rc=$?"; tail -2 gpurun_out/r02k_tests_v1.log
