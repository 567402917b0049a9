#!/bin/bash
# round 2, state P: fused vote + score
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02p_tests.log
for fz in 1 0; do
DEEPGRP_KNOBS="forward_fuse_score=$fz" timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sections "" > gpurun_out/r02p_bench_fuse$fz.json 2> gpurun_out/r02p_bench_fuse$fz.err; echo "fuse=$fz rc=$?"
done
python - <<'PY'
import json
for fz in (1, 0):
    d = json.load(open("gpurun_out/r02p_bench_fuse%d.json" % fz))
    print("fuse", fz, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["stages_ms"], d["gpu_launches"])
PY
