#!/bin/bash
# round 2, state T: shared-memory vote (no window probabilities in HBM)
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02t_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02t_tests.log
for rep in 1 2; do for v in 1 0; do
DEEPGRP_KNOBS="forward_smem_vote=$v" timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sections "" > gpurun_out/r02t_sv${v}_$rep.json 2> gpurun_out/r02t_sv${v}_$rep.err; echo "smem_vote=$v rep=$rep rc=$?"
done; done
DEEPGRP_KNOBS="forward_smem_vote=1" timeout -s KILL 200 python bench.py --bases 24800000 --vecsize 512 --units 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02t_cfg5b_sv1.json 2> gpurun_out/r02t_cfg5b_sv1.err; echo "5b sv1 rc=$?"
DEEPGRP_KNOBS="forward_smem_vote=0" timeout -s KILL 200 python bench.py --bases 24800000 --vecsize 512 --units 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02t_cfg5b_sv0.json 2> gpurun_out/r02t_cfg5b_sv0.err; echo "5b sv0 rc=$?"
python - <<'PY'
import json
for f in ("sv1_1", "sv0_1", "sv1_2", "sv0_2", "cfg5b_sv1", "cfg5b_sv0"):
    try:
        d = json.load(open("gpurun_out/r02t_%s.json" % f))
        print(f, "total ms", round(d["stages_ms"]["total_ms"], 2), "fwd", round(d["stages_ms"]["forward_ms"], 2), "score", round(d["stages_ms"]["score_ms"], 2), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "MHz", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "failed", e)
PY
