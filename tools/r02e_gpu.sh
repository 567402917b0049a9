#!/bin/bash
# round 2, state E: MSS shadow trajectories (rounding regime), uploader thread; bench with sections; 5b at 248 Mbp
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02e_tests.log
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02e_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02e_bench.json"))
    print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "stages", d["stages_ms"], "rounds", d["mss_rounds"])
    for k in ("x4", "strong", "genome"):
        v = d.get(k) or {}
        print(k, {kk: v[kk] for kk in v if kk in ("value", "e2e", "ms_per_step", "e2e_ms_per_step", "stages_ms", "seconds", "rank_seconds", "rank_forward_seconds", "rows_per_step", "mss_rounds")})
except Exception as e:
    print("bench failed", e)
PY
timeout -s KILL 200 python bench.py --bases 248000000 --vecsize 512 --units 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02e_cfg5b_248.json 2> gpurun_out/r02e_cfg5b_248.err; echo "5b rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02e_cfg5b_248.json"))
    print("5b value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "stages", d["stages_ms"], "rounds", d["mss_rounds"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"])
except Exception as e:
    print("5b failed", e)
PY
