#!/bin/bash
# round 2, state S: what the driver runs at round end on one GPU -- GPU tests, smoke(), both bench arms
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02s_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02s_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02s_smoke.log
timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02s_bench_ref.json 2> gpurun_out/r02s_bench_ref.err; echo "ref rc=$?"
timeout -s KILL 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02s_bench.json"))
r = json.load(open("gpurun_out/r02s_bench_ref.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ref", round(r["value"],4), "e2e ratio", round(d["e2e"]["value"]/r["value"]), "launches", d["gpu_launches"], "frac", round(d["roofline"]["frac"],4))
print("x4", d["x4"]["value"], d["x4"]["e2e"], "strong", d["strong"]["value"], "genome", d["genome"]["value"], d["genome"]["seconds"])
PY
