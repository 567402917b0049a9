#!/bin/bash
# round 2, state X: MUFU.TANH in the attention scores of the half-precision-sum mode
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/r02x_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r02x_tests.log | head -20
for rep in 1 2; do
timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sections x4 > gpurun_out/r02x_bench_$rep.json 2> gpurun_out/r02x_bench_$rep.err; echo "bench rep=$rep rc=$?"
done
timeout -s KILL 200 python bench.py --bases 24800000 --vecsize 512 --units 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02x_cfg5b.json 2> gpurun_out/r02x_cfg5b.err; echo "5b rc=$?"
timeout -s KILL 300 python tools/accuracy_study.py > gpurun_out/r02x_accuracy.txt 2>&1; echo "acc rc=$?"; tail -12 gpurun_out/r02x_accuracy.txt
python - <<'PY'
import json
for f in ("bench_1", "bench_2", "cfg5b"):
    try:
        d = json.load(open("gpurun_out/r02x_%s.json" % f))
        print(f, "total ms", round(d["stages_ms"]["total_ms"], 2), "fwd", round(d["stages_ms"]["forward_ms"], 2), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "TF", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 4), "MHz", d["clocks"]["sm_mhz"], ("x4 %.1f" % d["x4"]["value"]) if "x4" in d else "")
    except Exception as e:
        print(f, "failed", e)
PY
