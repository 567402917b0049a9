#!/bin/bash
# round 2, state H: 2..4 column blocks, default-scope remote arrives, mapped-memory count readbacks
set -u
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "wide or lstm or config5 or stream" > gpurun_out/r02h_tests_wide.log 2>&1; echo "wide tests rc=$?"; tail -4 gpurun_out/r02h_tests_wide.log
for knobs in "forward_ub=32" "forward_ub=64" "forward_ub=64,forward_overlap=0"; do
tag=$(echo $knobs | tr ',=' '__')
DEEPGRP_KNOBS="$knobs" timeout -s KILL 200 python bench.py --bases 24800000 --vecsize 512 --units 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02h_cfg5b_$tag.json 2> gpurun_out/r02h_cfg5b_$tag.err; echo "5b $knobs rc=$?"
done
timeout -s KILL 200 python bench.py --bases 46700000 --vecsize 342 --units 60 --rnn LSTM --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02h_lstm.json 2> gpurun_out/r02h_lstm.err; echo "lstm rc=$?"
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02h_tests.log
timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sections genome > gpurun_out/r02h_bench_genome.json 2> gpurun_out/r02h_bench_genome.err; echo "genome rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02h_cfg5b_*.json")) + ["gpurun_out/r02h_lstm.json"]:
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"],1), "fwd ms", round(d["stages_ms"]["forward_ms"],1), "TF", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "failed", e)
try:
    d = json.load(open("gpurun_out/r02h_bench_genome.json"))
    g = d["genome"]
    print({k: g[k] for k in ("value", "seconds", "rank_forward_seconds", "rank0_gpu_seconds", "rank0_waits_ms")})
    print("main value", d["value"], "e2e", d["e2e"])
except Exception as e:
    print("genome failed", e)
PY
