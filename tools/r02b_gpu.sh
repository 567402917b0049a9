#!/bin/bash
# round 2, state B: whole GPU suite, the plain bench line, config 5b / LSTM through the wide tcgen05 kernel
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02b_tests.log
timeout -s KILL 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
timeout -s KILL 120 python bench.py --bases 248000000 --vecsize 512 --units 128 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_cfg5b_n1.json 2> gpurun_out/r02b_cfg5b.err; echo "5b rc=$?"
timeout -s KILL 120 python bench.py --bases 46700000 --vecsize 342 --units 60 --rnn LSTM --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_lstm_n1.json 2> gpurun_out/r02b_lstm.err; echo "lstm rc=$?"
timeout -s KILL 120 python bench.py --bases 46700000 --vecsize 150 --units 32 --rnn LSTM --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_lstm32_n1.json 2> gpurun_out/r02b_lstm32.err; echo "lstm32 rc=$?"
for f in bench cfg5b_n1 lstm_n1 lstm32_n1; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02b_$f.json"))
    print("$f", round(d["value"],1), "Mbp/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), "; roofline", round(d["roofline"]["achieved"],1), "TF", round(d["roofline"]["frac"],4), d["roofline"]["kernel"][:40], d["stages_ms"], d["clocks"])
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r02b_$f.err").read()[-1500:])
PY
done
