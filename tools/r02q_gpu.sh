#!/bin/bash
# round 2, state Q (HEAD): the driver's N = 1 command, the reference arm, the other BASELINE configs at N = 1, launch list
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02q_bench_ref.json 2> gpurun_out/r02q_bench_ref.err; echo "ref rc=$?"
timeout -s KILL 200 python bench.py --bases 248000000 --vecsize 512 --units 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02q_cfg5b_248.json 2> gpurun_out/r02q_cfg5b_248.err; echo "5b rc=$?"
timeout -s KILL 200 python bench.py --bases 248000000 --vecsize 260 --units 50 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02q_cfg5a_248.json 2> gpurun_out/r02q_cfg5a_248.err; echo "5a rc=$?"
timeout -s KILL 200 python bench.py --bases 1000000 --vecsize 150 --units 32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_cfg1.json 2> gpurun_out/r02q_cfg1.err; echo "cfg1 rc=$?"
timeout -s KILL 200 python bench.py --bases 46700000 --rnn LSTM --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02q_lstm.json 2> gpurun_out/r02q_lstm.err; echo "lstm rc=$?"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02q_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sections "" > gpurun_out/r02q_launches.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json
for f in ("bench", "cfg5b_248", "cfg5a_248", "cfg1", "lstm"):
    try:
        d = json.load(open("gpurun_out/r02q_%s.json" % f))
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "TF", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],4), d["stages_ms"], "rounds", d["mss_rounds"], d["clocks"]["sm_mhz"])
        for k in ("x4", "strong", "genome", "cpu_baseline"):
            if k in d: print("   ", k, json.dumps(d[k])[:400])
    except Exception as e:
        print(f, "failed", e)
print(open("gpurun_out/r02q_bench_ref.json").read()[:900])
PY
