#!/bin/bash
# round 2, state F: overlapped MMA / gate protocol of the wide kernel; stream wait diagnostics
set -u
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "wide or lstm or config5 or stream" > gpurun_out/r02f_tests_wide.log 2>&1; echo "wide tests rc=$?"; tail -4 gpurun_out/r02f_tests_wide.log
for ov in 1 0; do
DEEPGRP_KNOBS="forward_overlap=$ov" timeout -s KILL 200 python bench.py --bases 24800000 --vecsize 512 --units 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02f_cfg5b_ov$ov.json 2> gpurun_out/r02f_cfg5b_ov$ov.err; echo "5b ov=$ov rc=$?"
DEEPGRP_KNOBS="forward_overlap=$ov" timeout -s KILL 200 python bench.py --bases 46700000 --vecsize 342 --units 60 --rnn LSTM --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02f_lstm_ov$ov.json 2> gpurun_out/r02f_lstm_ov$ov.err; echo "lstm ov=$ov rc=$?"
done
timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sections genome > gpurun_out/r02f_bench_genome.json 2> gpurun_out/r02f_bench_genome.err; echo "genome rc=$?"
python - <<'PY'
import json
for f in ("cfg5b_ov1", "cfg5b_ov0", "lstm_ov1", "lstm_ov0"):
    try:
        d = json.load(open("gpurun_out/r02f_%s.json" % f))
        print(f, "value", round(d["value"],1), "fwd ms", round(d["stages_ms"]["forward_ms"],1), "TF", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],4), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
try:
    d = json.load(open("gpurun_out/r02f_bench_genome.json"))
    print(json.dumps(d["genome"])[:1500])
except Exception as e:
    print("genome failed", e)
PY
