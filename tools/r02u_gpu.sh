#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02u_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02u_tests.log
for v in 1 0; do
DEEPGRP_KNOBS="forward_smem_vote=$v" timeout -s KILL 200 python bench.py --bases 1000000 --vecsize 150 --units 32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02u_cfg1_sv$v.json 2> gpurun_out/r02u_cfg1_sv$v.err; echo "cfg1 sv=$v rc=$?"
DEEPGRP_KNOBS="forward_smem_vote=$v" timeout -s KILL 200 python bench.py --bases 46700000 --rnn LSTM --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02u_lstm_sv$v.json 2> gpurun_out/r02u_lstm_sv$v.err; echo "lstm sv=$v rc=$?"
done
python - <<'PY'
import json
for f in ("cfg1_sv1", "cfg1_sv0", "lstm_sv1", "lstm_sv0"):
    try:
        d = json.load(open("gpurun_out/r02u_%s.json" % f))
        print(f, "total ms", round(d["stages_ms"]["total_ms"], 3), d["stages_ms"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
    except Exception as e:
        print(f, "failed", e)
PY
