#!/bin/bash
# round 2, state V: MSS chunk size on the 1 Mbp record of config 1; ncu of the final two-tile kernel; launch list
set -u
mkdir -p gpurun_out
for ch in 0 256 512 1024 2048 4096 8192; do
DEEPGRP_KNOBS="mss_chunk=$ch" timeout -s KILL 100 python bench.py --bases 1000000 --vecsize 150 --units 32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02v_cfg1_ch$ch.json 2> gpurun_out/r02v_cfg1_ch$ch.err
done
python - <<'PY'
import json
for ch in (0, 256, 512, 1024, 2048, 4096, 8192):
    try:
        d = json.load(open("gpurun_out/r02v_cfg1_ch%d.json" % ch))
        print("chunk", ch, "mss ms", round(d["stages_ms"]["mss_ms"], 3), "rounds", d["mss_rounds"], "total", round(d["stages_ms"]["total_ms"], 3), "value", round(d["value"], 1))
    except Exception as e:
        print(ch, "failed", e)
PY
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gru_tc_attention -c 1 -o gpurun_out/r02v_fwd_tc \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --sections "" > gpurun_out/r02v_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02v_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sections "" > gpurun_out/r02v_launches.log 2>&1; echo "ncu list rc=$?"
