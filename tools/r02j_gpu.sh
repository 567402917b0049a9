#!/bin/bash
# round 2, state J: new parity tests (1 Mbp config-2 shape vs oracle, config-1 TSV row diff), MSS chunk tuning
set -u
mkdir -p gpurun_out
DGRP_TSV_DIFF_REPORT=gpurun_out/r02j_tsv_diff_config1.json timeout -s KILL 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -s -k "config2_shape_vs_oracle or tsv_row_diff" > gpurun_out/r02j_parity.log 2>&1; echo "parity rc=$?"; grep -v "^$" gpurun_out/r02j_parity.log | tail -12
for ch in 256 512 1024 2048 4096; do
DEEPGRP_KNOBS="mss_chunk=$ch" timeout -s KILL 200 python bench.py --shard chunk --bases 248000000 --steps 2 --warmup 1 > gpurun_out/r02j_chunk_mss$ch.json 2> gpurun_out/r02j_chunk_mss$ch.err; echo "mss_chunk=$ch rc=$?"
done
python - <<'PY'
import json, glob
for ch in (256, 512, 1024, 2048, 4096):
    try:
        d = json.load(open("gpurun_out/r02j_chunk_mss%d.json" % ch))
        print(ch, d["stages_ms"], d["mss_rounds"])
    except Exception as e:
        print(ch, "failed", e)
PY
