#!/bin/bash
# round 2, state O: tiled MSS scan; memcheck of the new kernels on small cases
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02o_tests.log
for tl in 1 0; do
DEEPGRP_KNOBS="mss_tiled=$tl" timeout -s KILL 200 python bench.py --shard chunk --bases 248000000 --steps 3 --warmup 1 > gpurun_out/r02o_chunk_tiled$tl.json 2> gpurun_out/r02o_chunk_tiled$tl.err; echo "tiled=$tl rc=$?"
done
python - <<'PY'
import json
for tl in (1, 0):
    try:
        d = json.load(open("gpurun_out/r02o_chunk_tiled%d.json" % tl))
        print("tiled", tl, d["stages_ms"], d["mss_rounds"], d["rows_per_step"])
    except Exception as e:
        print(tl, "failed", e)
PY
timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sections x4 > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02o_bench.json"))
print("value", d["value"], d["stages_ms"], "x4", d["x4"]["value"], d["x4"]["stages_ms"])
PY
timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "config5 or lstm_variant or test_fasta_stream_equals" > gpurun_out/r02o_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/r02o_memcheck.log
