#!/bin/bash
# round 2, state M: end-to-end leg at N = 8 -- host piece size of the stream against the one-shot call
set -u
mkdir -p gpurun_out
run() { tag=$1; shift; timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 "$@" > gpurun_out/r02m_$tag.json 2> gpurun_out/r02m_$tag.err; echo "$tag rc=$?"; }
DEEPGRP_KNOBS="stream_slot_mb=64" run s64 --sections ""
DEEPGRP_KNOBS="stream_slot_mb=256" run s256 --sections "genome"
DEEPGRP_KNOBS="stream_slot_mb=1024" run s1024 --sections ""
run oneshot --sections "" --e2e-api oneshot
python - <<'PY'
import json
for tag in ("s64", "s256", "s1024", "oneshot"):
    try:
        d = json.loads([l for l in open("gpurun_out/r02m_%s.json" % tag) if l.startswith("{")][-1])
        print(tag, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],1))
        g = d.get("genome")
        if g: print("  genome", g["value"], g["seconds"], g["rank_seconds"], g["rank0_waits_ms"])
    except Exception as e:
        print(tag, "failed", e)
PY
