#!/bin/bash
# round 2, state W: the driver's command at N = 8 with the final kernels (weak + x4 + strong + config5b + genome)
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02w_bench_n8.json 2> gpurun_out/r02w_bench_n8.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/r02w_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02w_bench_n8.json") if l.startswith("{")][-1])
    print("value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],1))
    for k in ("x4", "strong", "config5b", "genome"):
        v = d.get(k) or {}
        print(k, {kk: v[kk] for kk in v if kk in ("value", "e2e", "ms_per_step", "e2e_ms_per_step", "stages_ms", "seconds", "rank_seconds", "rows_per_step", "mss_rounds", "n_gpus", "scale", "roofline")})
except Exception as e:
    print("bench failed", e)
PY
timeout -s KILL 300 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "abandoned or stream" 2>&1 | tail -3
