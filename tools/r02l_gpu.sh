#!/bin/bash
# round 2, state L: the driver's command at N = 8 (weak line + x4 + strong + genome sections)
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02l_bench_n8.json 2> gpurun_out/r02l_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r02l_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02l_bench_n8.json") if l.startswith("{")][-1])
    print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["e2e"]["ms_per_step"], "numa", d["numa"])
    for k in ("x4", "strong", "genome"):
        v = d.get(k) or {}
        print(k, {kk: v[kk] for kk in v if kk in ("value", "e2e", "ms_per_step", "e2e_ms_per_step", "stages_ms", "seconds", "rank_seconds", "rank_forward_seconds", "rows_per_step", "mss_rounds", "n_gpus", "scale", "generate_seconds", "rank0_waits_ms")})
except Exception as e:
    print("bench failed", e)
PY
nvidia-smi topo -m > gpurun_out/r02l_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r02l_topo.txt; free -g >> gpurun_out/r02l_topo.txt
