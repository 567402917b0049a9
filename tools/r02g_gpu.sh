#!/bin/bash
# round 2, state G: stream host timing; ncu of the overlapped 5b kernel
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sections genome > gpurun_out/r02g_bench_genome.json 2> gpurun_out/r02g_bench_genome.err; echo "genome rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02g_bench_genome.json"))
    g = d["genome"]
    print({k: g[k] for k in ("value", "seconds", "rank_forward_seconds", "rank0_gpu_seconds", "rank0_waits_ms")})
    print("main e2e", d["e2e"])
except Exception as e:
    print("genome failed", e)
PY
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:rnn_tcw -c 1 -o gpurun_out/r02g_fwd_tcw_5b_ovl \
  python bench.py --bases 24800000 --vecsize 512 --units 128 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02g_ncu_tcw.log 2>&1; echo "ncu tcw rc=$?"
