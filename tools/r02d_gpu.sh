#!/bin/bash
# round 2, state D: streaming pipeline tests, the new bench line (x4 / strong / genome sections), ncu of the two forward kernels
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "stream" > gpurun_out/r02d_stream_tests.log 2>&1; echo "stream tests rc=$?"; tail -15 gpurun_out/r02d_stream_tests.log
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02d_tests.log
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r02d_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02d_bench.json"))
    print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["e2e"], "numa", d["numa"])
    for k in ("x4", "strong", "genome", "cpu_baseline"):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print("bench failed", e)
PY
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sections "" > gpurun_out/r02d_launches.log 2>&1; echo "ncu list rc=$?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gru_tc_attention -c 1 -o gpurun_out/r02d_fwd_tc \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --sections "" > gpurun_out/r02d_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:rnn_tcw -c 1 -o gpurun_out/r02d_fwd_tcw_5b \
  python bench.py --bases 24800000 --vecsize 512 --units 128 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02d_ncu_tcw.log 2>&1; echo "ncu tcw rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
nvidia-smi topo -m > gpurun_out/r02d_topo.txt 2>&1; lscpu | head -25 >> gpurun_out/r02d_topo.txt
