#!/bin/bash
# round 2, state N: launch list of the chunk-sharded step at 248 Mbp (which kernels make up the owner's serial "finish")
set -u
mkdir -p gpurun_out
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02n_launches_248.csv \
  python bench.py --shard chunk --bases 248000000 --steps 1 --warmup 1 > gpurun_out/r02n_launches.log 2>&1; echo "ncu rc=$?"
python tools/launch_table.py gpurun_out/r02n_launches_248.csv 2>&1 | head -50
