#!/bin/bash
set -u
mkdir -p gpurun_out
for mb in 248000000; do
( while sleep 10; do nvidia-smi --query-gpu=utilization.gpu,memory.used,power.draw --format=csv,noheader; done ) &
MON=$!
timeout -s KILL 300 python bench.py --bases $mb --vecsize 512 --units 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02c_cfg5b_$mb.json 2> gpurun_out/r02c_cfg5b_$mb.err; echo "5b $mb rc=$?"
kill $MON
tail -5 gpurun_out/r02c_cfg5b_$mb.err; cut -c1-400 gpurun_out/r02c_cfg5b_$mb.json
done
