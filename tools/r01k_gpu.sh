#!/bin/bash
# Round 1, state K: one box, N = 1.  Parity tests, the plain bench line, the ncu launch list of the same command,
# then the remaining BASELINE.json configs at N = 1 (config 5a / 5b through the chunk route, config 1).
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q > gpurun_out/r01k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r01k_tests.log
timeout 70 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01k_bench.json 2> gpurun_out/r01k_bench.err; echo "bench rc=$?"
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01k_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01k_launches.log 2>&1; echo "ncu rc=$?"
timeout 45 python bench.py --shard chunk --bases 248000000 --vecsize 260 --units 50 --steps 2 --warmup 3 > gpurun_out/r01k_cfg5a_n1.json 2> gpurun_out/r01k_cfg5a.err; echo "5a rc=$?"
timeout 45 python bench.py --shard chunk --bases 24800000 --vecsize 512 --units 128 --steps 1 --warmup 3 > gpurun_out/r01k_cfg5b_n1.json 2> gpurun_out/r01k_cfg5b.err; echo "5b rc=$?"
timeout 40 python bench.py --bases 1000000 --vecsize 150 --units 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01k_cfg1.json 2> gpurun_out/r01k_cfg1.err; echo "cfg1 rc=$?"
cut -c1-300 gpurun_out/r01k_bench.json
