#!/bin/bash
# Round 1, state L: long-segment gap fill.  Parity tests, config 5b through the chunk route (its whole record is
# one maximal segment), the plain bench line.
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q > gpurun_out/r01l_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r01l_tests.log
timeout 40 python bench.py --shard chunk --bases 24800000 --vecsize 512 --units 128 --steps 1 --warmup 3 > gpurun_out/r01l_cfg5b_n1.json 2> gpurun_out/r01l_cfg5b.err; echo "5b rc=$?"
timeout 40 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01l_bench.json 2> gpurun_out/r01l_bench.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/r01l_bench.json; grep -o '"stages_ms.*' gpurun_out/r01l_cfg5b_n1.json | cut -c1-300
