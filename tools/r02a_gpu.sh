#!/bin/bash
# round 2, first GPU pass: the wide tcgen05 kernel (single CTA, then CTA pair), each under its own timeout
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL 420 "$@" > gpurun_out/r02a_$name.log 2>&1; echo "exit $? ($name)"; tail -n 15 gpurun_out/r02a_$name.log; }
run wide1 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "wide_kernel and (1-150 or 1-342 or 1-100)"
run lstm python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "lstm"
run slabs python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "slabs"
run wide2 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "wide_kernel and (2-342 or 2-150)"
run cfg5 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "config5"
nvidia-smi --query-gpu=name,memory.used --format=csv
