# Round 2, second session: the gpurun commands behind profiles/r02s2_* (one call each; `bash tools/r02s2_gpu.sh <step>` on the box).
# Files land in gpurun_out/ and were copied to profiles/ under the names given in profiles/r02s2_summary.md.
set -u
mkdir -p gpurun_out
B="python bench.py"
case "${1:-final}" in
  early_ab)      # r02s2_e2e_early0.json / r02s2_e2e_early1_first.json: the e2e leg with and without early rows, one box
    $B --sections x4 > gpurun_out/bench.json
    $B --sections '' --no-cpu-baseline --steps 10 --early-rows 0 > gpurun_out/bench_early0.json
    for v in "3 45" "4 55" "4 70" "5 55"; do set -- $v; $B --sections '' --no-cpu-baseline --steps 10 --early-slabs $1 --early-ratio $2 > gpurun_out/bench_s$1_r$2.json; done ;;
  chunk)         # r02s2_mss_chunk*.json
    for v in 0 128 256 512; do $B --sections '' --no-cpu-baseline --steps 10 --mss-chunk $v > gpurun_out/bench_ch$v.json; done ;;
  x4)            # r02s2_x4_*.json
    $B --sections '' --no-cpu-baseline --steps 5 --weight-scale 4 > gpurun_out/bench_x4.json
    $B --sections '' --no-cpu-baseline --steps 5 --weight-scale 4 --early-rows 0 > gpurun_out/bench_x4_e0.json ;;
  launches)      # r02s2_launches.csv, r02s2_launches_248.csv
    ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $B --steps 2 --warmup 1 --sections '' --no-cpu-baseline
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_248.csv $B --shard chunk --bases 248000000 --steps 2 --warmup 1 --no-cpu-baseline ;;
  hbm_ncu)       # r02s2_hbm_kernels_ncu_raw.csv
    ncu --set full --clock-control none --import-source on -k regex:"mss_scan_kernel|segv_scatter_kernel|segv_count_kernel|mss_count_runs4_kernel" -c 6 -o gpurun_out/hbm_kernels $B --sections '' --no-cpu-baseline --steps 1 --warmup 1
    ncu -i gpurun_out/hbm_kernels.ncu-rep --page raw --csv > gpurun_out/hbm_kernels_raw.csv ;;
  n2|n8)         # r02s2_bench_n2.json / r02s2_bench_n8.json (gpurun --gpus 2 / 8)
    N=${1#n}
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json ;;
  final)         # r02s2_final_*: what the driver runs at round end
    python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
    $B --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json
    $B > gpurun_out/bench.json
    ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $B --steps 2 --warmup 1 --sections '' --no-cpu-baseline ;;
esac
