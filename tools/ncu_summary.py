"""Summarise an ncu report: key raw metrics + per-opcode / per-instruction stall samples.
usage: python tools/ncu_summary.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__registers_per_thread", "smsp__inst_executed.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__average_warp_latency_per_inst_issued.ratio"]
    for i, h in enumerate(hdr):
        if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
                         and float(vals[i] or 0) > 0.05) or h in (
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"):
            print("%-90s %-10s %s" % (h, units[i], vals[i]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr, data = rows[1], rows[2:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot_s = sum(int(r[isamp]) for r in data)
    tot_e = sum(int(r[iex]) for r in data)
    print("SASS instructions %d, samples %d, warp-instructions executed %d" % (len(data), tot_s, tot_e))
    by_s, by_e = Counter(), Counter()
    for r in data:
        parts = r[isrc].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = op.split(".")[0]
        by_s[op] += int(r[isamp])
        by_e[op] += int(r[iex])
    for op, c in by_s.most_common(18):
        print("  %-10s samples %.3f  executed %.3f" % (op, c / tot_s, by_e[op] / tot_e))
    print("top instructions by samples:")
    for i, r in sorted(enumerate(data), key=lambda t: -int(t[1][isamp]))[:topn]:
        print("  #%-5d %.4f  exec %-10s %s" % (i, int(r[isamp]) / tot_s, r[iex], r[isrc][:100]))


if __name__ == "__main__":
    main()
