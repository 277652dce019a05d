"""Key metrics of an ncu --set full capture (read on the CPU box): python scripts/ncu_summary.py x.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active", "lts__t_bytes.sum ",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "local_load", "smsp__inst_executed_op_local",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]


def main():
    txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(w.strip()) for w in WANT):
            if h.startswith("smsp__average_warps_issue_stalled") and float(v.replace(",", "")) < 0.05:
                continue
            print("%-86s %-10s %s" % (h, u, v))


if __name__ == "__main__":
    main()
