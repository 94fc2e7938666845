"""Condense an `ncu --page raw --csv` export to the handful of columns the profiles/ summaries quote.
usage: python tools/ncu_condense.py raw.csv > condensed.csv"""
import csv, sys

COLS = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__cluster_size",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "smsp__cycles_active.avg"]
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
keep = [hdr.index(c) for c in COLS if c in hdr]
w = csv.writer(sys.stdout)
for r in rows[h:]:
    w.writerow([r[i] if i < len(r) else "" for i in keep])
