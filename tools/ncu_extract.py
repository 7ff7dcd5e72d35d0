"""Reduce `ncu -i X.ncu-rep --page raw --csv` to the columns the profile notes cite (one row per launch):
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python tools/ncu_extract.py > profiles/name_raw.csv"""
import csv, sys
KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
STALL = "smsp__pcsamp_warps_issue_stalled_"
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith(STALL) and not h.endswith("_not_issued")]
cols = [k for k in KEEP if k in idx] + stalls
w = csv.writer(sys.stdout)
w.writerow([c.replace(STALL, "stall_") for c in cols])
for r in rows[1:]:
    w.writerow([r[idx[c]] if idx[c] < len(r) else "" for c in cols])
