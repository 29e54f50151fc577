"""Summarise an ncu report (read here, without a GPU): one CSV row per profiled launch with the
metrics the roofline discussion uses.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.csv"""
import csv
import subprocess
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = [hdr.index(c) if c in hdr else None for c in COLS]
w = csv.writer(sys.stdout)
w.writerow(COLS)
w.writerow([units[i] if i is not None else "" for i in idx])
for r in rows[2:]:
    w.writerow([r[i] if i is not None else "" for i in idx])
