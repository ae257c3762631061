"""Summarise an .ncu-rep (read here, no GPU needed) into one CSV row per kernel launch with the counters quoted in
DESIGN.md.  usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.csv"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [hdr.index(k) for k in KEYS if k in hdr]
kn = hdr.index("Kernel Name")
w = csv.writer(sys.stdout)
w.writerow(["Kernel Name"] + [hdr[c] for c in cols])
w.writerow([""] + [units[c] for c in cols])
for r in data:
    w.writerow([r[kn].split("(")[0]] + [r[c] for c in cols])
