"""Summarise an ncu --set full report: one CSV row per launch with the metrics the roofline discussion uses."""
import csv
import subprocess
import sys

rep = sys.argv[1]
METRICS = ["launch__cluster_size", "launch__registers_per_thread", "gpu__time_duration.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
units = rows[1]
idx = [h.index(m) for m in METRICS if m in h]
w = csv.writer(sys.stdout)
w.writerow(["Kernel Name"] + [h[i] for i in idx])
w.writerow([""] + [units[i] for i in idx])
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    w.writerow([name[:60]] + [r[i] for i in idx])
