"""Trim an `ncu --metrics gpu__time_duration.sum --csv` log into id,kernel,grid,duration rows."""
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi, gi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Metric Unit")
w = csv.writer(sys.stdout)
w.writerow(["id", "kernel", "grid", "gpu__time_duration.sum", "unit"])
for r in rows[1:]:
    w.writerow([r[0], r[ki].split("(")[0], r[gi], r[vi], r[ui]])
