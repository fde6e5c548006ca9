"""Print a per-launch table from an `ncu --csv` log (metrics as columns)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d, order = {}, []
for r in rows[1:]:
    k = (int(r[ii]), r[ki])
    if k not in d:
        d[k] = {}
        order.append(k)
    d[k][r[mi]] = r[vi].replace(",", "")
tot = 0.0
for k in order:
    m = d[k]
    t = float(m.get("gpu__time_duration.sum", 0)) / 1e3
    tot += t
    tp = m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "")
    rd = float(m.get("dram__bytes_read.sum", 0)) / 1e6 if "dram__bytes_read.sum" in m else float("nan")
    wr = float(m.get("dram__bytes_write.sum", 0)) / 1e6 if "dram__bytes_write.sum" in m else float("nan")
    name = k[1].split("(")[0][-70:]
    print(f"{k[0]:4d} {t:9.1f} us  tensor {tp:>6}%  rd {rd:8.1f} MB  wr {wr:8.1f} MB  {name}")
print(f"total {tot:.1f} us over {len(order)} launches")
