"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
one line per launch (kernel, grid, ms, GB read, GB written)."""
import csv, sys, re
rows = {}
order = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    i = int(r["ID"])
    if i not in rows:
        rows[i] = {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), "grid": r["Grid Size"], "block": r["Block Size"]}
        order.append(i)
    rows[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
tot = 0
for i in order:
    r = rows[i]
    ms = r.get("gpu__time_duration.sum", 0) / 1e6
    tot += ms
    print("%3d %-34s %-22s %9.3f ms  rd %7.2f GB  wr %7.2f GB" % (i, r["name"][:34], r["grid"] + r["block"], ms,
          r.get("dram__bytes_read.sum", 0) / 1e9, r.get("dram__bytes_write.sum", 0) / 1e9))
print("total %.1f ms" % tot)
