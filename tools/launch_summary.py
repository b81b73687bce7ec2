"""Aggregate an `ncu --metrics gpu__time_duration.sum[,...] --csv --log-file X.csv` launch list: kernel, launches, mean ns
(and mean DRAM bytes when those metrics were collected).
    python tools/launch_summary.py gpurun_out/launches.csv ["header comment"] > profiles/r02_launches_<tag>.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
per = collections.OrderedDict()  # launch id -> {kernel, metrics}
for r in rows[hi + 1:]:
    if len(r) > iv:
        per.setdefault(r[0], {"k": r[ik].split("(")[0]})[r[im]] = float(r[iv].replace(",", ""))
agg = collections.OrderedDict()
for v in per.values():
    a = agg.setdefault(v["k"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += v.get("gpu__time_duration.sum", 0.0)
    a[2] += v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("# per-launch times are cold-cache and serialised: compare SHARES.  kernel, launches, mean ns, mean DRAM bytes")
for k, (n, t, b) in agg.items():
    print("%-66s %4d %12.0f %14.0f" % (k[:66], n, t / n, b / n))
