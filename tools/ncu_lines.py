"""Per source line: executed warp instructions and stall samples of the first kernel in an .ncu-rep (needs -lineinfo and
--import-source on).   python tools/ncu_lines.py x.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ""
hdr = None
agg = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        iex, ist = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        try:
            agg.append((int(r[iex] or 0), int(r[ist] or 0), cur_file, int(r[0])))
        except ValueError:
            pass
tot = sum(a[0] for a in agg) or 1
tots = sum(a[1] for a in agg) or 1
src = {}
for a in sorted(agg, reverse=True)[:top]:
    f = a[2]
    if f not in src:
        try:
            import glob
            p = glob.glob("/root/repo/spectralelements.jl_b200/csrc/" + f)
            src[f] = open(p[0]).read().split("\n") if p else []
        except Exception:
            src[f] = []
    line = src[f][a[3] - 1].strip()[:90] if len(src[f]) >= a[3] else ""
    print("%5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100.0 * a[0] / tot, 100.0 * a[1] / tots, f, a[3], line))
