"""Summarise an .ncu-rep (read on the CPU box): key metrics + stall breakdown + executed-instruction mix.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("=" * 100)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-72s %-12s %s" % (k, units[i], r[i]))
    tot = 0
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            v = float(r[i].replace(",", "") or 0)
            st.append((v, h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            tot += v
    print("-- warp stall samples (pc sampling), total %d" % tot)
    for v, n in sorted(st, reverse=True)[:10]:
        print("   %-28s %8d  %5.1f%%" % (n, v, 100 * v / max(tot, 1)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
try:
    h2 = rows[1]
    isrc, iex = h2.index("Source"), h2.index("Instructions Executed")
    c = Counter()
    for r in rows[2:]:
        try:
            op = r[isrc].split()
            op = op[1] if op[0].startswith("@") else op[0]
            c[op.split(".")[0]] += int(r[iex] or 0)
        except Exception:
            pass
    tt = sum(c.values())
    print("-- executed warp instructions by opcode (first kernel in the report), total %d" % tt)
    for k, v in c.most_common(18):
        print("   %-10s %12d  %5.1f%%" % (k, v, 100.0 * v / tt))
except Exception as e:
    print("(no source page: %s)" % e)
