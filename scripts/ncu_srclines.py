"""Per-source-line totals from ncu's own CUDA<->SASS correlation (needs --import-source on).
usage: ncu_srclines.py report.ncu-rep kernel_regex [table_index] [top]"""
import collections
import csv
import subprocess
import sys

rep, rx = sys.argv[1:3]
tab = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "-k", "regex:" + rx], capture_output=True, text=True).stdout
tables, cur, hdr = [], None, None
for r in csv.reader(out.splitlines()):
    if len(r) > 10 and r[0] == "Line No":
        hdr, cur = r, []
        tables.append((hdr, cur))
    elif cur is not None and hdr is not None and len(r) == len(hdr):
        cur.append(r)
hdr, rows = tables[tab]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


agg = collections.OrderedDict()
ti = ts = 0
for r in rows:
    key = (r[0], r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0, 0, 0])
    e, th, s = num(r[ix["Instructions Executed"]]), num(r[ix["Thread Instructions Executed"]]), num(r[ix["# Samples"]])
    a[0] += e; a[1] += th; a[2] += s; a[3] += num(r[ix["stall_long_sb"]])
    ti += e; ts += s
print(f"tables {len(tables)}; warp instructions {ti}, samples {ts}")
print(f"{'exec%':>6s} {'thr':>5s} {'samp%':>6s} {'lsb%':>6s}  line")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/max(ti,1):6.2f} {a[1]/max(a[0],1):5.1f} {100*a[2]/max(ts,1):6.2f} {100*a[3]/max(ts,1):6.2f}  {ln}: {src}")
