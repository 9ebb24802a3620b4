"""Join an ncu SASS profile of one kernel with nvdisasm line info: executed warp instructions,
thread efficiency and stall samples per source line (inlined-at chains collapsed to the innermost
line).  usage: ncu_lines.py report.ncu-rep libmoire_b200.so mangled_kernel_substring [top]
NCU_FILTER="-k regex:name -s N -c 1" picks one result of a multi-kernel report."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = ""
for cub in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    if kname in out:
        dis = out
        break
lines, cur, active = [], None, False
for ln in dis.splitlines():
    if ".section" in ln and ".text." in ln:
        active = kname in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        if "inlined at" not in ln or cur is None or True:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines.append((int(m.group(1), 16), m.group(2), cur))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + os.environ.get("NCU_FILTER", "").split(), capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tables, cur_t = [], None
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    if r[0] == "Address":
        cur_t = []
        tables.append(cur_t)
    elif cur_t is not None:
        cur_t.append(r)
data = tables[int(os.environ.get("NCU_TABLE", "0"))]  # a multi-result report: which table
base = int(data[0][ix["Address"]], 16)
byoff = {off: loc for off, _, loc in lines}
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
ti = ts = 0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    loc = byoff.get(off, ("?", 0))
    a = agg[loc]
    e, th, s = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]])
    a[0] += e; a[1] += th; a[2] += s; a[3] += int(r[ix["stall_barrier"]]); a[4] += int(r[ix["stall_no_inst"]])
    ti += e; ts += s
print(f"{'file:line':28s} {'exec%':>7s} {'thr/inst':>8s} {'samp%':>7s} {'barrier%':>8s} {'noinst%':>8s}")
key = int(os.environ.get("SORTCOL", "0"))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
    print(f"{loc[0] + ':' + str(loc[1]):28s} {100 * a[0] / ti:7.2f} {a[1] / max(a[0], 1):8.1f} {100 * a[2] / ts:7.2f} {100 * a[3] / ts:8.2f} {100 * a[4] / ts:8.2f}")
# ---- thread-instruction share by code region (line ranges of the current sources)
if os.environ.get("REGIONS"):
    regs = [tuple(x.split(":")) for x in os.environ["REGIONS"].split(",")]  # name:file:lo:hi
    tot = sum(a[1] for a in agg.values())
    out = collections.Counter()
    for (f, l), a in agg.items():
        for name, rf, lo, hi in regs:
            if f == rf and int(lo) <= l <= int(hi):
                out[name] += a[1]
                break
        else:
            out["other:" + f] += a[1]
    for k, vv in out.most_common():
        print(f"  {k:30s} {100 * vv / tot:6.2f}% of thread instructions")
