"""Kernel sequence of the last update_p in an ncu launch list, one line; usage: seq_summary.py launches.csv [n]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1]))); hdr = None; seq = []
for r in rows:
    if hdr is None:
        if "Kernel Name" in r: hdr = {h: i for i, h in enumerate(r)}
        continue
    if len(r) < len(hdr): continue
    n = re.sub(r"\(.*", "", r[hdr["Kernel Name"]]).replace("void moire::", "").replace("moire::", "")
    v = float(r[hdr["Metric Value"]].replace(",", "")); u = r[hdr["Metric Unit"]]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    seq.append((n, v))
idx = [i for i, (n, v) in enumerate(seq) if n == 'k_pf_prepare'][-1]
cnt = int(sys.argv[2]) if len(sys.argv) > 2 else 45
print(' '.join(f"{n.replace('k_pf_','').replace('k_flat_','')}:{v:.0f}" for n, v in seq[idx:idx + cnt]))
