"""One line per captured launch of an ncu report: duration, DRAM bytes and throughput, issue slots,
fp64 pipe, lanes per instruction, L2 hit rate, top stall reasons.  usage: ncu_raw_digest.py rep.ncu-rep"""
import csv
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print(f"{'kernel':34s} {'us':>7s} {'grid':>6s} {'blk':>4s} {'reg':>3s} {'rdMB':>6s} {'wrMB':>6s} {'dram%':>5s} {'issue%':>6s} {'fp64%':>5s} {'lanes':>5s} {'L2hit':>5s} {'warps%':>6s}  top stalls")
for r in rows[2:]:
    g = lambda n: float(r[ix[n]].replace(",", "")) if n in ix and r[ix[n]] not in ("", "n/a") else float("nan")
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("moire::", "")
    dur = g("gpu__time_duration.sum")
    unit = rows[1][ix["gpu__time_duration.sum"]]
    dur = dur / 1e3 if unit == "ns" else dur * 1e3 if unit == "ms" else dur
    def mb(n):
        v = g(n); u = rows[1][ix[n]]
        return v / 1e6 if u == "byte" else v / 1e3 if u == "Kbyte" else v if u == "Mbyte" else v * 1e3 if u == "Gbyte" else v
    st = sorted(((g(s), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls), reverse=True)
    tot = sum(v for v, _ in st if v == v) or 1.0
    top = " ".join(f"{n}:{100 * v / tot:.0f}" for v, n in st[:4])
    print(f"{name[:34]:34s} {dur:7.1f} {int(g('launch__grid_size')):6d} {int(g('launch__block_size')):4d} {int(g('launch__registers_per_thread')):3d} "
          f"{mb('dram__bytes_read.sum'):6.1f} {mb('dram__bytes_write.sum'):6.1f} {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} {g('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):5.1f} "
          f"{g('smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f} {g('lts__t_sector_hit_rate.pct'):5.1f} {g('sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f}  {top}")
