"""Digest of an ncu report: headline metrics, stall mix, opcode mix, hot SASS windows."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    g = lambda n: r[hdr.index(n)] if n in hdr else None
    print("==", g("Kernel Name"))
    for n in ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
              "launch__registers_per_thread", "launch__occupancy_limit_registers",
              "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
              "sm__warps_active.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "smsp__thread_inst_executed_per_inst_executed.ratio",
              "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]:
        print(f"  {n:70s} {g(n)}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in data if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
if "--first" in sys.argv:   # several captured launches: keep the first kernel's rows only
    seen, keep = set(), []
    for r in data:
        if r[ix["Address"]] in seen:
            break
        seen.add(r[ix["Address"]])
        keep.append(r)
    data = keep
ts = sum(int(r[ix["# Samples"]]) for r in data)
ti = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("SASS instructions", len(data), "samples", ts, "warp instr", ti)
op, ops, thr = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    w = r[ix["Source"]].split()
    o = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
    op[o] += int(r[ix["Instructions Executed"]])
    ops[o] += int(r[ix["# Samples"]])
    thr[o] += int(r[ix["Thread Instructions Executed"]])
for o, c in op.most_common(18):
    print(f"  {o:10s} exec {c / ti:6.3f} samples {ops[o] / ts:6.3f} avg threads {thr[o] / max(c, 1):5.1f}")
for st in ["stall_barrier", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_math",
           "stall_long_sb", "stall_selected", "stall_branch_resolving", "stall_not_selected",
           "stall_dispatch", "stall_mio", "stall_lg"]:
    print(f"  {st:24s} {sum(int(r[ix[st]]) for r in data) / ts:6.3f}")
