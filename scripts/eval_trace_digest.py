"""Digest of MOIRE_B200_TRACE written by a -DMOIRE_EVAL_TRACE build: the last k_flat_eval launch,
per block {entry, n words, then per task: previous task done at, claimed at, task id}.
usage: eval_trace_digest.py trace.bin"""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(1024, 256).astype(np.int64)
blocks = [b for b in range(600) if a[b, 0] > 0 and a[b, 1] >= 5]
t0 = min(a[b, 0] for b in blocks)
rows = []
for b in blocks:
    n = int(a[b, 1])
    rec = a[b, 2:n].reshape(-1, 3)            # done_prev, claimed, task
    entry = a[b, 0] - t0
    first_claim = rec[0, 1] - t0
    claim_gap = (rec[:, 1] - rec[:, 0])       # barrier + atomic + barrier
    work = rec[1:, 0] - rec[:-1, 1]           # task durations (block level)
    end = rec[-1, 1] - t0
    rows.append((entry, first_claim, claim_gap.sum(), work.sum(), end, len(rec) - 1, work.max() if len(work) else 0,
                 rec[np.argmax(work), 2] if len(work) else -1))
r = np.array(rows, dtype=np.float64)
us = 1e-3
print(f"{len(blocks)} blocks; launch spans {r[:, 4].max() * us:.1f} us (first block entry to last block exit)")
print(f"block entry:       mean {r[:, 0].mean() * us:6.2f}  max {r[:, 0].max() * us:6.2f} us after the first")
print(f"first task claimed mean {r[:, 1].mean() * us:6.2f}  max {r[:, 1].max() * us:6.2f} us (prologue: tables into shared memory)")
print(f"claim gaps / block mean {r[:, 2].mean() * us:6.2f}  max {r[:, 2].max() * us:6.2f} us over {r[:, 5].mean():.1f} tasks per block ({r[:, 2].sum() / max(1, r[:, 5].sum() + len(blocks)) * us:.2f} us per claim)")
print(f"work / block       mean {r[:, 3].mean() * us:6.2f}  max {r[:, 3].max() * us:6.2f} us")
print(f"block exit         mean {r[:, 4].mean() * us:6.2f}  min {r[:, 4].min() * us:6.2f}  max {r[:, 4].max() * us:6.2f} us")
order = np.argsort(-r[:, 6])[:8]
print("longest single tasks (us, task id, block exit us):", [(round(r[i, 6] * us, 1), int(r[i, 7]), round(r[i, 4] * us, 1)) for i in order])
# task duration by task id decile
alltasks = []
for b in blocks:
    n = int(a[b, 1]); rec = a[b, 2:n].reshape(-1, 3)
    for i in range(len(rec) - 1):
        alltasks.append((rec[i, 2], rec[i + 1, 0] - rec[i, 1], rec[i, 1] - t0))
alltasks.sort()
at = np.array(alltasks, dtype=np.float64)
if len(at):
    for lo in range(0, 10):
        seg = at[len(at) * lo // 10: len(at) * (lo + 1) // 10]
        print(f"  tasks {int(seg[0, 0]):6d}..{int(seg[-1, 0]):6d}: mean {seg[:, 1].mean() * us:6.2f} us max {seg[:, 1].max() * us:6.2f}  started {seg[:, 2].min() * us:6.1f}..{seg[:, 2].max() * us:6.1f} us")
# cooperative cells' phase clocks (cycles): {k, cnt, normalise, subset terms + ordered sums, gather, mixture}
c = a.reshape(-1)[600 * 256:]
n = int(c[0])
if n:
    r = c[8:8 + 8 * min(n, 8000)].reshape(-1, 8).astype(np.float64)
    print(f"cooperative cells traced: {len(r)} (of {n}); cycles per phase, mean by k:")
    for k in sorted(set(r[:, 0].astype(int))):
        s = r[r[:, 0] == k]
        print(f"  k {k:2d}: {len(s):5d} cells, cnt mean {s[:, 1].mean():5.1f}  normalise {s[:, 2].mean():7.0f}  subsets+sums {s[:, 3].mean():8.0f}  gather {s[:, 4].mean():6.0f}  mixture {s[:, 5].mean():7.0f}  total {s[:, 2:6].sum(1).mean():8.0f} (max {s[:, 2:6].sum(1).max():8.0f})")
