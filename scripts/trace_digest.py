"""Digest of MOIRE_B200_TRACE: per (block, proposal) the spread of the warps' evaluation-phase end
times in update_p."""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(64, 16, 16, 2).astype(np.int64)
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rows = []
for b in range(64):
    for r in range(10):
        t0, t1 = a[b, r, :nw, 0], a[b, r, :nw, 1]
        if t1.min() == 0:
            continue
        dur = t1 - t0.min()
        rows.append((dur.max(), dur.mean(), dur.min(), np.sort(dur)))
rows.sort(key=lambda x: -x[0])
mx = np.array([r[0] for r in rows]); mean = np.array([r[1] for r in rows]); mn = np.array([r[2] for r in rows])
print(f"{len(rows)} (block, proposal) phases; cycles: slowest warp mean {mx.mean():.0f}, average warp {mean.mean():.0f}, fastest {mn.mean():.0f}")
print(f"waiting share = 1 - mean/max = {1 - (mean / mx).mean():.3f}")
for r in rows[:5] + rows[len(rows) // 2:len(rows) // 2 + 3]:
    print("  warps' durations:", (r[3] / 1000).round(1).tolist(), "kcycles")
