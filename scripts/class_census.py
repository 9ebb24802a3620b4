"""Which cells set an evaluator launch's critical path: (k, m - k + 1) census of a workload's chain
states after a few steps.  usage: class_census.py c2 [steps]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moire_b200.workloads import make_workload
from moire_b200.engine import Engine
name = sys.argv[1]; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
wl = make_workload(name)
temps = wl["temps"]
eng = Engine(wl["packed"], temps, seed=1, **wl["model"])
eng.init_chains()
for it in range(steps):
    eng.step(it)
pop = np.vectorize(lambda x: bin(int(x)).count("1"))
for c in (0, len(temps) // 2, len(temps) - 1):
    st = eng.dump_state(c)
    k = pop(st["latent"]); m = st["m"][:, None] + 0 * k
    cnt = m - k + 1
    miss = wl["packed"].is_missing != 0
    k, cnt, mm = k[~miss], cnt[~miss], m[~miss]
    print(f"chain {c} temp {temps[c]:.3f}: m mean {st['m'].mean():.2f} max {st['m'].max()}  k mean {k.mean():.2f} max {k.max()}  cnt max {cnt.max()}")
    for lo, hi in ((1, 2), (3, 5), (6, 7), (8, 10), (11, 64)):
        sel = (k >= lo) & (k <= hi)
        if sel.any():
            print(f"   k {lo}-{hi}: {sel.sum()} cells, cnt mean {cnt[sel].mean():.1f} max {cnt[sel].max()}")
