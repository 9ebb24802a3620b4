"""Step time with the captured step graph vs plain launches (MOIRE_B200_NO_GRAPH=1), no per-kind
events.  usage: graph_ab.py [config] [chains]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moire_b200.engine import Engine
from moire_b200.workloads import make_workload

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
wl = make_workload(name)
temps = np.asarray(wl["temps"])
if len(sys.argv) > 2:
    temps = temps[:int(sys.argv[2])]
for mode in ("graph", "plain"):
    if mode == "plain":
        os.environ["MOIRE_B200_NO_GRAPH"] = "1"
    else:
        os.environ.pop("MOIRE_B200_NO_GRAPH", None)
    eng = Engine(wl["packed"], temps, seed=3, burnin=1000, **wl["model"])
    eng.init_chains()
    for it in range(5):
        eng.step(it)
    eng.scalars()
    ms, n = 0.0, 20
    t0 = time.perf_counter()
    for it in range(5, 5 + n):
        eng.timer_start()
        eng.step(it)
        ms += eng.timer_stop()
        eng.scalars()
    wall = time.perf_counter() - t0
    print(mode, name, len(temps), "chains: device ms/step", round(ms / n, 4), "wall ms/step", round(1e3 * wall / n, 4),
          "llik0", eng.scalars()[0][0])
    eng.close()
