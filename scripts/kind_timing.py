"""Per-update-kind device time of one engine step (CUDA events around every launch)."""
import argparse
import json
import sys
import os
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moire_b200.workloads import make_workload  # noqa: E402
from moire_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--chains", type=int, default=None)
ap.add_argument("--pipeline", default="auto")
ap.add_argument("--mask-words", type=int, default=None,
                help="2: run the data on the two-word build (libmoire_b200_w128.so), zero-padded masks")
a = ap.parse_args()
wl = make_workload(a.config)
if a.mask_words:
    wl["packed"] = wl["packed"].widened(a.mask_words)
temps = wl["temps"] if a.chains is None else wl["temps"][:a.chains]
eng = Engine(wl["packed"], temps, seed=1, pipeline=a.pipeline, **wl["model"])
t0 = time.time()
info = eng.init_chains()
print("init s", time.time() - t0, "ill", info["ill_count"].sum())
for it in range(a.warmup):
    eng.step(it)
eng.synchronize()
ev0, _ = eng.counters()
eng.set_timing(True)
t0 = time.time()
for it in range(a.warmup, a.warmup + a.steps):
    eng.step(it)
eng.synchronize()
wall = time.time() - t0
tm = eng.timing()
ev1, _ = eng.counters()
tot = sum(v[0] for v in tm.values())
print(json.dumps(dict(config=a.config, mask_words=wl["packed"].mask_words, pipeline=a.pipeline, chains=len(temps), steps=a.steps, wall_ms_per_step=1e3 * wall / a.steps,
                      device_ms_per_step=tot / a.steps, evals_per_step=(ev1 - ev0) / a.steps,
                      kinds={k: round(v[0] / a.steps, 4) for k, v in tm.items()})))
llik, prior = eng.scalars()
print("llik[0..3]", llik[:4], "mean m", eng.dump_state(0)["m"].mean())
