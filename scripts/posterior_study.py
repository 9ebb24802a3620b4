#!/usr/bin/env python
"""Posterior equivalence across seeds (BASELINE.json's second correctness criterion) at the hard
shapes: the engine's `run_mcmc` -- as shipped (fp64) and with MOIRE_NUMERICS_REF32 (the float
reference's clamp emulated) -- against the reference's own C++ sampler, fp32 as shipped and its
mechanical fp64 variant (tests/golden/posterior_<shape>.npz, minted on the CPU by
scripts/posterior_reference.py).  Prints one JSON document per shape: across-seed mean and
between-seed sd of every summary for the four samplers, and the engine-minus-reference differences
in units of their standard error.

    python scripts/posterior_study.py c3s c5s [--seeds 4]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from moire_b200.workloads import make_posterior_shape  # noqa: E402
from test_gpu_posterior import engine_runs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("shapes", nargs="+")
ap.add_argument("--seeds", type=int, default=4)
a = ap.parse_args()
KEYS = ("m", "r", "eps_neg", "eps_pos", "p0")


def summarise(per_run):
    """per_run: {key: [seeds][...]} -> {key: (mean over seeds of the run mean, sd over seeds)}."""
    out = {}
    for k, v in per_run.items():
        x = np.array([np.mean(r) for r in v])
        out[k] = dict(mean=float(x.mean()), sd=float(x.std(ddof=1)) if len(x) > 1 else 0.0, n=len(x))
    return out


for shape in a.shapes:
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"posterior_{shape}.npz")))
    wl = make_posterior_shape(shape)
    samplers = {}
    for prec in (32, 64):
        samplers[f"reference_fp{prec}"] = dict({k: list(g[f"ref{prec}.{k}"]) for k in KEYS},
                                               mean_coi=list(g[f"ref{prec}.mean_coi"]),
                                               llik=list(g[f"ref{prec}.llik_sample_mean"]))
    for name, kw in (("engine_fp64", {}), ("engine_ref32_emulation", dict(ref32_emulation=True))):
        runs = engine_runs(wl, seeds=range(11, 11 + a.seeds), **kw)
        samplers[name] = dict({k: [r[k] for r in runs] for k in KEYS}, mean_coi=[r["mean_coi"] for r in runs],
                              llik=[r["llik"] for r in runs])
    doc = dict(shape=shape, samples=wl["packed"].num_samples, loci=wl["packed"].num_loci, burnin=wl["burnin"],
               draws=wl["samples"], summaries={n: summarise(v) for n, v in samplers.items()}, differences={},
               per_sample_coi_correlation={})
    for eng, ref in (("engine_fp64", "reference_fp64"), ("engine_fp64", "reference_fp32"),
                     ("engine_ref32_emulation", "reference_fp32"), ("reference_fp64", "reference_fp32")):
        d = {}
        for k in list(KEYS) + ["mean_coi", "llik"]:
            x, y = doc["summaries"][eng][k], doc["summaries"][ref][k]
            se = np.sqrt(x["sd"] ** 2 / x["n"] + y["sd"] ** 2 / y["n"])
            d[k] = dict(diff=x["mean"] - y["mean"], se=float(se), z=float((x["mean"] - y["mean"]) / se) if se > 0 else None)
        doc["differences"][f"{eng} - {ref}"] = d
        ma, mb = np.mean(samplers[eng]["m"], axis=0), np.mean(samplers[ref]["m"], axis=0)
        doc["per_sample_coi_correlation"][f"{eng} vs {ref}"] = float(np.corrcoef(ma, mb)[0, 1])
    print(json.dumps(doc))
