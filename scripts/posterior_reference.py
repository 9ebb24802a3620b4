#!/usr/bin/env python
"""Posterior summaries of the REFERENCE's own sampler (oracle/_ref: the unmodified C++, fp32 as
shipped and its mechanical fp64 variant) on the hard shapes of BASELINE.json, scaled to what the
CPU finishes in minutes -- the fixtures behind tests/test_gpu_posterior.py and
scripts/posterior_study.py.  Runs the seeds side by side (one process each) and writes
tests/golden/posterior_<shape>.npz.  Needs /root/reference (through oracle/_ref); the fixture it
writes travels to the GPU box, the reference does not have to.

    python scripts/posterior_reference.py c3s|c5s [--seeds32 4] [--seeds64 2]"""
import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moire_b200.workloads import make_posterior_shape  # noqa: E402


def one_run(args):
    shape, prec, seed, burnin, samples = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle.pyoracle import Ref
    wl = make_posterior_shape(shape)
    pk = wl["packed"]
    ref = Ref(prec)
    ref.set_data(pk.num_alleles, pk.obs_mask, pk.is_missing)
    ref.set_params(pt_chains=[1.0], burnin=burnin, samples=samples, max_coi=wl["model"]["max_coi"])
    h = ref.mcmc_new(seed=seed)
    t0 = time.time()
    for i in range(burnin):
        ref.lib.ref_mcmc_burnin(h, i)
    for i in range(samples):
        ref.lib.ref_mcmc_sample(h, i)
    sm = ref.mcmc_summary(h)
    llb = np.zeros(burnin)
    lls = np.zeros(samples)
    ref.lib.ref_mcmc_trace(h, llb.ctypes.data, burnin, lls.ctypes.data, samples)
    ref.mcmc_free(h)
    return dict(prec=prec, seed=seed, m=sm["m"], r=sm["r"], eps_neg=sm["eps_neg"], eps_pos=sm["eps_pos"],
                p0=sm["p"][:, 0], mean_coi=sm["mean_coi"], llik_sample_mean=float(lls.mean()),
                llik_sample_sd=float(lls.std()), seconds=time.time() - t0)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("shape")
    ap.add_argument("--seeds32", type=int, default=4)
    ap.add_argument("--seeds64", type=int, default=2)
    a = ap.parse_args()
    wl = make_posterior_shape(a.shape)
    jobs = [(a.shape, 32, 100 + s, wl["burnin"], wl["samples"]) for s in range(a.seeds32)] + \
           [(a.shape, 64, 300 + s, wl["burnin"], wl["samples"]) for s in range(a.seeds64)]
    with mp.Pool(len(jobs)) as pool:
        runs = pool.map(one_run, jobs)
    out = {"shape": a.shape, "burnin": wl["burnin"], "samples": wl["samples"]}
    for prec in (32, 64):
        rs = [r for r in runs if r["prec"] == prec]
        if not rs:
            continue
        for k in ("m", "r", "eps_neg", "eps_pos", "p0"):
            out[f"ref{prec}.{k}"] = np.array([r[k] for r in rs])          # [seeds][N or L]
        for k in ("mean_coi", "llik_sample_mean", "llik_sample_sd", "seconds", "seed"):
            out[f"ref{prec}.{k}"] = np.array([r[k] for r in rs], dtype=np.float64)
    path = os.path.join(ROOT, "tests", "golden", f"posterior_{a.shape}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
