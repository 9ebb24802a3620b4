#!/usr/bin/env python
"""CPU only: the reference's sampler in its two precisions (oracle/_ref: the unmodified fp32 build
and the mechanical float->double variant) on the C1-shaped data of scripts/posterior_study.py, one
process per seed.  Tells whether a posterior difference between the engine (fp64) and the fp32
reference comes from the precision.  usage: posterior_ref64.py [seeds] [burnin] [samples]"""
import json
import os
import sys
from multiprocessing import Pool

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(job):
    prec, seed, burnin, samples = job
    from moire_b200.workloads import make_workload
    from oracle.pyoracle import Ref
    pk = make_workload("c1")["packed"]
    ref = Ref(prec)
    ref.set_data(pk.num_alleles, pk.obs_mask, pk.is_missing)
    ref.set_params(pt_chains=[1.0], burnin=burnin, samples=samples)
    h = ref.mcmc_new(seed=seed)
    for i in range(burnin):
        ref.lib.ref_mcmc_burnin(h, i)
    for i in range(samples):
        ref.lib.ref_mcmc_sample(h, i)
    sm = ref.mcmc_summary(h)
    ref.mcmc_free(h)
    return prec, dict(m=float(np.mean(sm["m"])), r=float(np.mean(sm["r"])), eps_neg=float(np.mean(sm["eps_neg"])),
                      eps_pos=float(np.mean(sm["eps_pos"])), p0=float(np.mean(sm["p"][:, 0])),
                      lam=float(sm["mean_coi"]))


if __name__ == "__main__":
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    burnin = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    samples = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
    jobs = [(p, 100 + s, burnin, samples) for s in range(seeds) for p in (64, 32)]
    with Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        res = pool.map(one, jobs)
    out = {}
    for prec in (32, 64):
        runs = [r for p, r in res if p == prec]
        out[f"ref{prec}"] = {k: dict(mean=float(np.mean([r[k] for r in runs])),
                                     sd=float(np.std([r[k] for r in runs], ddof=1))) for k in runs[0]}
    print(json.dumps(dict(seeds=seeds, burnin=burnin, samples=samples, **out), indent=1))
