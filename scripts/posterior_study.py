#!/usr/bin/env python
"""Posterior equivalence across seeds (BASELINE.json's second correctness criterion): the engine's
`run_mcmc` against the reference's own C++ sampler (oracle/_ref, fp32, unmodified) on C1-shaped
synthetic data (100 samples x 30 loci, data-raw/mcmc_results.R:4-24 recipe), several seeds each.

Prints one JSON document: per-statistic mean over seeds, between-seed sd within each sampler, and
the difference between samplers in units of the pooled standard error."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moire_b200.mcmc import run_mcmc  # noqa: E402
from moire_b200.workloads import make_workload  # noqa: E402
from oracle.pyoracle import Ref  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--burnin", type=int, default=1500)
ap.add_argument("--samples", type=int, default=1500)
ap.add_argument("--seeds", type=int, default=4)
ap.add_argument("--config", default="c1")
a = ap.parse_args()

wl = make_workload(a.config)
pk, sim = wl["packed"], wl["sim"]
N, L = pk.num_samples, pk.num_loci


def stats_from(m, r, en, ep, p0, mean_coi):
    return dict(mean_coi_per_sample=float(np.mean(m)), frac_coi_gt1=float(np.mean(np.asarray(m) > 1.5)),
                mean_relatedness=float(np.mean(r)), mean_eps_neg=float(np.mean(en)),
                mean_eps_pos=float(np.mean(ep)), mean_first_allele_freq=float(np.mean(p0)),
                lam_coi=float(mean_coi), per_sample_m=np.asarray(m), per_locus_p0=np.asarray(p0))


ref = Ref(32)
ref.set_data(pk.num_alleles, pk.obs_mask, pk.is_missing)
ref.set_params(pt_chains=[1.0], burnin=a.burnin, samples=a.samples)
ref_runs, our_runs = [], []
for s in range(a.seeds):
    h = ref.mcmc_new(seed=100 + s)
    for i in range(a.burnin):
        ref.lib.ref_mcmc_burnin(h, i)
    for i in range(a.samples):
        ref.lib.ref_mcmc_sample(h, i)
    sm = ref.mcmc_summary(h)
    ref.mcmc_free(h)
    ref_runs.append(stats_from(sm["m"], sm["r"], sm["eps_neg"], sm["eps_pos"], sm["p"][:, 0], sm["mean_coi"]))
    ch = run_mcmc(sim, sim["is_missing"], burnin=a.burnin, samples_per_chain=a.samples, verbose=False,
                  seed=200 + s)["chains"][0]
    our_runs.append(stats_from(np.mean(ch["coi"], axis=1), np.mean(ch["relatedness"], axis=1),
                               np.mean(ch["eps_neg"], axis=1), np.mean(ch["eps_pos"], axis=1),
                               [np.mean(ch["allele_freqs"][j], axis=0)[0] for j in range(L)],
                               np.mean(ch["lam_coi"])))

out = dict(config=a.config, samples=N, loci=L, burnin=a.burnin, draws=a.samples, seeds=a.seeds, stats={})
for k in ref_runs[0]:
    if k.startswith("per_"):
        continue
    x = np.array([r[k] for r in ref_runs])
    y = np.array([r[k] for r in our_runs])
    se = np.sqrt(x.var(ddof=1) / len(x) + y.var(ddof=1) / len(y))
    out["stats"][k] = dict(reference=float(x.mean()), engine=float(y.mean()), sd_reference=float(x.std(ddof=1)),
                           sd_engine=float(y.std(ddof=1)), diff_in_se=float((y.mean() - x.mean()) / max(se, 1e-300)))
rm = np.mean([r["per_sample_m"] for r in ref_runs], axis=0)
om = np.mean([r["per_sample_m"] for r in our_runs], axis=0)
rp = np.mean([r["per_locus_p0"] for r in ref_runs], axis=0)
op = np.mean([r["per_locus_p0"] for r in our_runs], axis=0)
truth = np.asarray(sim["sample_cois"])
out["per_sample_coi"] = dict(corr_engine_vs_reference=float(np.corrcoef(rm, om)[0, 1]),
                             mean_abs_diff=float(np.abs(rm - om).mean()),
                             between_seed_mean_abs_diff_reference=float(np.mean(
                                 [np.abs(ref_runs[i]["per_sample_m"] - ref_runs[j]["per_sample_m"]).mean()
                                  for i in range(a.seeds) for j in range(i)])),
                             mae_vs_truth_reference=float(np.abs(rm - truth).mean()),
                             mae_vs_truth_engine=float(np.abs(om - truth).mean()))
out["per_locus_first_allele_freq"] = dict(max_abs_diff=float(np.abs(rp - op).max()),
                                          corr=float(np.corrcoef(rp, op)[0, 1]))
print(json.dumps(out, indent=1))
