#!/usr/bin/env python
"""Mints tests/golden/packaged_data.npz from the reference's packaged data files
(/root/reference/data/*.rda), decoded with moire_b200.loaders.read_rda:

* simulated_data (100 samples x 30 loci; the input of the reference's demo run) as bit masks,
* namibia_data (long form, 97 214 rows) through load_long_form_data, as bit masks,
* posterior summaries of mcmc_results (the reference authors' own run on simulated_data with
  80 temperatures, data-raw/mcmc_results.R) and the arguments that produced it.

    python tests/golden/make_packaged_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from moire_b200.data import pack_data  # noqa: E402
from moire_b200.loaders import load_long_form_data, read_rda  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.dirname(os.path.abspath(__file__))

sd = read_rda(os.path.join(REF, "simulated_data.rda"))["simulated_data"]
pk = pack_data(sd["data"], sd["is_missing"] != 0)
g = {"sim.obs_mask": pk.obs_mask, "sim.is_missing": pk.is_missing, "sim.num_alleles": pk.num_alleles,
     "sim.true_coi": np.asarray(sd["sample_cois"]).astype(np.int32),
     "sim.true_relatedness": np.asarray(sd["sample_relatedness"], dtype=np.float64)}
nd = read_rda(os.path.join(REF, "namibia_data.rda"))["namibia_data"]
nam = load_long_form_data(dict(sample_id=nd["sample_id"], locus=nd["locus"], allele=nd["allele"]))
npk = pack_data(nam["data"], nam["is_missing"])
g.update({"nam.obs_mask": npk.obs_mask, "nam.is_missing": npk.is_missing,
          "nam.num_alleles": npk.num_alleles, "nam.loci": np.array(nam["loci"]),
          "nam.first_samples": np.array(nam["sample_ids"][:5])})
mr = read_rda(os.path.join(REF, "mcmc_results.rda"))["mcmc_results"]
ch, args = mr["chains"][0], mr["args"]
for k in ("eps_pos_alpha", "eps_pos_beta", "eps_neg_alpha", "eps_neg_beta", "r_alpha", "r_beta",
          "mean_coi_shape", "mean_coi_scale", "max_eps_pos", "max_eps_neg", "max_coi", "burnin",
          "samples_per_chain", "pre_adapt_steps", "temp_adapt_steps"):
    g["res.arg." + k] = np.asarray(args[k], dtype=np.float64).reshape(-1)[0]
g["res.pt_chains"] = np.asarray(args["pt_chains"], dtype=np.float64)
g["res.llik_sample"] = np.asarray(ch["llik_sample"], dtype=np.float64)
g["res.lam_coi"] = np.asarray(ch["lam_coi"], dtype=np.float64)
g["res.mean_coi_per_sample"] = np.array([np.mean(c) for c in ch["coi"]])
g["res.mean_eps_neg"] = np.array([np.mean(c) for c in ch["eps_neg"]])
g["res.mean_eps_pos"] = np.array([np.mean(c) for c in ch["eps_pos"]])
g["res.mean_relatedness"] = np.array([np.mean(c) for c in ch["relatedness"]])
g["res.mean_first_allele_freq"] = np.array([np.mean([d[0] for d in loc]) for loc in ch["allele_freqs"]])
g["res.temp_gradient"] = np.asarray(ch["temp_gradient"], dtype=np.float64)
np.savez_compressed(os.path.join(OUT, "packaged_data.npz"), **g)
print(os.path.getsize(os.path.join(OUT, "packaged_data.npz")))
