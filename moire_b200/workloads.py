"""The named benchmark shapes of BASELINE.json / SURVEY.md section 8(d), generated with the
restated ``simulate_data``.  Seeds: Philox(17325 + config number)."""
from __future__ import annotations

import numpy as np

from .data import pack_data
from .simulate import simulate_data, simulate_sample_coi


def ladder(T, grad=1.0):
    """R/mcmc.R:139-141: seq(1, 0, length.out = T) ** pt_grad."""
    if T == 1:
        return np.array([1.0])
    return np.linspace(1.0, 0.0, T) ** grad


def make_workload(name: str):
    name = name.lower()
    model = dict(allow_relatedness=True, max_coi=40)
    if name == "c1":     # packaged simulated_data shape (data-raw/mcmc_results.R:4-24), T = 1
        sim = dict(mean_coi=5, num_samples=100, epsilon_pos=.01, epsilon_neg=.1,
                   locus_freq_alphas=[np.ones(5)] * 15 + [np.ones(10)] * 15,
                   internal_relatedness_alpha=.1, internal_relatedness_beta=1, seed=17325 + 1)
        T = 1
    elif name == "c2":   # namibia_data shape: 2585 x 26, A_j 7..35 (sum 451), 11.4% missing, T = 10
        rng = np.random.Generator(np.random.Philox(99))
        A = np.array([7, 8, 9, 10, 10, 11, 12, 12, 13, 14, 15, 15, 16, 17, 18, 18, 19, 20, 21, 22,
                      23, 24, 26, 28, 28, 35])
        assert A.sum() == 451 and len(A) == 26
        sim = dict(mean_coi=1.6, num_samples=2585, epsilon_pos=.01, epsilon_neg=.1,
                   locus_freq_alphas=[np.full(a, .5) for a in A],
                   internal_relatedness_alpha=.1, internal_relatedness_beta=1,
                   missingness=.114, seed=17325 + 2)
        T = 10
    elif name == "c3":   # the north-star shape: 1000 x 100, A = 10, COI <= 10, T = 40
        rng = np.random.Generator(np.random.Philox(17325 + 3))
        coi = np.minimum(simulate_sample_coi(1000, 4.0, rng), 10)
        sim = dict(sample_cois=coi, epsilon_pos=.01, epsilon_neg=.1,
                   locus_freq_alphas=[np.ones(10)] * 100,
                   internal_relatedness_alpha=.1, internal_relatedness_beta=1, seed=17325 + 3)
        T = 40
    elif name == "c4":   # 5000 x 250, A_j ~ U{5..30}, T = 64 (8 GPUs)
        rng = np.random.Generator(np.random.Philox(17325 + 4))
        A = rng.integers(5, 31, size=250)
        coi = np.minimum(simulate_sample_coi(5000, 4.0, rng), 10)
        sim = dict(sample_cois=coi, epsilon_pos=.01, epsilon_neg=.1,
                   locus_freq_alphas=[np.ones(a) for a in A],
                   internal_relatedness_alpha=.1, internal_relatedness_beta=1, seed=17325 + 4)
        T = 64
    elif name == "c5":   # high-COI stress: 2000 x 50, A = 30, COI ~ U{1..30}, relatedness on
        rng = np.random.Generator(np.random.Philox(17325 + 5))
        coi = rng.integers(1, 31, size=2000)
        sim = dict(sample_cois=coi, epsilon_pos=.01, epsilon_neg=.1,
                   locus_freq_alphas=[np.ones(30)] * 50,
                   internal_relatedness_alpha=1, internal_relatedness_beta=1, seed=17325 + 5)
        T = 32
    elif name == "tiny":
        sim = dict(mean_coi=3, num_samples=48, epsilon_pos=.02, epsilon_neg=.1,
                   locus_freq_alphas=[np.ones(6)] * 12, internal_relatedness_alpha=.1,
                   internal_relatedness_beta=1, seed=17325)
        T = 4
    else:
        raise ValueError(f"unknown workload {name!r}")
    d = simulate_data(**sim)
    pk = pack_data(d["data"], d["is_missing"])
    sum_alleles = int(pk.num_alleles.sum())
    nonmissing = float(1.0 - pk.is_missing.mean())
    # nominal evaluations per chain-step: 6 N L + N sum_j A_j with relatedness
    nominal = (6 * pk.num_samples * pk.num_loci + pk.num_samples * sum_alleles) * nonmissing
    return dict(name=name, sim=d, packed=pk, temps=ladder(T), model=model,
                nominal_evals_per_chain_step=nominal)


def make_posterior_shape(name: str):
    """The hard shapes of BASELINE.json scaled to what the reference's CPU sampler finishes in
    minutes (scripts/posterior_reference.py mints the reference side, tests/test_gpu_posterior.py
    and scripts/posterior_study.py run the engine on the same data): ``c3s`` is C3-shaped (A = 10,
    COI <= 10 from a zero-truncated Poisson(4), sparse relatedness) at 500 samples x 40 loci;
    ``c5s`` is C5-shaped (A = 30, COI uniform on 1..30, relatedness Beta(1, 1): cells with more than
    ten latent alleles, coverage probabilities far below the fp32 reference's 2^-23 floor) at
    150 samples x 20 loci.  One chain (T = 1), ``max_coi`` 40."""
    name = name.lower()
    if name == "c3s":
        rng = np.random.Generator(np.random.Philox(17325 + 33))
        coi = np.minimum(simulate_sample_coi(500, 4.0, rng), 10)
        sim = dict(sample_cois=coi, epsilon_pos=.01, epsilon_neg=.1,
                   locus_freq_alphas=[np.ones(10)] * 40,
                   internal_relatedness_alpha=.1, internal_relatedness_beta=1, seed=17325 + 33)
        burnin, samples = 1000, 1000
    elif name == "c5s":
        rng = np.random.Generator(np.random.Philox(17325 + 55))
        coi = rng.integers(1, 31, size=150)
        sim = dict(sample_cois=coi, epsilon_pos=.01, epsilon_neg=.1,
                   locus_freq_alphas=[np.ones(30)] * 20,
                   internal_relatedness_alpha=1, internal_relatedness_beta=1, seed=17325 + 55)
        burnin, samples = 700, 700
    else:
        raise ValueError(f"unknown posterior shape {name!r}")
    d = simulate_data(**sim)
    pk = pack_data(d["data"], d["is_missing"])
    return dict(name=name, sim=d, packed=pk, temps=np.array([1.0]), burnin=burnin, samples=samples,
                model=dict(allow_relatedness=True, max_coi=40))
