"""Synthetic genotyping data: a NumPy restatement of moire's ``simulate_data``.

Follows the generative model of the reference's R simulator (R/simulator.R:154-212):
allele frequencies ~ Dirichlet(alpha_j) per locus (:165-175), COI ~ zero-truncated
Poisson(mean_coi) by inverse CDF (:44-46), within-host relatedness ~ Beta(a, b) (:181-183; the
default a = 0 is R's point mass at 0), per locus ``related ~ Binom(coi - 1, r)`` and allele
counts ~ Multinom(coi - related, freqs) (:63-69), then the observation process: a present allele
is seen w.p. ``1 - eps_neg / A_j``, an absent one w.p. ``eps_pos / A_j``, and the whole cell is
blanked w.p. ``missingness`` (:85-107).  ``is_missing`` = no allele observed (:195-199).

R's RNG stream cannot be reproduced without R; draws come from ``numpy.random.Generator(Philox)``.
This is the benchmark / test input generator, not part of the likelihood path.
"""
from __future__ import annotations

import numpy as np


def _ztpois_icdf(u, lam):
    """qpois(u, lam) for u already restricted to (dpois(0), 1): smallest k with CDF(k) >= u."""
    out = np.zeros(u.shape, dtype=np.int64)
    pmf = np.exp(-lam)
    cdf = np.full(u.shape, pmf)
    k = 0
    todo = cdf < u
    while todo.any() and k < 10000:
        k += 1
        pmf = pmf * lam / k
        cdf = cdf + pmf
        out[todo] = k
        todo = todo & (cdf < u)
    return np.maximum(out, 1)


def simulate_sample_coi(num_samples, mean_coi, rng):
    p0 = np.exp(-mean_coi)
    u = rng.uniform(p0, 1.0, size=num_samples)
    return _ztpois_icdf(u, float(mean_coi))


def simulate_data(mean_coi=None, num_samples=None, epsilon_pos=0.0, epsilon_neg=0.0,
                  sample_cois=None, locus_freq_alphas=None, allele_freqs=None,
                  internal_relatedness_alpha=0.0, internal_relatedness_beta=1.0,
                  internal_relatedness=None, missingness=0.0, seed=0):
    """Returns a dict with the fields of the reference's ``simulate_data`` result that the
    sampler consumes (``data``, ``sample_ids``, ``loci``, ``is_missing``) plus the simulated
    truth (``allele_freqs``, ``sample_cois``, ``sample_relatedness``, ``true_genotypes``).

    ``data[j]`` is an int8 array [num_samples][A_j] of 0/1 (the reference's
    ``data[[locus]][[sample]]`` vectors stacked); ``is_missing`` is bool [L][N]."""
    rng = np.random.Generator(np.random.Philox(seed))
    if allele_freqs is None:
        allele_freqs = [rng.dirichlet(np.asarray(a, dtype=np.float64)) for a in locus_freq_alphas]
    else:
        allele_freqs = [np.asarray(f, dtype=np.float64) for f in allele_freqs]
    L = len(allele_freqs)
    if sample_cois is None:
        sample_cois = simulate_sample_coi(num_samples, mean_coi, rng)
    sample_cois = np.asarray(sample_cois, dtype=np.int64)
    N = len(sample_cois)
    if internal_relatedness is None:
        if internal_relatedness_alpha <= 0:
            internal_relatedness = np.zeros(N)
        else:
            internal_relatedness = rng.beta(internal_relatedness_alpha, internal_relatedness_beta,
                                            size=N)
    internal_relatedness = np.asarray(internal_relatedness, dtype=np.float64)

    true_genotypes, data = [], []
    is_missing = np.zeros((L, N), dtype=bool)
    for j, freqs in enumerate(allele_freqs):
        A = len(freqs)
        related = rng.binomial(sample_cois - 1, internal_relatedness)
        counts = rng.multinomial(sample_cois - related, freqs)          # [N][A]
        present = counts > 0
        prob = np.where(present, 1.0 - epsilon_neg / A, epsilon_pos / A)
        obs = (rng.uniform(size=(N, A)) < prob)
        blank = rng.uniform(size=N) < missingness
        obs[blank, :] = False
        is_missing[j] = ~obs.any(axis=1)
        true_genotypes.append(counts)
        data.append(obs.astype(np.int8))
    return dict(data=data, sample_ids=[f"S{i + 1}" for i in range(N)],
                loci=[f"L{j + 1}" for j in range(L)], is_missing=is_missing,
                allele_freqs=allele_freqs, sample_cois=sample_cois,
                sample_relatedness=internal_relatedness, true_genotypes=true_genotypes)


def simulate_data_gpu(mean_coi=None, num_samples=None, epsilon_pos=0.0, epsilon_neg=0.0,
                      sample_cois=None, locus_freq_alphas=None, allele_freqs=None,
                      internal_relatedness_alpha=0.0, internal_relatedness_beta=1.0,
                      internal_relatedness=None, missingness=0.0, seed=0, device=0):
    """``simulate_data`` with the per-cell work (N x L x A_j: strains, alleles, observation noise,
    missingness) on the GPU (``moire_simulate_data``); the per-locus frequencies and per-sample
    COIs / relatedness -- a few thousand numbers -- are drawn here exactly as ``simulate_data``
    draws them.  Returns the packed masks the engine consumes plus the truth:
    ``dict(packed=PackedData, true_mask, allele_freqs, sample_cois, sample_relatedness)``.
    Same model, different random streams: the two generators agree in distribution."""
    import ctypes as C

    from . import _lib
    from .data import PackedData
    rng = np.random.Generator(np.random.Philox(seed))
    if allele_freqs is None:
        allele_freqs = [rng.dirichlet(np.asarray(a, dtype=np.float64)) for a in locus_freq_alphas]
    else:
        allele_freqs = [np.asarray(f, dtype=np.float64) for f in allele_freqs]
    L = len(allele_freqs)
    if sample_cois is None:
        sample_cois = simulate_sample_coi(num_samples, mean_coi, rng)
    sample_cois = np.ascontiguousarray(sample_cois, dtype=np.int32)
    N = len(sample_cois)
    if internal_relatedness is None:
        internal_relatedness = (np.zeros(N) if internal_relatedness_alpha <= 0 else
                                rng.beta(internal_relatedness_alpha, internal_relatedness_beta, size=N))
    rel = np.ascontiguousarray(internal_relatedness, dtype=np.float64)
    na = np.array([len(f) for f in allele_freqs], dtype=np.int32)
    AP = int(na.max())
    fr = np.zeros((L, AP))
    for j, f in enumerate(allele_freqs):
        fr[j, :len(f)] = f
    W = _lib.words_for(AP)  # one mask word, or two past 64 alleles per locus
    shape = (N, L) if W == 1 else (N, L, W)
    obs, tru = np.zeros(shape, np.uint64), np.zeros(shape, np.uint64)
    mis = np.zeros((N, L), np.uint8)
    lib = _lib.load(W)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.moire_simulate_data(N, L, ptr(na), ptr(fr), AP, ptr(sample_cois), ptr(rel), float(epsilon_pos),
                                 float(epsilon_neg), float(missingness), int(seed), int(device), ptr(obs),
                                 ptr(mis), ptr(tru))
    if rc != 0:
        raise RuntimeError(f"moire_simulate_data failed ({rc}): {lib.moire_last_error().decode()}")
    bits = np.vectorize(lambda x: bin(int(x)).count("1"))(obs)
    coi = (bits if W == 1 else bits.sum(axis=2)).max(axis=1).astype(np.int32)
    return dict(packed=PackedData(N, L, na, obs, mis, coi), true_mask=tru, allele_freqs=allele_freqs,
                sample_cois=sample_cois, sample_relatedness=rel)
