#!/usr/bin/env python
"""Mints the golden vectors under tests/golden/ from the reference itself.

The reference's own tests hold no vectors for the likelihood path, so the known
answers are generated here by calling the UNMODIFIED reference sources, compiled in place from
/root/reference/src against oracle/shim (``make -C oracle ref`` -> oracle/_ref/libmoire_ref32.so
in the float arithmetic the reference runs, and libmoire_ref64.so, its mechanical float->double
variant).  Run in the build container (the only place /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden.py

Everything is seeded; the outputs are committed so that the GPU box (which has no
/root/reference) can check both oracle/moire_oracle.c and the CUDA engine against them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from moire_b200.data import pack_data  # noqa: E402
from moire_b200.simulate import simulate_data  # noqa: E402
from oracle.pyoracle import Ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def simplexes(rng, k):
    """Fixed + random + near-degenerate simplexes of size k."""
    out = [np.full(k, 1.0 / k)]
    out.append(rng.dirichlet(np.ones(k)))
    out.append(rng.dirichlet(np.full(k, 0.3)))
    if k > 1:
        q = np.full(k, 1e-4)
        q[0] = 1 - 1e-4 * (k - 1)
        out.append(q)
        q = rng.dirichlet(np.ones(k)) * 1e-3
        q[-1] = 1 - q[:-1].sum()
        out.append(q)
    return out


def golden_pure(prec):
    ref = Ref(prec)
    rng = np.random.default_rng(1234)
    g = {}
    # (i) combination sequences, 1 <= r <= n <= 12 (+ the degenerate r = 0 and r = n + 1)
    combo_rows, combo_meta = [], []
    for n in range(1, 13):
        for r in range(0, n + 2):
            seq, nc, gen = ref.combinations(n, r)
            combo_meta.append([n, r, len(seq), nc, gen])
            flat = np.full((len(seq), 12), -1, np.int8)
            flat[:, :r if r <= 12 else 12] = seq[:, :12]
            combo_rows.append(flat)
    g["combo_meta"] = np.array(combo_meta, np.int64)
    g["combo_rows"] = np.concatenate(combo_rows)

    # (ii) PAM scalar & vectorised, k = 1..14
    q_all, k_all, n_all, s_all, v_all = [], [], [], [], []
    for k in range(1, 15):
        for q in simplexes(rng, k):
            if prec == 32:
                q = q.astype(np.float32).astype(np.float64)
            qq = np.zeros(14)
            qq[:k] = q
            vec = ref.pam_vectorized(q, 40)
            for n in sorted(set(list(range(max(1, k - 1), min(k + 6, 41))) + [20, 33, 40])):
                q_all.append(qq)
                k_all.append(k)
                n_all.append(n)
                s_all.append(ref.pam_scalar(q, n))
                v_all.append(vec)
    g["pam_q"] = np.array(q_all)
    g["pam_k"] = np.array(k_all, np.int32)
    g["pam_n"] = np.array(n_all, np.int32)
    g["pam_scalar"] = np.array(s_all)
    g["pam_vectorized"] = np.array(v_all)

    # a tiny dataset so a Chain can be constructed to reach its private term functions
    d = simulate_data(mean_coi=3, num_samples=4, epsilon_pos=.01, epsilon_neg=.1,
                      locus_freq_alphas=[np.ones(5)] * 2, seed=9)
    pk = pack_data(d["data"], d["is_missing"])
    ref.set_data(pk.num_alleles, pk.obs_mask, pk.is_missing)

    # (iii) transmission on a (k, m, r) grid, both relatedness modes
    rows = []
    for allow in (1, 0):
        ref.set_params(allow_relatedness=allow)
        h = ref.chain_new(1.0, seed=3)
        for A in (5, 12, 30):
            for rep in range(4):
                p = rng.dirichlet(np.ones(A) * (0.5 if rep % 2 else 1.0))
                if prec == 32:
                    p = p.astype(np.float32).astype(np.float64)
                for k in sorted(set([1, 2, 3, min(5, A), min(8, A), min(10, A), min(11, A),
                                     min(14, A)])):
                    idx = np.sort(rng.choice(A, size=k, replace=False))
                    for m in sorted(set([k, k + 1, k + 3, min(40, k + 9), 40])):
                        for r in ((0.0,) if not allow else (1e-3, 0.2, 0.77)):
                            val = ref.chain_transmission(h, idx, p, m, r)
                            mask = int(sum(1 << int(i) for i in idx))
                            pp = np.zeros(30)
                            pp[:A] = p
                            rows.append((allow, A, mask, m, r, val, pp))
        ref.chain_free(h)
    g["tr_allow"] = np.array([x[0] for x in rows], np.int32)
    g["tr_A"] = np.array([x[1] for x in rows], np.int32)
    g["tr_mask"] = np.array([x[2] for x in rows], np.uint64)
    g["tr_m"] = np.array([x[3] for x in rows], np.int32)
    g["tr_r"] = np.array([x[4] for x in rows])
    g["tr_val"] = np.array([x[5] for x in rows])
    g["tr_p"] = np.array([x[6] for x in rows])

    # (iv) observation for random mask pairs
    ref.set_params(allow_relatedness=1)
    h = ref.chain_new(1.0, seed=3)
    rows = []
    for A in (2, 5, 10, 35, 64):
        for rep in range(12):
            lat = rng.random(A) < 0.3
            if not lat.any():
                lat[rng.integers(A)] = True
            obs = np.where(rng.random(A) < 0.8, lat, ~lat)
            en, ep = rng.uniform(1e-4, 0.9), rng.uniform(1e-4, 0.9)
            if prec == 32:
                en, ep = float(np.float32(en)), float(np.float32(ep))
            val = ref.chain_observation(h, np.flatnonzero(lat), obs.astype(np.int32), en, ep)
            lm = int(sum(1 << int(i) for i in np.flatnonzero(lat)))
            om = int(sum(1 << int(i) for i in np.flatnonzero(obs)))
            rows.append((A, lm, om, en, ep, val))
    ref.chain_free(h)
    g["ob_A"] = np.array([x[0] for x in rows], np.int32)
    g["ob_latent"] = np.array([x[1] for x in rows], np.uint64)
    g["ob_obs"] = np.array([x[2] for x in rows], np.uint64)
    g["ob_eps_neg"] = np.array([x[3] for x in rows])
    g["ob_eps_pos"] = np.array([x[4] for x in rows])
    g["ob_val"] = np.array([x[5] for x in rows])

    # (v) sample_latent_genotype: the drawn set and its log proposal density
    rows = []
    for A in (2, 5, 10, 35):
        for rep in range(40):
            obs = (rng.random(A) < rng.choice([0.0, 0.15, 0.4, 0.9])).astype(np.int32)
            coi = int(rng.integers(1, 13))
            ep, en = rng.uniform(.005, .6), rng.uniform(.005, .6)
            if prec == 32:
                en, ep = float(np.float32(en)), float(np.float32(ep))
            idx, lp = ref.sample_latent_genotype(1000 + rep, obs, coi, ep, en)
            lm = int(sum(1 << int(i) for i in idx))
            om = int(sum(1 << int(i) for i in np.flatnonzero(obs)))
            rows.append((A, om, coi, ep, en, lm, lp))
    g["lg_A"] = np.array([x[0] for x in rows], np.int32)
    g["lg_obs"] = np.array([x[1] for x in rows], np.uint64)
    g["lg_coi"] = np.array([x[2] for x in rows], np.int32)
    g["lg_eps_pos"] = np.array([x[3] for x in rows])
    g["lg_eps_neg"] = np.array([x[4] for x in rows])
    g["lg_latent"] = np.array([x[5] for x in rows], np.uint64)
    g["lg_log_prob"] = np.array([x[6] for x in rows])

    # (vi) sample_constrained as a function of its normal draw
    rows = []
    ms = ref.min_sampled()
    for s in range(60):
        for curr, sd, lo, hi in ((0.03, 1.0, ms, 1.0), (0.6, 0.3, ms, .99), (3.7, 0.8, 1.0, 40.0),
                                 (5.2, 1.0, 1e-5, 40.0)):
            if prec == 32:
                curr, lo = float(np.float32(curr)), float(np.float32(lo))
                hi = float(np.float32(hi))
            z, prop, adj = ref.sample_constrained(s, curr, sd, lo, hi)
            rows.append((z, curr, lo, hi, prop, adj))
    g["sc"] = np.array(rows)

    # (vii) SALT proposal
    rows = []
    for A in (2, 3, 5, 10, 30):
        for rep in range(25):
            p = rng.dirichlet(np.ones(A) * (0.2 if rep % 3 == 0 else 1.0))
            p = np.maximum(p, 1e-9)
            p /= p.sum()
            if prec == 32:
                p = p.astype(np.float32).astype(np.float64)
            idx = int(rng.integers(A))
            lprop = ref.logit(p[idx]) + rng.normal() * (3.0 if rep % 5 == 0 else 1.0)
            if prec == 32:
                lprop = float(np.float32(lprop))
            out, adj = ref.salt(p, idx, lprop)
            pp, oo = np.zeros(30), np.zeros(30)
            pp[:A], oo[:A] = p, out
            rows.append((A, idx, lprop, adj, pp, oo))
    g["salt_A"] = np.array([x[0] for x in rows], np.int32)
    g["salt_idx"] = np.array([x[1] for x in rows], np.int32)
    g["salt_logit_prop"] = np.array([x[2] for x in rows])
    g["salt_log_adj"] = np.array([x[3] for x in rows])
    g["salt_p"] = np.array([x[4] for x in rows])
    g["salt_prop_p"] = np.array([x[5] for x in rows])

    # (viii) priors
    rows = []
    for x in np.concatenate([rng.uniform(1e-6, 1 - 1e-6, 20), [0.0, 1.0]]):
        if prec == 32:
            x = float(np.float32(x))
        rows.append((x, ref.beta_log_prior(x, 1, 1), ref.beta_log_prior(x, .1, 9.9),
                     ref.beta_log_prior(x, 2.5, 1.0), ref.gamma_log_prior(x * 20, .1, 10),
                     ref.gamma_log_prior(x * 20, 2.0, 1.5)))
    g["prior_beta_gamma"] = np.array(rows)
    rows = []
    for lam in (0.3, 1.0, 4.2, 17.5):
        for x in (1, 2, 5, 17, 40, 100):
            rows.append((x, lam, ref.dztpois(x, lam)))
    g["prior_dztpois"] = np.array(rows)
    rows = []
    for size in (0, 1, 4, 9, 39):
        for x in sorted(set([0, size // 2, size])):
            for pr in (1e-3, .2, .77):
                if prec == 32:
                    pr = float(np.float32(pr))
                rows.append((x, size, pr, ref.dbinom(x, size, pr)))
    g["prior_dbinom"] = np.array(rows)
    return g


STATE_CONFIGS = {
    # name: (simulate_data kwargs, params, steps)
    "c1like_rel": (dict(mean_coi=4, num_samples=24, epsilon_pos=.03, epsilon_neg=.1,
                        locus_freq_alphas=[np.ones(5)] * 5 + [np.ones(10)] * 5,
                        internal_relatedness_alpha=.1, internal_relatedness_beta=1,
                        missingness=.05, seed=11), dict(allow_relatedness=1), 6),
    "c1like_norel": (dict(mean_coi=4, num_samples=24, epsilon_pos=.03, epsilon_neg=.1,
                          locus_freq_alphas=[np.ones(5)] * 5 + [np.ones(10)] * 5,
                          missingness=.05, seed=12), dict(allow_relatedness=0), 6),
    "highcoi_rel": (dict(sample_cois=np.arange(1, 21), epsilon_pos=.02, epsilon_neg=.05,
                         locus_freq_alphas=[np.ones(30)] * 6 + [np.ones(12)] * 2,
                         internal_relatedness_alpha=1, internal_relatedness_beta=1,
                         missingness=.02, seed=13), dict(allow_relatedness=1), 4),
    "highcoi_norel": (dict(sample_cois=np.arange(1, 21), epsilon_pos=.02, epsilon_neg=.05,
                           locus_freq_alphas=[np.ones(30)] * 6 + [np.ones(12)] * 2,
                           missingness=.02, seed=14), dict(allow_relatedness=0), 4),
}


def golden_states(prec):
    """(ix) whole chain state after k reference steps, with every cached likelihood term."""
    ref = Ref(prec)
    g = {}
    for name, (simkw, prm, steps) in STATE_CONFIGS.items():
        d = simulate_data(**simkw)
        pk = pack_data(d["data"], d["is_missing"])
        ref.set_data(pk.num_alleles, pk.obs_mask, pk.is_missing)
        ref.set_params(**prm)
        h = ref.chain_new(1.0, seed=21)
        for it in range(steps):
            ref.chain_step(h, it)
        st = ref.chain_get_state(h)
        ref.chain_free(h)
        g[f"{name}.num_alleles"] = pk.num_alleles
        g[f"{name}.obs_mask"] = pk.obs_mask
        g[f"{name}.is_missing"] = pk.is_missing
        g[f"{name}.allow_relatedness"] = np.array(prm["allow_relatedness"])
        for k, v in st.items():
            g[f"{name}.{k}"] = np.asarray(v)
    return g


def main():
    for prec in (32, 64):
        np.savez_compressed(os.path.join(OUT, f"pure_ref{prec}.npz"), **golden_pure(prec))
        np.savez_compressed(os.path.join(OUT, f"states_ref{prec}.npz"), **golden_states(prec))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
