"""The CPU restatement (oracle/moire_oracle.c) against the golden vectors minted from the
reference itself (tests/golden/make_golden.py).  Integer work and pure +,-,* arithmetic must be
bit-exact; terms that go through libm get a few ulps."""
import numpy as np
import pytest

from conftest import mask_to_idx

TOL = {32: 2e-6, 64: 1e-13}


def close(a, b, prec):
    a, b = np.asarray(a, float), np.asarray(b, float)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    both_nan = np.isnan(a) & np.isnan(b)
    ok = same_inf | both_nan | (np.abs(a - b) <= TOL[prec] * np.maximum(1.0, np.abs(b)))
    return bool(np.all(ok))


def test_combination_sequences_bit_exact(oracle, golden_pure):
    g = golden_pure[64]
    row = 0
    for n, r, nrows, nc, gen in g["combo_meta"]:
        seq, onc, ogen = oracle.combinations(int(n), int(r))
        assert len(seq) == nrows and onc == nc and ogen == gen, (n, r)
        assert np.array_equal(seq[:, :12], g["combo_rows"][row:row + nrows, :min(r, 12)]), (n, r)
        row += nrows
    assert row == len(g["combo_rows"])


@pytest.mark.parametrize("prec", [32, 64])
def test_pam_bit_exact(oracle, golden_pure, prec):
    g = golden_pure[prec]
    for q, k, n, s, vec in zip(g["pam_q"], g["pam_k"], g["pam_n"], g["pam_scalar"],
                               g["pam_vectorized"]):
        assert oracle.pam_scalar(prec, q[:k], int(n)) == s, (k, n)
        assert np.array_equal(oracle.pam_vectorized(prec, q[:k], 40), vec), (k, n)


@pytest.mark.parametrize("prec", [32, 64])
def test_transmission(oracle, golden_pure, prec):
    g = golden_pure[prec]
    for allow, A, mask, m, r, val, p in zip(g["tr_allow"], g["tr_A"], g["tr_mask"], g["tr_m"],
                                            g["tr_r"], g["tr_val"], g["tr_p"]):
        got = oracle.transmission(prec, mask_to_idx(mask), p[:A], int(m), float(r), int(allow))
        assert close(got, val, prec), (allow, A, mask, m, r, got, val)


@pytest.mark.parametrize("prec", [32, 64])
def test_observation(oracle, golden_pure, prec):
    g = golden_pure[prec]
    for A, lat, obs, en, ep, val in zip(g["ob_A"], g["ob_latent"], g["ob_obs"], g["ob_eps_neg"],
                                        g["ob_eps_pos"], g["ob_val"]):
        ov = [(int(obs) >> a) & 1 for a in range(A)]
        got = oracle.observation(prec, mask_to_idx(lat), ov, float(en), float(ep))
        assert close(got, val, prec)


@pytest.mark.parametrize("prec", [32, 64])
def test_latent_genotype_log_prob(oracle, golden_pure, prec):
    """log_prob of sample_latent_genotype as a pure function of the counts it drew."""
    g = golden_pure[prec]
    for A, obs, coi, ep, en, lat, lp in zip(g["lg_A"], g["lg_obs"], g["lg_coi"], g["lg_eps_pos"],
                                            g["lg_eps_neg"], g["lg_latent"], g["lg_log_prob"]):
        obs, lat = int(obs), int(lat)
        obs_pos = bin(obs).count("1")
        fn = bin(lat & ~obs).count("1")
        fp = bin(obs & ~lat).count("1")
        k = bin(lat).count("1")
        assert 1 <= k <= coi
        got = oracle.latent_log_prob(prec, int(A), obs_pos, fn, fp, int(coi), float(ep), float(en))
        assert close(got, lp, prec), (A, obs, coi, lat)


@pytest.mark.parametrize("prec", [32, 64])
def test_constrained_proposal(oracle, golden_pure, prec):
    for z, curr, lo, hi, prop, adj in golden_pure[prec]["sc"]:
        gp, ga = oracle.constrained(prec, z, curr, lo, hi)
        assert close(gp, prop, prec) and close(ga, adj, prec)


@pytest.mark.parametrize("prec", [32, 64])
def test_salt_proposal(oracle, golden_pure, prec):
    g = golden_pure[prec]
    for A, idx, lprop, adj, p, out in zip(g["salt_A"], g["salt_idx"], g["salt_logit_prop"],
                                          g["salt_log_adj"], g["salt_p"], g["salt_prop_p"]):
        gp, ga = oracle.salt(prec, p[:A], int(idx), float(lprop))
        assert close(gp, out[:A], prec) and close(ga, adj, prec)


@pytest.mark.parametrize("prec", [32, 64])
def test_priors(oracle, golden_pure, prec):
    g = golden_pure[prec]
    for x, b11, b2, b3, g1, g2 in g["prior_beta_gamma"]:
        assert close(oracle.beta_log_prior(prec, x, 1, 1), b11, prec)
        assert close(oracle.beta_log_prior(prec, x, .1, 9.9), b2, prec)
        assert close(oracle.beta_log_prior(prec, x, 2.5, 1.0), b3, prec)
        assert close(oracle.gamma_log_prior(prec, x * 20, .1, 10), g1, prec)
        assert close(oracle.gamma_log_prior(prec, x * 20, 2.0, 1.5), g2, prec)
    for x, lam, val in g["prior_dztpois"]:
        assert close(oracle.dztpois(prec, int(x), lam), val, prec)
    for x, size, pr, val in g["prior_dbinom"]:
        assert close(oracle.dbinom(prec, int(x), int(size), pr), val, prec)


@pytest.mark.parametrize("prec", [32, 64])
def test_whole_chain_states(oracle, golden_states, prec):
    """Per-cell log-likelihoods of full chain states dumped after k reference steps."""
    for name, st in golden_states[prec].items():
        allow = int(st["allow_relatedness"])
        t, o = oracle.eval_cells(prec, st["num_alleles"], st["obs_mask"], st["is_missing"],
                                 st["latent"], st["m"], st["r"], st["eps_neg"], st["eps_pos"],
                                 st["p"], allow)
        if prec == 32:
            cell = (t.astype(np.float32) + o.astype(np.float32)).astype(np.float64)
        else:
            cell = t + o
        assert close(cell, st["cell_llik"], prec), name
        assert np.all(cell[st["is_missing"] != 0] == 0)
        # k <= m everywhere and the index <-> bitmask round trip is loss-free
        k = np.vectorize(lambda x: bin(int(x)).count("1"))(st["latent"])
        assert np.all(k >= 1) and np.all(k <= st["m"][:, None]), name
        # Chain::llik is the fp sum of the cells (src/chain.cpp:978-991)
        assert abs(cell.sum() - st["llik"]) <= (2e-3 if prec == 32 else 1e-9) * abs(st["llik"])
