"""The restatement against the compiled reference itself, on fresh random inputs.  Runs wherever
oracle/_ref/*.so exists (built here from /root/reference; the .so travels to the GPU box)."""
import numpy as np
import pytest

from oracle.pyoracle import Ref, ref_available


@pytest.mark.parametrize("prec", [32, 64])
def test_random_pam_and_chain_state(oracle, prec):
    if not ref_available(prec):
        pytest.skip("oracle/_ref not built")
    from moire_b200.data import pack_data
    from moire_b200.simulate import simulate_data
    ref = Ref(prec)
    rng = np.random.default_rng(77 + prec)
    for k in range(1, 15):
        q = rng.dirichlet(np.ones(k) * 0.7)
        if prec == 32:
            q = q.astype(np.float32).astype(np.float64)
        assert np.array_equal(oracle.pam_vectorized(prec, q, 40), ref.pam_vectorized(q, 40))
        for n in (k, k + 1, k + 7):
            assert oracle.pam_scalar(prec, q, n) == ref.pam_scalar(q, n)
    d = simulate_data(mean_coi=5, num_samples=40, epsilon_pos=.05, epsilon_neg=.1,
                      locus_freq_alphas=[np.ones(7)] * 4 + [np.ones(25)] * 4,
                      internal_relatedness_alpha=.5, internal_relatedness_beta=1,
                      missingness=.05, seed=5)
    pk = pack_data(d["data"], d["is_missing"])
    for allow in (1, 0):
        ref.set_data(pk.num_alleles, pk.obs_mask, pk.is_missing)
        ref.set_params(allow_relatedness=allow)
        h = ref.chain_new(1.0, seed=8)
        for it in range(3):
            ref.chain_step(h, it)
        st = ref.chain_get_state(h)
        ref.chain_free(h)
        t, o = oracle.eval_cells(prec, pk.num_alleles, pk.obs_mask, pk.is_missing, st["latent"],
                                 st["m"], st["r"], st["eps_neg"], st["eps_pos"], st["p"], allow)
        cell = (t.astype(np.float32) + o.astype(np.float32)).astype(np.float64) if prec == 32 \
            else t + o
        assert np.array_equal(cell, st["cell_llik"])
