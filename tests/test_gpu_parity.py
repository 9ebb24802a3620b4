"""Parity of the CUDA engine, called through the C ABI, against the oracle and the golden
vectors minted from the reference (fp64; tolerance 1e-9 relative as BASELINE.json states --
observed differences are at the libm level, ~1e-15)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.array_equal(a[~fin], b[~fin], equal_nan=True)
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(1.0, np.abs(b[fin]))))


def engine_for(st_or_pk, temps=(1.0,), allow=1, seed=3, **kw):
    from moire_b200.data import PackedData
    from moire_b200.engine import Engine
    if isinstance(st_or_pk, dict):
        st = st_or_pk
        obs = st["obs_mask"]
        pk = PackedData(obs.shape[0], obs.shape[1], st["num_alleles"].astype(np.int32), obs,
                        st["is_missing"].astype(np.uint8),
                        np.array([max(bin(int(x)).count("1") for x in row) for row in obs],
                                 np.int32))
    else:
        pk = st_or_pk
    return Engine(pk, temps=list(temps), seed=seed, allow_relatedness=bool(allow), **kw), pk


def test_cells_match_reference_golden_states(golden_states):
    """Identical chain state dumped from the reference (fp64 variant) -> per-cell llik."""
    for name, st in golden_states[64].items():
        eng, pk = engine_for(st, allow=int(st["allow_relatedness"]))
        eng.load_state(0, dict(m=st["m"], r=st["r"], eps_pos=st["eps_pos"], eps_neg=st["eps_neg"],
                               p=st["p"], latent=st["latent"], lg_adj=st["lg_adj"],
                               mean_coi=float(st["mean_coi"])))
        out = eng.dump_state(0)
        cell = out["cell_t"] + out["cell_o"]
        assert rel_err(cell, st["cell_llik"]) < RTOL, name
        assert np.array_equal(out["latent"], st["latent"])          # bit-exact mask round trip
        pri = st["priors"]
        assert rel_err(out["prior_eps_neg"], pri[0]) < RTOL
        assert rel_err(out["prior_eps_pos"], pri[1]) < RTOL
        assert rel_err(out["prior_coi"], pri[2]) < RTOL
        assert rel_err(out["prior_r"], pri[3]) < RTOL
        assert rel_err(out["mean_coi_hyper_prior"], st["mean_coi_hyper_prior"]) < RTOL
        llik, prior = eng.scalars()
        assert abs(llik[0] - st["llik"]) < 1e-9 * abs(st["llik"]), name
        assert abs(prior[0] - st["prior"]) < 1e-9 * max(1.0, abs(st["prior"])), name
        assert rel_err(eng.sample_llik(0), cell.sum(axis=1)) < RTOL
        eng.close()


@pytest.mark.parametrize("allow", [1, 0])
def test_cells_match_oracle_on_device_drawn_states(oracle, allow):
    """Starting states drawn on the device, then advanced by full sweeps, recomputed by the
    oracle: covers IE (k<=10), the n==k closed form, EGF (k>10) and the relatedness mixture."""
    from moire_b200.data import pack_data
    from moire_b200.simulate import simulate_data
    d = simulate_data(sample_cois=np.r_[np.arange(1, 31), np.arange(1, 31)], epsilon_pos=.03,
                      epsilon_neg=.08,
                      locus_freq_alphas=[np.ones(30)] * 5 + [np.ones(8)] * 5 + [np.ones(2)] * 2,
                      internal_relatedness_alpha=1, internal_relatedness_beta=1,
                      missingness=.03, seed=41)
    pk = pack_data(d["data"], d["is_missing"])
    eng, _ = engine_for(pk, temps=(1.0, 0.3, 0.0), allow=allow, burnin=50)
    info = eng.init_chains()
    assert not info["still_ill"].any()
    kmax = 0
    for sweep in range(3):
        for c in range(3):
            st = eng.dump_state(c)
            t, o = oracle.eval_cells(64, pk.num_alleles, pk.obs_mask, pk.is_missing, st["latent"],
                                     st["m"], st["r"], st["eps_neg"], st["eps_pos"],
                                     st["p"][:, :pk.max_alleles], allow)
            assert rel_err(st["cell_t"], t) < RTOL and rel_err(st["cell_o"], o) < RTOL
            k = np.vectorize(lambda x: bin(int(x)).count("1"))(st["latent"])
            assert (k >= 1).all() and (k <= st["m"][:, None]).all()
            kmax = max(kmax, int(k.max()))
            assert np.allclose(st["p"][:, :pk.max_alleles].sum(1), 1.0, atol=1e-9)
        eng.step(sweep)
    assert kmax > 10  # the EGF branch was exercised
    eng.close()


def test_runs_are_reproducible_and_sharding_invariant():
    """Streams are keyed by global chain id: two engines holding halves of a ladder reproduce
    the chains of one engine holding all of it, bit for bit."""
    from moire_b200.data import pack_data
    from moire_b200.simulate import simulate_data
    d = simulate_data(mean_coi=3, num_samples=40, epsilon_pos=.02, epsilon_neg=.1,
                      locus_freq_alphas=[np.ones(6)] * 7, internal_relatedness_alpha=.2,
                      internal_relatedness_beta=1, seed=5)
    pk = pack_data(d["data"], d["is_missing"])
    temps = [1.0, 0.7, 0.4, 0.1]
    from moire_b200.engine import Engine
    whole = Engine(pk, temps, seed=9, burnin=20)
    lo = Engine(pk, temps[:2], seed=9, burnin=20, chain_offset=0)
    hi = Engine(pk, temps[2:], seed=9, burnin=20, chain_offset=2)
    for e in (whole, lo, hi):
        e.init_chains()
        for it in range(25):
            e.step(it)
    lw, pw = whole.scalars()
    ll, pl = lo.scalars()
    lh, ph = hi.scalars()
    assert np.array_equal(lw, np.r_[ll, lh]) and np.array_equal(pw, np.r_[pl, ph])
    a, b = whole.dump_state(3), hi.dump_state(1)
    for k in ("m", "r", "eps_pos", "eps_neg", "p", "latent", "lg_adj", "cell_t", "cell_o"):
        assert np.array_equal(a[k], b[k]), k
    for e in (whole, lo, hi):
        e.close()
