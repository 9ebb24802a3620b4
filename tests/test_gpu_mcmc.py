"""The drop-in call on the GPU: run_mcmc's result list, the engine-backed driver, and
distributional agreement with the reference's own sampler (compiled under oracle/_ref)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ["llik_burnin", "llik_sample", "prior_burnin", "prior_sample", "posterior_burnin",
          "posterior_sample", "data_llik", "coi", "lam_coi", "allele_freqs", "eps_neg", "eps_pos",
          "relatedness", "latent_genotypes", "observed_coi", "swap_store", "swap_acceptances",
          "swap_barriers", "temp_gradient", "acceptance_rates", "sampling_variances",
          "max_runtime_reached", "total_runtime"]     # src/main.cpp:164-186


def small_data(seed=31, N=60):
    from moire_b200.simulate import simulate_data
    return simulate_data(mean_coi=3, num_samples=N, epsilon_pos=.01, epsilon_neg=.08,
                         locus_freq_alphas=[np.ones(5)] * 8 + [np.ones(9)] * 6,
                         internal_relatedness_alpha=.1, internal_relatedness_beta=1,
                         missingness=.03, seed=seed)


def test_run_mcmc_result_layout():
    from moire_b200.mcmc import run_mcmc
    d = small_data()
    res = run_mcmc(d, d["is_missing"], burnin=80, samples_per_chain=40, pt_chains=5, thin=2,
                   record_latent_genotypes=True, verbose=False, seed=4)
    assert set(res) == {"runtime", "chains", "args"} and len(res["chains"]) == 1
    ch = res["chains"][0]
    assert [f for f in FIELDS if f not in ch] == [] and "mean_coi" in ch
    N, L = 60, 14
    D = len(ch["lam_coi"])
    assert 1 <= D <= 20                                   # thin = 2 of 40 sampling steps
    assert len(ch["llik_burnin"]) == 80 and len(ch["llik_sample"]) == 40
    assert len(ch["swap_store"]) == 120 and len(ch["temp_gradient"]) == 5
    assert np.all(np.isfinite(ch["llik_sample"])) and np.all(np.isfinite(ch["posterior_burnin"]))
    assert len(ch["coi"]) == N and len(ch["coi"][0]) == D and ch["coi"][0].dtype.kind == "i"
    assert len(ch["allele_freqs"]) == L and len(ch["allele_freqs"][0]) == D
    assert len(ch["allele_freqs"][0][0]) == 5 and len(ch["allele_freqs"][13][0]) == 9
    assert abs(ch["allele_freqs"][3][-1].sum() - 1) < 1e-9
    lg = ch["latent_genotypes"]
    assert len(lg) == N and len(lg[0]) == L and len(lg[0][0]) == D
    assert all(1 <= len(lg[i][j][-1]) <= ch["coi"][i][-1] for i in range(N) for j in range(L))
    assert len(ch["acceptance_rates"]) == 5 and len(ch["sampling_variances"]) == 5
    assert set(ch["acceptance_rates"][0]) == {"allele_freq_accept", "coi_accept", "eps_neg_accept",
                                              "eps_pos_accept", "r_accept", "m_r_accept",
                                              "full_sample_accept", "mean_coi_accept"}
    assert ch["temp_gradient"][0] == 1.0 and np.all(np.diff(ch["temp_gradient"]) <= 0)
    assert np.allclose(ch["mean_coi"], ch["lam_coi"] / (1 - np.exp(-ch["lam_coi"])))
    # per-sample data_llik sums to the trace of the stored steps
    tot = np.array([x for x in np.sum(ch["data_llik"], axis=0)])
    assert np.allclose(tot, np.asarray(ch["llik_sample"])[::2][-D:], rtol=1e-9)


def test_single_chain_stores_every_draw_and_is_reproducible():
    from moire_b200.mcmc import run_mcmc
    d = small_data(N=30)
    a = run_mcmc(d, d["is_missing"], burnin=30, samples_per_chain=12, verbose=False, seed=9)
    b = run_mcmc(d, d["is_missing"], burnin=30, samples_per_chain=12, verbose=False, seed=9)
    c = run_mcmc(d, d["is_missing"], burnin=30, samples_per_chain=12, verbose=False, seed=10)
    assert len(a["chains"][0]["lam_coi"]) == 12
    assert np.array_equal(a["chains"][0]["llik_sample"], b["chains"][0]["llik_sample"])
    assert not np.array_equal(a["chains"][0]["llik_sample"], c["chains"][0]["llik_sample"])


@pytest.mark.parametrize("T,latent", [(5, True), (1, False)])
def test_device_ladder_reproduces_host_ladder(monkeypatch, T, latent):
    """Row f3: swap pass, barrier accumulators, traces, hot flags and the stored draws on the device
    (csrc/ladder.cu; one graph launch per step, batched read-back) against the same run with the
    ladder on the host after a per-step read-back (MOIRE_B200_HOST_LADDER=1, round 1's path, itself
    pinned to the reference's MCMC::swap_chains by tests/test_mcmc_driver.py).  Same streams, one
    swap_pair() for both: every integer result identical; floating results identical or, where the
    device's exp/log feed the barrier sums and through them the adapted ladder, equal to rounding."""
    from moire_b200.mcmc import run_mcmc
    d = small_data(N=50)
    kw = dict(burnin=90, samples_per_chain=45, pt_chains=T, thin=2, record_latent_genotypes=latent,
              pre_adapt_steps=10, temp_adapt_steps=7, verbose=False, seed=6)
    dev = run_mcmc(d, d["is_missing"], **kw)["chains"][0]
    monkeypatch.setenv("MOIRE_B200_HOST_LADDER", "1")
    host = run_mcmc(d, d["is_missing"], **kw)["chains"][0]
    assert len(dev["swap_store"]) == len(host["swap_store"]) == (135 if T > 1 else 45)
    for k in ("swap_store", "swap_acceptances", "coi", "observed_coi"):
        assert np.array_equal(np.asarray(dev[k]), np.asarray(host[k])), k
    for k in ("llik_burnin", "prior_burnin", "posterior_burnin", "llik_sample", "prior_sample",
              "posterior_sample", "lam_coi", "temp_gradient", "swap_barriers", "eps_neg", "eps_pos",
              "relatedness", "data_llik"):
        a, b = np.asarray(dev[k], float), np.asarray(host[k], float)
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-10, atol=1e-12, equal_nan=True), k
    assert len(dev["lam_coi"]) == len(host["lam_coi"]) and (T > 1 or len(dev["lam_coi"]) == 23)
    for la, lb in zip(dev["allele_freqs"], host["allele_freqs"]):
        assert np.allclose(np.asarray(la), np.asarray(lb), rtol=1e-10)
    if latent:
        assert dev["latent_genotypes"] == host["latent_genotypes"]
    assert dev["acceptance_rates"][0]["coi_accept"].tolist() == host["acceptance_rates"][0]["coi_accept"].tolist()


def test_progress_and_stop_with_the_ladder_on_the_device():
    """moire_mcmc_run in batches: the callback still hears about every step, each with that step's
    own posterior (the trace entry), and a stop request ends the run at the end of its batch."""
    from moire_b200.engine import MoireError
    from moire_b200.mcmc import run_mcmc
    d = small_data(N=30)
    seen = []
    res = run_mcmc(d, d["is_missing"], burnin=70, samples_per_chain=10, pt_chains=3, verbose=False,
                   seed=2, progress=lambda ph, st, post, hot: seen.append((ph, st, post, hot)) and False)
    ch = res["chains"][0]
    assert [(p, s) for p, s, _, _ in seen] == [(0, s) for s in range(1, 71)] + [(1, s) for s in range(1, 11)]
    assert np.array_equal([x[2] for x in seen[:70]], ch["posterior_burnin"])
    assert np.array_equal([x[3] for x in seen], ch["swap_store"])
    with pytest.raises(MoireError, match="interrupted"):
        run_mcmc(d, d["is_missing"], burnin=70, samples_per_chain=10, pt_chains=3, verbose=False, seed=2,
                 progress=lambda ph, st, post, hot: ph == 0 and st == 40)


def test_posterior_agrees_with_reference_sampler():
    """Posterior COI / allele-frequency / error-rate summaries from the engine against the
    reference's own C++ sampler on the same data, three reference seeds and six engine seeds: the
    mean COI must agree within four standard errors of the difference, estimated from the
    between-seed spread of each sampler (with a floor: three seeds give a rough sd); the other
    summaries within fixed bands a few times their Monte-Carlo error."""
    from oracle.pyoracle import Ref, ref_available
    if not ref_available(32):
        pytest.skip("oracle/_ref not built")
    from moire_b200.data import pack_data
    from moire_b200.mcmc import run_mcmc
    d = small_data(seed=77, N=80)
    pk = pack_data(d["data"], d["is_missing"])
    burn, samp = 600, 600
    ref = Ref(32)
    ref.set_data(pk.num_alleles, pk.obs_mask, pk.is_missing)
    ref.set_params(pt_chains=[1.0], burnin=burn, samples=samp)
    refs = []
    for seed in (1, 2, 3):
        h = ref.mcmc_new(seed=seed)
        for s in range(burn):
            ref.lib.ref_mcmc_burnin(h, s)
        for s in range(samp):
            ref.lib.ref_mcmc_sample(h, s)
        refs.append(ref.mcmc_summary(h))
        ref.mcmc_free(h)
    ours = []
    for seed in (11, 12, 13, 14, 15, 16):
        ch = run_mcmc(d, d["is_missing"], burnin=burn, samples_per_chain=samp, verbose=False,
                      seed=seed, pipeline="flat" if seed % 2 else "fused")["chains"][0]
        ours.append(dict(m=np.mean(ch["coi"], axis=1), r=np.mean(ch["relatedness"], axis=1),
                         eps_neg=np.mean(ch["eps_neg"], axis=1), eps_pos=np.mean(ch["eps_pos"], axis=1),
                         p0=np.array([np.mean(ch["allele_freqs"][j], axis=0)[0] for j in range(pk.num_loci)]),
                         mean_coi=float(np.mean(ch["lam_coi"]))))
    ref_m = np.mean([r["m"] for r in refs], axis=0)
    our_m = np.mean([o["m"] for o in ours], axis=0)
    ref_spread = np.abs(refs[0]["m"] - refs[1]["m"]).mean()
    ref_runs = np.array([r["m"].mean() for r in refs])
    our_runs = np.array([o["m"].mean() for o in ours])
    se = np.sqrt(max(ref_runs.std(ddof=1), 0.1) ** 2 / len(ref_runs) +
                 max(our_runs.std(ddof=1), 0.1) ** 2 / len(our_runs))
    assert abs(ref_m.mean() - our_m.mean()) < 4 * se, (ref_runs, our_runs, se)
    assert np.abs(ref_m - our_m).mean() < max(4 * ref_spread, 0.25), (np.abs(ref_m - our_m).mean(), ref_spread)
    assert np.corrcoef(ref_m, our_m)[0, 1] > 0.95
    ref_p0 = np.mean([r["p"][:, 0] for r in refs], axis=0)
    our_p0 = np.mean([o["p0"] for o in ours], axis=0)
    assert np.abs(ref_p0 - our_p0).max() < 0.03, np.abs(ref_p0 - our_p0).max()
    for k in ("eps_neg", "eps_pos", "r"):
        a = np.mean([r[k] for r in refs], axis=0).mean()
        b = np.mean([o[k] for o in ours], axis=0).mean()
        assert abs(a - b) < 0.03, (k, a, b)
    assert abs(np.mean([r["mean_coi"] for r in refs]) - np.mean([o["mean_coi"] for o in ours])) < 0.4


def packaged(prefix):
    from conftest import load_golden
    from moire_b200.data import PackedData
    g = load_golden("packaged_data.npz")
    obs = g[prefix + ".obs_mask"]
    coi = np.vectorize(lambda x: bin(int(x)).count("1"))(obs).max(axis=1).astype(np.int32)
    return PackedData(obs.shape[0], obs.shape[1], g[prefix + ".num_alleles"].astype(np.int32), obs,
                      g[prefix + ".is_missing"].astype(np.uint8), coi), g


def test_packaged_demo_run_reproduces_the_reference_authors_posterior():
    """The reference ships the result of its demo run on `simulated_data` (data/mcmc_results.rda,
    recipe data-raw/mcmc_results.R: 80 temperatures, 1000 + 1000 steps, eps priors Beta(.1, 9.9)).
    Same data, same arguments through `run_mcmc` here.  That fixture was produced by an OLDER build
    of moire, so it is a sanity band, not a pin: the current reference C++ itself
    (oracle/_ref, same data and priors, 8 temperatures) gives mean COI 6.85, lam 6.85, relatedness
    0.33 where the fixture has 6.22, 6.21, 0.27 -- and the engine lands on the current reference
    (6.81 / 6.8 / 0.33).  The log-likelihood at stationarity agrees with the fixture to well
    within its own trace sd."""
    from moire_b200.mcmc import run_mcmc
    pk, g = packaged("sim")
    a = {k[len("res.arg."):]: float(v) for k, v in g.items() if k.startswith("res.arg.")}
    res = run_mcmc(pk, burnin=int(a["burnin"]), samples_per_chain=int(a["samples_per_chain"]),
                   pt_chains=g["res.pt_chains"], eps_pos_alpha=a["eps_pos_alpha"],
                   eps_pos_beta=a["eps_pos_beta"], eps_neg_alpha=a["eps_neg_alpha"],
                   eps_neg_beta=a["eps_neg_beta"], r_alpha=a["r_alpha"], r_beta=a["r_beta"],
                   mean_coi_shape=a["mean_coi_shape"], mean_coi_scale=a["mean_coi_scale"],
                   max_eps_pos=a["max_eps_pos"], max_eps_neg=a["max_eps_neg"], max_coi=int(a["max_coi"]),
                   pre_adapt_steps=int(a["pre_adapt_steps"]), temp_adapt_steps=int(a["temp_adapt_steps"]),
                   verbose=False, seed=17325)
    ch = res["chains"][0]
    assert len(ch["temp_gradient"]) == 80 and len(ch["lam_coi"]) > 900
    ll = np.asarray(ch["llik_sample"])
    assert abs(ll.mean() - g["res.llik_sample"].mean()) < 60, (ll.mean(), g["res.llik_sample"].mean())
    m = np.mean(ch["coi"], axis=1)
    assert abs(m.mean() - g["res.mean_coi_per_sample"].mean()) < 1.0     # older build: 6.22
    assert abs(m.mean() - 6.85) < 0.35                                   # current reference C++
    # two independent 1000-draw runs: 0.96-0.98 across engine seeds
    assert np.corrcoef(m, g["res.mean_coi_per_sample"])[0, 1] > 0.95
    assert abs(np.mean(ch["lam_coi"]) - 6.85) < 0.6
    assert abs(np.mean(ch["eps_neg"]) - g["res.mean_eps_neg"].mean()) < 0.002
    assert abs(np.mean(ch["eps_pos"]) - g["res.mean_eps_pos"].mean()) < 0.002
    assert abs(np.mean(ch["relatedness"]) - 0.326) < 0.05
    p0 = np.array([np.mean(ch["allele_freqs"][j], axis=0)[0] for j in range(30)])
    assert np.abs(p0 - g["res.mean_first_allele_freq"]).max() < 0.03
    # the adapted ladder keeps its end points and stays ordered
    assert ch["temp_gradient"][0] == 1.0 and ch["temp_gradient"][-1] == 0.0
    assert np.all(np.diff(ch["temp_gradient"]) <= 0)


def test_namibia_data_runs_with_ten_temperatures():
    """configs[1] of BASELINE.json on the real data shape: 2585 samples x 26 loci, A_j 7..35,
    11.4 % missing cells, 10 temperatures."""
    from moire_b200.mcmc import run_mcmc
    pk, _ = packaged("nam")
    res = run_mcmc(pk, burnin=30, samples_per_chain=10, pt_chains=10, verbose=False, seed=3)
    ch = res["chains"][0]
    assert np.all(np.isfinite(ch["llik_burnin"])) and np.all(np.isfinite(ch["posterior_sample"]))
    assert ch["llik_burnin"][-1] > ch["llik_burnin"][0]              # burn-in climbs
    assert len(ch["coi"]) == 2585 and len(ch["allele_freqs"]) == 26 and len(ch["allele_freqs"][22][0]) == 35


def test_acceptance_rates_and_adapted_scales_track_the_reference():
    """The Metropolis machinery as a whole (proposal rules, validity checks, adaptation toward
    0.23): after the same burn-in on the same data, the per-kind acceptance rates and the adapted
    proposal scales of the engine's chain agree with the reference's own Chain (oracle/_ref, fp32)
    -- two independent random streams, so the comparison is distributional (means over samples)."""
    from oracle.pyoracle import Ref, ref_available
    if not ref_available(32):
        pytest.skip("oracle/_ref not built")
    from moire_b200.data import pack_data
    from moire_b200.engine import Engine
    d = small_data(seed=5, N=120)
    pk = pack_data(d["data"], d["is_missing"])
    steps = 500
    ref = Ref(32)
    ref.set_data(pk.num_alleles, pk.obs_mask, pk.is_missing)
    ref.set_params(burnin=steps)
    h = ref.chain_new(1.0, seed=3)
    for it in range(steps):
        ref.chain_step(h, it)
    racc, rsd = ref.chain_get_tuning(h)
    rp_acc, rp_att, rp_sd = ref.chain_get_p_tuning(h)
    ref.chain_free(h)
    eng = Engine(pk, [1.0], seed=8, burnin=steps, pipeline="flat")   # the benchmarked kernels
    eng.init_chains()
    for it in range(steps):
        eng.step(it)
    dg = eng.diagnostics(0)
    eng.close()
    names = ["eps_neg_accept", "eps_pos_accept", "m_accept", "r_accept", "m_r_accept", "sample_accept"]
    for row, name in enumerate(names):
        a, b = racc[row].mean() / steps, dg[name].mean() / steps
        assert abs(a - b) < 0.04, (name, a, b)
    for row, name in enumerate(["eps_neg_var", "eps_pos_var", "r_var", "m_r_var"]):
        a, b = np.median(rsd[row]), np.median(dg[name])
        assert abs(a - b) < 0.25 * max(a, b) + 0.02, (name, a, b)
    # the SALT sampler (update_p): per-allele attempts, acceptance rate and adapted scale
    A = pk.num_alleles
    sel = np.arange(rp_att.shape[1])[None, :] < A[:, None]
    # A_j proposals per locus and step, fewer where a proposal fell below 1e-12 and the locus was
    # left early (src/chain.cpp:503-518)
    for att in (rp_att[sel].sum(), dg["p_attempt"][:, :rp_att.shape[1]][sel].sum()):
        assert 0.8 * steps * A.sum() < att <= steps * A.sum(), att
    ra = rp_acc[sel].sum() / rp_att[sel].sum()
    ea = dg["p_accept"][:, :rp_att.shape[1]][sel].sum() / dg["p_attempt"][:, :rp_att.shape[1]][sel].sum()
    assert abs(ra - ea) < 0.03 and 0.15 < ea < 0.35, (ra, ea)
    rs, es = np.median(rp_sd[sel]), np.median(dg["p_prop_var"][:, :rp_att.shape[1]][sel])
    assert abs(rs - es) < 0.2 * max(rs, es), (rs, es)
    # the adapted single-parameter updates sit near the 0.23 target on both sides
    for row in (0, 1, 3):
        assert 0.15 < dg[names[row]].mean() / steps < 0.35
