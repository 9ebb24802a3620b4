"""Full-size shapes and edge cases on the GPU: size-independent properties of the engine state
(idempotent re-evaluation, per-chain sums, simplex, k <= m, counters) on the BASELINE.json
configurations, plus spot checks of individual cells against the oracle and degenerate inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def popcount(a, pk=None):
    """Set bits per allele set; masks of W words ([..][W], the two-word build) are summed over words."""
    bits = np.vectorize(lambda x: bin(int(x)).count("1"))(a)
    return bits.sum(axis=-1) if pk is not None and pk.mask_words > 1 else bits


def bits_above(mask_words, A):
    """True if an allele set (its words, lowest first) has a bit at or above allele A."""
    x = sum(int(w) << (64 * i) for i, w in enumerate(np.atleast_1d(mask_words)))
    return (x >> int(A)) != 0


def check_invariants(eng, pk, chain, oracle=None, allow=1, ncheck=400, seed=0):
    st = eng.dump_state(chain)
    cell = st["cell_t"] + st["cell_o"]
    assert np.all(np.isfinite(cell))
    assert np.all(cell[pk.is_missing != 0] == 0.0)                 # missing cells contribute nothing
    k = popcount(st["latent"], pk)
    assert (k >= 1).all() and (k <= st["m"][:, None]).all()
    A = pk.num_alleles
    assert not any(bits_above(x, A[j]) for j in range(pk.num_loci) for x in st["latent"][:, j][:50])
    assert np.allclose(st["p"][:, :pk.max_alleles].sum(1), 1.0, atol=1e-9)
    assert (st["p"][np.arange(pk.num_loci)[:, None], np.arange(st["p"].shape[1])[None, :]]
            [np.arange(st["p"].shape[1])[None, :] >= A[:, None]] == 0).all()   # padding stays 0
    assert (st["m"] >= 1).all() and (st["m"] <= eng.model["max_coi"]).all()
    assert ((st["eps_neg"] > 0) & (st["eps_neg"] < 1) & (st["eps_pos"] > 0) & (st["eps_pos"] < 1)).all()
    # the per-chain sums are the sums of the cells / priors
    llik, prior = eng.scalars()
    assert abs(llik[chain] - cell.sum()) <= 1e-9 * max(1.0, abs(llik[chain]))
    pri = (st["prior_eps_neg"] + st["prior_eps_pos"] + st["prior_r"] + st["prior_coi"]).sum() + \
        st["mean_coi_hyper_prior"]
    assert abs(prior[chain] - pri) < 1e-9 * max(1.0, abs(prior[chain]))
    assert np.allclose(eng.sample_llik(chain), cell.sum(1), rtol=1e-12, atol=1e-9)
    if oracle is not None:   # a random subset of samples, every cell, against the CPU restatement
        rng = np.random.default_rng(seed)
        rows = np.sort(rng.choice(pk.num_samples, size=min(pk.num_samples, max(1, ncheck // pk.num_loci)),
                                  replace=False))
        t, o = oracle.eval_cells(64, pk.num_alleles, pk.obs_mask[rows], pk.is_missing[rows],
                                 st["latent"][rows], st["m"][rows], st["r"][rows], st["eps_neg"][rows],
                                 st["eps_pos"][rows], st["p"][:, :pk.max_alleles], allow)
        err_t = np.abs(st["cell_t"][rows] - t) / np.maximum(1.0, np.abs(t))
        err_o = np.abs(st["cell_o"][rows] - o) / np.maximum(1.0, np.abs(o))
        assert err_t.max() < 1e-9 and err_o.max() < 1e-9, (err_t.max(), err_o.max())
    return st, k


@pytest.mark.parametrize("name,chains,steps", [("c2", 3, 3), ("c3", 3, 3), ("c4", 2, 2), ("c5", 3, 3)])
def test_baseline_shapes_hold_their_invariants(oracle, name, chains, steps):
    """The named shapes of BASELINE.json (Namibia-like A_j up to 35, 1000x100, 5000x250 with A_j up
    to 30, COI up to 30): a few chains of the ladder, a few full sweeps, then the properties, and
    idempotence of a full re-evaluation of the cached terms."""
    from moire_b200.engine import Engine
    from moire_b200.workloads import make_workload
    wl = make_workload(name)
    pk = wl["packed"]
    temps = wl["temps"][np.linspace(0, len(wl["temps"]) - 1, chains).astype(int)]
    eng = Engine(pk, temps, seed=5, burnin=100, **wl["model"])
    info = eng.init_chains()
    assert not info["still_ill"].any()
    ev0, _ = eng.counters()
    for it in range(steps):
        eng.step(it)
    ev1, _ = eng.counters()
    N, L, nonmiss = pk.num_samples, pk.num_loci, int((pk.is_missing == 0).sum())
    # upper bound of evaluations per chain-step: 6 N L + N sum_j A_j, missing cells excluded
    per_locus_nonmiss = (pk.is_missing == 0).sum(0)
    bound = chains * steps * (6 * nonmiss + int((per_locus_nonmiss * pk.num_alleles).sum()))
    assert 0.5 * bound < ev1 - ev0 <= bound
    kmax = 0
    for c in range(chains):
        st, k = check_invariants(eng, pk, c, oracle, ncheck=600, seed=c)
        kmax = max(kmax, int(k.max()))
        before = (st["cell_t"].copy(), st["cell_o"].copy())
    if name == "c5":
        assert kmax > 10          # the EGF branch is live at this shape
    l0, p0 = eng.scalars()
    eng.eval_all()                # recompute every cached term from the parameters
    l1, p1 = eng.scalars()
    assert np.allclose(l0, l1, rtol=1e-12) and np.allclose(p0, p1, rtol=1e-12)
    st2 = eng.dump_state(chains - 1)
    assert np.allclose(st2["cell_t"], before[0], rtol=1e-12, atol=1e-12)
    assert np.array_equal(st2["cell_o"], before[1])
    eng.close()


def _mini(N, As, seed=3, **kw):
    from moire_b200.data import pack_data
    from moire_b200.simulate import simulate_data
    base = dict(mean_coi=3, num_samples=N, epsilon_pos=.02, epsilon_neg=.1,
                locus_freq_alphas=[np.ones(a) for a in As], internal_relatedness_alpha=.2,
                internal_relatedness_beta=1, seed=seed)
    base.update(kw)
    d = simulate_data(**base)
    return pack_data(d["data"], d["is_missing"])


@pytest.mark.parametrize("N,As,T,kw", [
    (1, [2], 1, {}),                                   # one sample, one biallelic locus, one chain
    (3, [64, 2, 33], 2, {}),                           # the widest mask word
    (30, [100, 2, 65, 128, 64], 2, {}),                # past one word: the two-word build (128 alleles)
    (25, [128] * 3 + [90], 2, dict(mean_coi=14, missingness=.1)),   # ... with k > 10 (EGF) cells
    (70, [5] * 3, 3, dict(missingness=1.0)),           # every cell missing
    (33, [7] * 40, 2, dict(missingness=.5)),           # half missing, ragged tile (33 = 32 + 1)
    (2049, [4, 6], 2, {}),                             # N > 2048: the 512-thread update_p variant
    (40, [6] * 1100, 1, {}),                           # L > 1024: one sample per tile, long rows
])
def test_edge_shapes(oracle, N, As, T, kw):
    from moire_b200.engine import Engine
    pk = _mini(N, As, **kw)
    eng = Engine(pk, np.linspace(1, 0, T) if T > 1 else [1.0], seed=2, burnin=30, max_coi=20)
    assert not eng.init_chains()["still_ill"].any()
    for it in range(3):
        eng.step(it)
    for c in range(T):
        check_invariants(eng, pk, c, oracle, ncheck=300)
    eng.close()


def test_two_word_build_runs_the_one_word_chain_bit_for_bit():
    """libmoire_b200_w128.so is the same sources with two-word allele sets (mask.cuh).  On data whose
    loci fit one word it must run the very chain the one-word library runs -- same Philox streams,
    same arithmetic, same order -- so everything the one-word build is pinned to (the CPU stepper,
    the oracle, the reference's goldens) carries over to it: every state array bit-identical after
    several full sweeps, the second mask word empty."""
    from moire_b200.engine import Engine
    pk = _mini(60, [5] * 4 + [12] * 4 + [40, 64], seed=4, missingness=.1, mean_coi=4)
    temps = [1.0, .6, .2]
    states = []
    for data in (pk, pk.widened(2)):
        eng = Engine(data, temps, seed=12, burnin=40, max_coi=25, pipeline="flat")
        assert eng.lib.moire_engine_mask_words() == data.mask_words
        assert not eng.init_chains()["still_ill"].any()
        for it in range(6):
            eng.step(it)
        states.append(([eng.dump_state(c) for c in range(len(temps))], eng.scalars(), eng.diagnostics(0)))
        eng.close()
    (one, sc1, dg1), (two, sc2, dg2) = states
    for a, b in zip(one, two):
        for name in a:
            if name == "latent":
                assert np.array_equal(a[name], b[name][:, :, 0]) and not b[name][:, :, 1].any()
            else:
                assert np.array_equal(np.asarray(a[name]), np.asarray(b[name])), name
    assert np.array_equal(sc1[0], sc2[0]) and np.array_equal(sc1[1], sc2[1])
    for name in dg1:
        assert np.array_equal(dg1[name], dg2[name]), name


def test_run_mcmc_on_a_locus_with_more_than_64_alleles():
    """The user-facing call on data the one-word masks cannot hold (the reference has no allele limit,
    src/genotyping_data.cpp:17-48): the 23-field result comes back, recorded latent genotypes are
    index lists below each locus' allele count and within the sampled COI."""
    from moire_b200.mcmc import run_mcmc
    from moire_b200.simulate import simulate_data
    As = [100, 8, 70]
    d = simulate_data(mean_coi=3, num_samples=20, epsilon_pos=.02, epsilon_neg=.1,
                      locus_freq_alphas=[np.ones(a) for a in As], seed=5)
    res = run_mcmc(d, d["is_missing"], burnin=30, samples_per_chain=10, pt_chains=3,
                   record_latent_genotypes=True, verbose=False)
    ch = res["chains"][0]
    assert [np.asarray(f[0]).shape[0] for f in ch["allele_freqs"]] == As
    assert all(abs(np.asarray(f[-1]).sum() - 1) < 1e-9 for f in ch["allele_freqs"])
    lg = ch["latent_genotypes"]
    for i in range(20):
        for j, A in enumerate(As):
            for dr in range(10):
                g = lg[i][j][dr]
                assert len(g) >= 1 and all(0 <= a < A for a in g) and len(g) <= ch["coi"][i][dr]
    assert np.all(np.isfinite(np.asarray(ch["data_llik"])))


def test_salt_treats_alleles_past_the_first_mask_word_like_the_others():
    """update_p on 100-allele loci with every cell missing: the likelihood is constant, so SALT
    (src/chain.cpp:438-565) random-walks p under a law that is exchangeable in the alleles (the index
    of a proposal is uniform, the update is symmetric in the others).  The two-word build keeps four
    alleles per lane in the proposal (p_flat.cuh: kAPL); the alleles 64..99 must be proposed, accepted
    and adapted like the alleles 0..63, hold the same mass and move as much."""
    from moire_b200.engine import Engine
    A, steps = 100, 160
    pk = _mini(6, [A] * 8, missingness=1.0)
    eng = Engine(pk, [1.0], seed=5, burnin=steps, max_coi=10)
    assert eng.W == 2 and not eng.init_chains()["still_ill"].any()
    trace = []
    for it in range(steps):
        eng.update(2, it)                     # KIND_P
        if it >= 40:
            trace.append(eng.dump_state(0)["p"][:, :A].copy())
    trace = np.asarray(trace)                 # [step][locus][allele]
    assert np.allclose(trace.sum(axis=2), 1.0, atol=1e-9) and (trace > 0).all()
    d = eng.diagnostics(0)
    att, acc, var = d["p_attempt"][:, :A], d["p_accept"][:, :A], d["p_prop_var"][:, :A]
    lo, hi = slice(0, 64), slice(64, A)
    # a locus proposes until every allele had its turn or a frequency would leave [1e-12, 1] (:503-518)
    # (with nothing observed the frequencies wander down to that floor, so most sweeps end early)
    assert (att.sum(axis=1) <= A * steps).all() and (att.sum(axis=1) > 10 * steps).all()
    assert 0.9 < att[:, hi].mean() / att[:, lo].mean() < 1.1            # the index is uniform
    per_allele = att.sum(axis=0)                                         # over the 8 loci: ~240 each
    assert per_allele[hi].min() > 0.7 * per_allele[lo].mean() and per_allele[lo].min() > 0.7 * per_allele[hi].mean()
    rate = acc.sum(axis=0) / att.sum(axis=0)                             # per allele, over the loci
    assert abs(rate[hi].mean() - rate[lo].mean()) < 0.03 and rate[hi].min() > 0.05
    assert (var[att > 16] != 1.0).all() and (att > 16).mean() > 0.9     # the scales have adapted (:555-562)
    assert 0.8 < np.median(var[:, hi]) / np.median(var[:, lo]) < 1.25
    # exchangeable: the two groups hold the same mass per allele and move as much
    mean = trace.mean(axis=(0, 1))
    assert abs(mean[hi].mean() / mean[lo].mean() - 1.0) < 0.2
    move = np.abs(np.diff(np.log(trace), axis=0)).mean(axis=(0, 1))     # mean |step| of log p, per allele
    assert 0.8 < move[hi].mean() / move[lo].mean() < 1.25
    eng.close()


@pytest.mark.parametrize("max_coi,pipeline", [(64, "fused"), (100, "flat"), (127, "flat"), (127, "fused")])
def test_high_coi_limits(oracle, max_coi, pipeline):
    """max_coi up to the reference's own bound (127, its lgamma table, src/sampler.cpp:22-25): the
    wide mixture path (m - k >= 16), EGF bands wider than the scratch column and wider than a warp,
    cooperative sums over more than 64 mixture terms -- every cell against the oracle."""
    from moire_b200.engine import Engine
    pk = _mini(24, [30] * 4 + [3] * 4, sample_cois=np.arange(max_coi - 23, max_coi + 1), seed=9)
    eng = Engine(pk, [1.0, 0.2], seed=2, burnin=10, max_coi=max_coi, pipeline=pipeline)
    assert not eng.init_chains()["still_ill"].any()
    for it in range(2):
        eng.step(it)
    wide = 0
    for c in range(2):
        st, k = check_invariants(eng, pk, c, oracle, ncheck=10 ** 6)
        wide += int(((st["m"][:, None] - k + 1 > 16) & (k <= 10)).sum())
    assert wide > 0
    eng.close()


def test_argument_errors_are_reported_not_fatal():
    from moire_b200.engine import Engine, MoireError
    pk = _mini(5, [4, 4])
    with pytest.raises(MoireError, match="max_coi"):
        Engine(pk, [1.0], max_coi=128)
    with pytest.raises(MoireError, match="device_id"):
        Engine(pk, [1.0], device=99)
    with pytest.raises(MoireError, match="one-word masks only"):      # the two-word build is flat-only
        Engine(pk.widened(2), [1.0], pipeline="fused")
    wide_bad = _mini(5, [70, 4])
    wide_bad.obs_mask = wide_bad.obs_mask.copy()
    wide_bad.obs_mask[0, 0, 1] = 1 << 20                              # allele 84 of a 70-allele locus
    with pytest.raises(MoireError, match="bits above"):
        Engine(wide_bad, [1.0])
    bad = _mini(5, [4, 4])
    bad.obs_mask = bad.obs_mask.copy()
    bad.obs_mask[0, 0] = 1 << 10
    with pytest.raises(MoireError, match="bits above"):
        Engine(bad, [1.0])
    eng = Engine(pk, [1.0])
    with pytest.raises(MoireError, match="before init"):
        eng.step(0)
    eng.close()


def test_chain_states_do_not_depend_on_the_ladder_size():
    """Which evaluator a cell lands in (one thread or a cooperating warp) follows the launch's work
    per thread, i.e. the number of chains on the device; both produce the same bits, so a chain's
    trajectory is the same in a 2-chain and in a 6-chain engine (what makes sharding invisible)."""
    from moire_b200.engine import Engine
    from moire_b200.workloads import make_workload
    wl = make_workload("c3")
    pk = wl["packed"]
    temps = wl["temps"][np.linspace(0, len(wl["temps"]) - 1, 6).astype(int)]
    states = []
    for n in (2, 6):
        eng = Engine(pk, temps[:n], seed=9, burnin=100, **wl["model"])
        eng.init_chains()
        for it in range(3):
            eng.step(it)
        states.append([eng.dump_state(c) for c in range(2)])
        eng.close()
    for sa, sb in zip(*states):
        for key in sa:
            assert np.array_equal(np.asarray(sa[key]), np.asarray(sb[key]), equal_nan=True), key


def _run_pipeline(flat, pk, temps, model, steps, monkeypatch):
    from moire_b200.engine import Engine
    eng = Engine(pk, temps, seed=21, pipeline="flat" if flat else "fused", **model)
    eng.init_chains()
    for it in range(steps):
        eng.step(it)
    out = [eng.dump_state(c) for c in range(len(temps))]
    llik, prior = eng.scalars()
    eng.close()
    return out, llik, prior


def test_flat_and_fused_pipelines_agree_on_ragged_data_with_missing_cells(monkeypatch):
    """The two generations of kernels share streams and arithmetic: on a small ragged shape
    (A_j from 2 to 40, several distinct allele counts, 15 % missing cells, N not a multiple of
    anything) they reach the same integer state and the same terms (sums are taken in different
    orders, so the floating state is compared to 1e-9)."""
    from moire_b200.data import pack_data
    from moire_b200.engine import DEFAULT_MODEL
    from moire_b200.simulate import simulate_data
    rng = np.random.default_rng(3)
    L, N = 13, 157
    num_alleles = rng.integers(2, 41, size=L)
    sim = simulate_data(mean_coi=3.0, num_samples=N, epsilon_pos=.02, epsilon_neg=.05,
                        locus_freq_alphas=[np.full(int(a), .5) for a in num_alleles], seed=4)
    miss = rng.random((L, N)) < .15
    pk = pack_data(sim["data"], miss)
    temps = np.array([1.0, .6, .2])
    model = dict(DEFAULT_MODEL, max_coi=12, burnin=100)
    a, la, pa = _run_pipeline(True, pk, temps, model, 4, monkeypatch)
    b, lb, pb = _run_pipeline(False, pk, temps, model, 4, monkeypatch)
    assert np.allclose(la, lb, rtol=1e-9) and np.allclose(pa, pb, rtol=1e-9)
    for sa, sb in zip(a, b):
        for key in sa:
            x, y = np.asarray(sa[key]), np.asarray(sb[key])
            if x.dtype.kind in "iu":
                assert np.array_equal(x, y), key
            else:
                assert np.allclose(x, y, rtol=1e-9, atol=1e-12, equal_nan=True), key


@pytest.mark.parametrize("shape", ["ragged", "c5-like", "norel"])
def test_update_p_pam_cache_stays_on_the_from_scratch_chain(oracle, shape):
    """update_p's PAM cache (MOIRE_P_CACHE_ON, opt-in): cells that do not hold the proposed allele
    re-run only the mixture on cached P(any missing) vectors.  The cached vectors come from
    frequencies that were rescaled since (equal to the last bits), so the terms differ from a
    from-scratch evaluation by rounding x the cell's conditioning.  On a well-conditioned shape
    (ragged A_j, COI ~ 3) the run is the cache-off run: identical integer state and counters,
    floating state to 1e-9.  On COI-30 data (EGF cells, cooperative cells, coverage probabilities
    down to 1e-12) the terms are held to the oracle's from-scratch values with the bound that
    conditioning allows: 99 % of the cells within 1e-9, every cell within 1e-4, and a from-scratch
    re-evaluation on the device (eval_all) agrees with the oracle to 1e-9 again."""
    from moire_b200.data import pack_data
    from moire_b200.engine import DEFAULT_MODEL, Engine
    from moire_b200.simulate import simulate_data
    rng = np.random.default_rng(3)
    if shape == "ragged":
        L, N = 13, 157
        num_alleles = rng.integers(2, 41, size=L)
        sim = simulate_data(mean_coi=3.0, num_samples=N, epsilon_pos=.02, epsilon_neg=.05,
                            locus_freq_alphas=[np.full(int(a), .5) for a in num_alleles], seed=4)
        pk = pack_data(sim["data"], rng.random((L, N)) < .15)
        model = dict(DEFAULT_MODEL, max_coi=12, burnin=100)
    else:
        sim = simulate_data(sample_cois=rng.integers(1, 31, size=120), epsilon_pos=.01, epsilon_neg=.1,
                            locus_freq_alphas=[np.ones(30)] * 9, internal_relatedness_alpha=1,
                            internal_relatedness_beta=1, seed=8)
        pk = pack_data(sim["data"], sim["is_missing"])
        model = dict(DEFAULT_MODEL, max_coi=40, burnin=100, allow_relatedness=shape != "norel")
    allow = int(model["allow_relatedness"])
    temps = np.array([1.0, .6, .2])
    runs = {}
    for pipeline in ("flat-cache", "flat+cache"):
        eng = Engine(pk, temps, seed=21, pipeline=pipeline, **model)
        assert not eng.init_chains()["still_ill"].any()
        for it in range(5):
            eng.step(it)
        states = [eng.dump_state(c) for c in range(len(temps))]
        diags = [eng.diagnostics(c) for c in range(len(temps))]
        if pipeline == "flat+cache":
            worst = 0.0
            for c, st in enumerate(states):
                check_invariants(eng, pk, c)
                t, o = oracle.eval_cells(64, pk.num_alleles, pk.obs_mask, pk.is_missing, st["latent"],
                                         st["m"], st["r"], st["eps_neg"], st["eps_pos"],
                                         st["p"][:, :pk.max_alleles], allow)
                err = np.abs(st["cell_t"] - t) / np.maximum(1.0, np.abs(t))
                assert (err < 1e-9).mean() > .99 and err.max() < 1e-4, (err.max(), (err < 1e-9).mean())
                assert np.array_equal(st["cell_o"], o) or np.allclose(st["cell_o"], o, rtol=1e-12)
                worst = max(worst, float(err.max()))
            print(f"PAM cache, {shape}: largest relative difference of a cached term {worst:.3g}")
            eng.eval_all()
            for c in range(len(temps)):
                check_invariants(eng, pk, c, oracle, allow=allow, ncheck=2000, seed=c)
        runs[pipeline] = (states, diags)
        eng.close()
    if shape != "ragged":
        return
    (sa_all, da_all), (sb_all, db_all) = runs["flat-cache"], runs["flat+cache"]
    for sa, sb, da, db in zip(sa_all, sb_all, da_all, db_all):
        for a, b in ((sa, sb), (da, db)):
            for key in a:
                x, y = np.asarray(a[key]), np.asarray(b[key])
                if x.dtype.kind in "iu":
                    assert np.array_equal(x, y), key
                else:
                    assert np.allclose(x, y, rtol=1e-9, atol=1e-12, equal_nan=True), key


def test_step_graph_replays_the_plain_launch_sequence(monkeypatch):
    """moire_engine_step replays a captured CUDA graph; with MOIRE_B200_NO_GRAPH it issues the same
    kernels one by one.  Same state, bit for bit, and the same launch count."""
    from moire_b200.engine import Engine
    from moire_b200.workloads import make_workload
    wl = make_workload("c3")
    temps = wl["temps"][:3]
    outs = []
    for plain in (False, True):
        if plain:
            monkeypatch.setenv("MOIRE_B200_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("MOIRE_B200_NO_GRAPH", raising=False)
        eng = Engine(wl["packed"], temps, seed=2, burnin=100, **wl["model"])
        eng.init_chains()
        _, l0 = eng.counters()
        for it in range(4):
            eng.step(it)
        _, l1 = eng.counters()
        outs.append(([eng.dump_state(c) for c in range(3)], eng.scalars(), l1 - l0, eng.diagnostics_all(),
                     [eng.diagnostics(c) for c in range(3)]))
        eng.close()
    (sa, (la, pa), na, da, dsa), (sb, (lb, pb), nb, _, _) = outs
    assert np.array_equal(la, lb) and np.array_equal(pa, pb)
    assert na == nb + 4                      # plus the one-thread kernel that stores the iteration
    for x, y in zip(sa, sb):
        for key in x:
            assert np.array_equal(np.asarray(x[key]), np.asarray(y[key]), equal_nan=True), key
    # the bulk diagnostics read is the per-chain read, chain by chain
    for bulk, one in zip(da, dsa):
        for key in one:
            assert np.array_equal(np.asarray(bulk[key]), np.asarray(one[key])), key


def test_many_samples_need_no_shared_memory_budget():
    """9000 samples per locus: beyond what the fused update_p could stage in shared memory; the flat
    pipelines have no such bound."""
    from moire_b200.data import pack_data
    from moire_b200.engine import DEFAULT_MODEL, Engine
    from moire_b200.simulate import simulate_data
    sim = simulate_data(mean_coi=2.5, num_samples=9000, epsilon_pos=.01, epsilon_neg=.05,
                        locus_freq_alphas=[np.full(6, 1.0)] * 8, seed=8)
    pk = pack_data(sim["data"], sim["is_missing"])
    model = dict(DEFAULT_MODEL, max_coi=10, burnin=50)
    eng = Engine(pk, np.array([1.0, .5]), seed=4, **model)
    assert not eng.init_chains()["still_ill"].any()
    for it in range(3):
        eng.step(it)
    check_invariants(eng, pk, 0)
    check_invariants(eng, pk, 1)
    eng.close()


def test_simulate_data_on_the_device_follows_the_generative_model():
    """Row f2: the device generator against the NumPy restatement of R/simulator.R:154-212 on the
    same frequencies, COIs and relatedness -- two independent random streams, so the comparison is
    distributional: per-allele observation rates, strains per cell, missingness, and the exact
    structural facts (no bits above A_j, is_missing == nothing observed, present alleles <= COI)."""
    from moire_b200.data import pack_data
    from moire_b200.simulate import simulate_data, simulate_data_gpu
    rng = np.random.default_rng(7)
    alphas = [np.ones(a) for a in (2, 5, 12, 30, 64)]
    freqs = [rng.dirichlet(a) for a in alphas]
    N = 6000
    coi = rng.integers(1, 9, size=N)
    rel = rng.beta(.5, 2, size=N)
    kw = dict(sample_cois=coi, allele_freqs=freqs, internal_relatedness=rel, epsilon_pos=.4, epsilon_neg=.5,
              missingness=.07)
    g = simulate_data_gpu(seed=5, **kw)
    c = simulate_data(seed=5, **kw)
    pk, cp = g["packed"], pack_data(c["data"], c["is_missing"])
    assert pk.obs_mask.shape == (N, 5) and np.array_equal(pk.num_alleles, cp.num_alleles)
    assert np.array_equal(pk.is_missing != 0, pk.obs_mask == 0)
    pop = np.vectorize(lambda x: bin(int(x)).count("1"))
    present = pop(g["true_mask"])
    assert (present >= 1).all() and (present <= coi[:, None]).all()
    for j, A in enumerate(pk.num_alleles):
        assert A == 64 or not (pk.obs_mask[:, j] >> np.uint64(A)).any()
        bits_g = ((pk.obs_mask[:, j][:, None] >> np.arange(A, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(float)
        bits_c = ((cp.obs_mask[:, j][:, None] >> np.arange(A, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(float)
        se = np.sqrt((bits_c.var(0) + bits_g.var(0)) / N) + 1e-9
        assert (np.abs(bits_g.mean(0) - bits_c.mean(0)) < 5 * se).all(), j      # per-allele observation rate
        truth_c = (np.asarray(c["true_genotypes"][j]) > 0).sum(1)
        assert abs(present[:, j].mean() - truth_c.mean()) < 5 * np.sqrt(2 * truth_c.var() / N)
    assert abs(pk.is_missing.mean() - cp.is_missing.mean()) < 0.01
    # the generator's output drives the engine like any other data
    again = simulate_data_gpu(seed=5, **kw)["packed"]
    assert np.array_equal(again.obs_mask, pk.obs_mask)          # same seed, same data
