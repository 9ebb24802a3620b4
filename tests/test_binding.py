"""The R binding of this repository -- bindings/r/main_b200.cpp, the file a maintainer drops in as
moire's src/main.cpp -- compiled with the reference's own argument parsers (Parameters,
GenotypingData) and run without R against a functional Rcpp stand-in (oracle/shim_r, oracle/Makefile
target `binding`).  The harness builds the `args` list exactly as R/mcmc.R hands it over and calls
`run_mcmc(args)` as RcppExports would.

On the CPU the driver runs in ladder-only mode (MOIRE_B200_DEVICE=-1: nothing is evaluated, every
field of the result list must still come back with the reference's names, nesting and lengths,
src/main.cpp:138-189); on the GPU the same compiled binding must return what
`moire_b200.mcmc.run_mcmc` returns for the same seed."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libmoire_binding.so")

FIELDS = ["llik_burnin", "llik_sample", "prior_burnin", "prior_sample", "posterior_burnin",
          "posterior_sample", "data_llik", "coi", "lam_coi", "allele_freqs", "eps_neg", "eps_pos",
          "relatedness", "latent_genotypes", "observed_coi", "swap_store", "swap_acceptances",
          "swap_barriers", "temp_gradient", "acceptance_rates", "sampling_variances",
          "max_runtime_reached", "total_runtime"]     # src/main.cpp:164-186
SCALARS = ["verbose", "simple_verbose", "use_message", "allow_relatedness", "chain_number", "thin", "burnin",
           "samples", "pt_num_threads", "adapt_temp", "pre_adapt_steps", "temp_adapt_steps",
           "max_initialization_tries", "record_latent_genotypes", "max_runtime", "max_coi", "mean_coi_shape",
           "mean_coi_scale", "max_eps_pos", "eps_pos_alpha", "eps_pos_beta", "max_eps_neg", "eps_neg_alpha",
           "eps_neg_beta", "r_alpha", "r_beta"]
DEFAULTS = dict(verbose=0, simple_verbose=0, use_message=0, allow_relatedness=1, chain_number=1, thin=1,
                burnin=20, samples=10, pt_num_threads=1, adapt_temp=1, pre_adapt_steps=5, temp_adapt_steps=4,
                max_initialization_tries=100, record_latent_genotypes=0, max_runtime=1e9, max_coi=20,
                mean_coi_shape=.1, mean_coi_scale=10, max_eps_pos=2, eps_pos_alpha=1, eps_pos_beta=1,
                max_eps_neg=2, eps_neg_alpha=1, eps_neg_beta=1, r_alpha=1, r_beta=1)


LIB_W128 = os.path.join(ROOT, "oracle", "_ref", "libmoire_binding_w128.so")   # linked to libmoire_b200_w128.so


class Binding:
    def __init__(self, lib=LIB):
        if not os.path.exists(lib):
            pytest.skip(f"{lib} not built (needs /root/reference: make -C oracle binding)")
        self.lib = C.CDLL(lib)
        self.lib.binding_error.restype = C.c_char_p
        self.lib.binding_result_names.restype = C.c_char_p
        self.lib.binding_node_names.restype = C.c_char_p

    def run(self, sim, pt_chains, **kw):
        prm = dict(DEFAULTS)
        prm.update(kw)
        data = [np.asarray(m) for m in sim["data"]]
        L, N = len(data), data[0].shape[0]
        na = np.array([m.shape[1] for m in data], np.int32)
        obs = np.ascontiguousarray(np.concatenate([m.astype(np.uint8).ravel() for m in data]))
        mis = np.ascontiguousarray(np.asarray(sim["is_missing"], np.uint8))          # [L][N]
        sc = np.array([float(prm[k]) for k in SCALARS])
        pt = np.ascontiguousarray(pt_chains, dtype=np.float64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = self.lib.binding_run(L, N, p(na), p(obs), p(mis), p(sc), p(pt), len(pt))
        assert rc == 0, self.lib.binding_error().decode()
        return [n for n in self.lib.binding_result_names().decode().split(",") if n]

    def node(self, *path):
        pa = np.array(path, np.int32)
        t, n = C.c_int(), C.c_int()
        pp = pa.ctypes.data_as(C.c_void_p) if len(path) else None
        assert self.lib.binding_node(pp, len(path), C.byref(t), C.byref(n), None, 0) == 0, path
        vals = np.zeros(max(n.value, 1))
        self.lib.binding_node(pp, len(path), C.byref(t), C.byref(n), vals.ctypes.data_as(C.c_void_p), n.value)
        return t.value, n.value, vals[:n.value]

    def names(self, *path):
        pa = np.array(path, np.int32)
        return [x for x in self.lib.binding_node_names(pa.ctypes.data_as(C.c_void_p), len(path)).decode().split(",") if x]


def small_sim(N=12, seed=3):
    from moire_b200.simulate import simulate_data
    return simulate_data(mean_coi=3, num_samples=N, epsilon_pos=.02, epsilon_neg=.1,
                         locus_freq_alphas=[np.ones(4)] * 3 + [np.ones(7)] * 2, missingness=.1, seed=seed)


def test_binding_returns_the_reference_result_list_in_ladder_only_mode(monkeypatch):
    monkeypatch.setenv("MOIRE_B200_DEVICE", "-1")
    monkeypatch.setenv("MOIRE_B200_SEED", "5")
    b = Binding()
    sim = small_sim()
    N, L, T = 12, 5, 4
    names = b.run(sim, np.linspace(1, 0, T), thin=2, record_latent_genotypes=1)
    assert names == FIELDS
    idx = {n: i for i, n in enumerate(names)}
    for k in ("llik_burnin", "prior_burnin", "posterior_burnin"):
        assert b.node(idx[k])[:2] == (14, 20)
    for k in ("llik_sample", "prior_sample", "posterior_sample"):
        assert b.node(idx[k])[:2] == (14, 10)
    assert b.node(idx["swap_store"])[:2] == (13, 30)
    assert b.node(idx["temp_gradient"])[1] == T and b.node(idx["swap_acceptances"])[1] == T - 1
    assert b.node(idx["swap_barriers"])[1] == T - 1
    t, n, v = b.node(idx["observed_coi"])
    obs_coi = np.max([np.asarray(m).sum(1) for m in sim["data"]], axis=0)
    assert (t, n) == (13, N) and np.array_equal(v, obs_coi)            # GenotypingData parsed the list
    # per-sample and per-locus nesting (no draws are stored before the first swap at position 0 /
    # adaptation in ladder-only mode with constant scalars; the containers must still be there)
    for k in ("coi", "eps_neg", "eps_pos", "relatedness", "data_llik", "latent_genotypes"):
        assert b.node(idx[k])[:2] == (19, N), k
    assert b.node(idx["allele_freqs"])[:2] == (19, L)
    assert b.node(idx["latent_genotypes"], 0)[:2] == (19, L)
    D = b.node(idx["lam_coi"])[1]
    assert b.node(idx["coi"], 0)[1] == D and b.node(idx["allele_freqs"], 0)[1] == D
    assert b.node(idx["max_runtime_reached"])[0] == 10 and b.node(idx["max_runtime_reached"])[2][0] == 0
    assert b.node(idx["total_runtime"])[0] == 14
    # max_runtime = 0 stops at once, like the reference's `elapsed < max_runtime`
    b.run(sim, np.linspace(1, 0, T), max_runtime=0)
    assert b.node(idx["llik_burnin"])[1] == 0 and b.node(idx["max_runtime_reached"])[2][0] == 1


@pytest.mark.gpu
def test_binding_matches_python_run_mcmc_on_the_gpu(monkeypatch):
    from moire_b200.mcmc import run_mcmc
    monkeypatch.setenv("MOIRE_B200_DEVICE", "0")
    monkeypatch.setenv("MOIRE_B200_SEED", "9")
    b = Binding()
    sim = small_sim(N=25, seed=8)
    T = 3
    names = b.run(sim, np.linspace(1, 0, T), burnin=40, samples=20, thin=2, record_latent_genotypes=1)
    idx = {n: i for i, n in enumerate(names)}
    # the reference's Parameters holds its reals as float (src/parameters.h:31-47): the binding hands
    # the engine 0.1f for mean_coi_shape, so does this call
    f32 = lambda x: float(np.float32(x))
    ch = run_mcmc(sim, sim["is_missing"], burnin=40, samples_per_chain=20, thin=2, pt_chains=np.linspace(1, 0, T),
                  max_coi=20, pre_adapt_steps=5, temp_adapt_steps=4, max_initialization_tries=100,
                  mean_coi_shape=f32(DEFAULTS["mean_coi_shape"]), mean_coi_scale=f32(DEFAULTS["mean_coi_scale"]),
                  record_latent_genotypes=True, verbose=False, seed=9)["chains"][0]
    for k in ("llik_burnin", "llik_sample", "prior_sample", "posterior_burnin", "lam_coi", "swap_store",
              "temp_gradient", "swap_barriers", "swap_acceptances", "observed_coi"):
        assert np.array_equal(b.node(idx[k])[2], np.asarray(ch[k], float)), k
    D = len(ch["lam_coi"])
    assert D > 0
    for i in (0, 7, 24):
        for k in ("coi", "eps_neg", "relatedness", "data_llik"):
            assert np.array_equal(b.node(idx[k], i)[2], np.asarray(ch[k][i], float)), (k, i)
    for j in (0, 4):
        for d in (0, D - 1):
            assert np.array_equal(b.node(idx["allele_freqs"], j, d)[2], ch["allele_freqs"][j][d])
            assert np.array_equal(b.node(idx["latent_genotypes"], 3, j, d)[2],
                                  np.asarray(ch["latent_genotypes"][3][j][d], float))
    assert b.names(idx["acceptance_rates"], 0) == ["allele_freq_accept", "coi_accept", "eps_neg_accept",
                                                   "eps_pos_accept", "r_accept", "m_r_accept",
                                                   "full_sample_accept", "mean_coi_accept"]
    assert b.node(idx["acceptance_rates"])[1] == T and b.node(idx["sampling_variances"])[1] == T
    assert np.array_equal(b.node(idx["acceptance_rates"], 1, 1)[2], ch["acceptance_rates"][1]["coi_accept"])
    assert np.array_equal(b.node(idx["sampling_variances"], 2, 0, 3)[2],
                          ch["sampling_variances"][2]["allele_freq_var"][3])


@pytest.mark.gpu
def test_binding_linked_to_the_two_word_library_runs_loci_past_64_alleles(monkeypatch):
    """The same compiled binding linked to libmoire_b200_w128.so (it reads the mask width from the
    library): a 100-allele locus goes through GenotypingData -> two-word masks -> the engine and comes
    back as the reference's result list, equal to what run_mcmc returns for the seed; linked to the
    one-word library the same call stops with a message instead."""
    from moire_b200.mcmc import run_mcmc
    from moire_b200.simulate import simulate_data
    monkeypatch.setenv("MOIRE_B200_DEVICE", "0")
    monkeypatch.setenv("MOIRE_B200_SEED", "4")
    sim = simulate_data(mean_coi=3, num_samples=15, epsilon_pos=.02, epsilon_neg=.1,
                        locus_freq_alphas=[np.ones(100), np.ones(6), np.ones(70)], missingness=.1, seed=2)
    T = 3
    b = Binding(LIB_W128)
    names = b.run(sim, np.linspace(1, 0, T), burnin=30, samples=12, record_latent_genotypes=1)
    assert names == FIELDS
    idx = {n: i for i, n in enumerate(names)}
    f32 = lambda x: float(np.float32(x))
    ch = run_mcmc(sim, sim["is_missing"], burnin=30, samples_per_chain=12, pt_chains=np.linspace(1, 0, T),
                  max_coi=20, pre_adapt_steps=5, temp_adapt_steps=4, max_initialization_tries=100,
                  mean_coi_shape=f32(DEFAULTS["mean_coi_shape"]), mean_coi_scale=f32(DEFAULTS["mean_coi_scale"]),
                  record_latent_genotypes=True, verbose=False, seed=4)["chains"][0]
    for k in ("llik_burnin", "llik_sample", "lam_coi", "swap_store", "observed_coi"):
        assert np.array_equal(b.node(idx[k])[2], np.asarray(ch[k], float)), k
    D = len(ch["lam_coi"])
    assert D > 0
    for j, A in enumerate((100, 6, 70)):
        t, n, v = b.node(idx["allele_freqs"], j, D - 1)
        assert n == A and np.array_equal(v, ch["allele_freqs"][j][D - 1])
        for i in (0, 9):
            g = b.node(idx["latent_genotypes"], i, j, D - 1)[2]
            assert np.array_equal(g, np.asarray(ch["latent_genotypes"][i][j][D - 1], float)) and (g < A).all()
    assert any((b.node(idx["latent_genotypes"], i, 0, d)[2] >= 64).any() for i in range(15) for d in range(D))
    narrow = Binding()
    data = [np.asarray(m) for m in sim["data"]]
    na = np.array([m.shape[1] for m in data], np.int32)
    obs = np.ascontiguousarray(np.concatenate([m.astype(np.uint8).ravel() for m in data]))
    mis = np.ascontiguousarray(np.asarray(sim["is_missing"], np.uint8))
    sc = np.array([float(DEFAULTS[k]) for k in SCALARS])
    pt = np.linspace(1, 0, T)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert narrow.lib.binding_run(3, 15, p(na), p(obs), p(mis), p(sc), p(pt), T) != 0
    assert b"at most 64 alleles" in narrow.lib.binding_error()
