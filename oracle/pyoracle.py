"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes bindings for the two checkers.

* ``Oracle``  -> oracle/liboracle.so, the plain-C restatement (travels as source, built by
  ``__graft_entry__.build()`` / ``make -C oracle oracle``).
* ``Ref``     -> oracle/_ref/libmoire_ref{32,64}.so, the UNMODIFIED reference compiled in place
  against oracle/shim (built only where /root/reference exists; the .so travels to the GPU box).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs
may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

c_double_p = C.POINTER(C.c_double)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _i32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))


def build_oracle():
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)


def ref_available(prec=64):
    return os.path.exists(os.path.join(HERE, "_ref", f"libmoire_ref{prec}.so"))


class Oracle:
    """Plain-C restatement.  ``prec`` = 32 or 64 on every call."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = self.lib = C.CDLL(path)
        d, i, vp = C.c_double, C.c_int, C.c_void_p
        L.orc_pam_scalar.restype = d
        L.orc_pam_scalar.argtypes = [i, vp, i, i]
        L.orc_pam_vectorized.argtypes = [i, vp, i, i, vp]
        L.orc_transmission.restype = d
        L.orc_transmission.argtypes = [i, vp, i, vp, i, i, d, i]
        L.orc_observation.restype = d
        L.orc_observation.argtypes = [i, vp, i, vp, i, d, d, d, d]
        L.orc_latent_log_prob.restype = d
        L.orc_latent_log_prob.argtypes = [i, i, i, i, i, i, d, d]
        L.orc_constrained.argtypes = [i, d, d, d, d, vp, vp]
        L.orc_salt.argtypes = [i, vp, i, i, d, vp, vp]
        L.orc_logit.restype = d
        L.orc_logit.argtypes = [i, d]
        L.orc_log_pq.argtypes = [i, d, vp, vp]
        L.orc_logit_sum.restype = d
        L.orc_logit_sum.argtypes = [i, vp, i]
        L.orc_logit_scale.argtypes = [i, vp, i, d, vp]
        L.orc_expit.restype = d
        L.orc_expit.argtypes = [i, d]
        L.orc_log_sum_exp.restype = d
        L.orc_log_sum_exp.argtypes = [i, vp, i]
        L.orc_dztpois.restype = d
        L.orc_dztpois.argtypes = [i, i, d]
        L.orc_dbinom.restype = d
        L.orc_dbinom.argtypes = [i, i, i, d]
        L.orc_beta_log_prior.restype = d
        L.orc_beta_log_prior.argtypes = [i, d, d, d]
        L.orc_gamma_log_prior.restype = d
        L.orc_gamma_log_prior.argtypes = [i, d, d, d]
        L.orc_combinations.restype = C.c_long
        L.orc_combinations.argtypes = [i, i, vp, C.c_long, vp, vp]
        L.orc_eval_cells.argtypes = [i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, d, d,
                                     vp, vp]
        L.orc_eval_cells_wide.argtypes = [i, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, d, d,
                                          vp, vp]

    # a6
    def combinations(self, n, r, max_rows=5000):
        out = np.zeros((max_rows, max(r, 1)), dtype=np.int8)
        nc = C.c_ulong(0)
        gen = C.c_ulong(0)
        rows = self.lib.orc_combinations(n, r, _p(out), max_rows, C.byref(nc), C.byref(gen))
        return out[:rows, :r].copy(), int(nc.value), int(gen.value)

    def pam_scalar(self, prec, q, n):
        q = _f64(q)
        return self.lib.orc_pam_scalar(prec, _p(q), len(q), n)

    def pam_vectorized(self, prec, q, n_max):
        q = _f64(q)
        out = np.zeros(n_max)
        self.lib.orc_pam_vectorized(prec, _p(q), len(q), n_max, _p(out))
        return out

    def transmission(self, prec, idx, p, coi, r, allow_relatedness):
        idx, p = _i32(idx), _f64(p)
        return self.lib.orc_transmission(prec, _p(idx), len(idx), _p(p), len(p), coi, r,
                                         int(allow_relatedness))

    def observation(self, prec, idx, obs, eps_neg, eps_pos, max_eps_neg=2.0, max_eps_pos=2.0):
        idx, obs = _i32(idx), _i32(obs)
        return self.lib.orc_observation(prec, _p(idx), len(idx), _p(obs), len(obs), eps_neg,
                                        eps_pos, max_eps_neg, max_eps_pos)

    def stepper_coi_delta(self, seed, n):
        out = np.zeros(n, np.int32)
        f = self.lib.orc_stp_coi_delta_many
        f.restype = None
        f.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
        f(seed, n, _p(out))
        return out

    def stepper_sample_latent(self, seed, n, obs_mask, A, coi, eps_pos, eps_neg):
        """n draws of the engine-semantics latent-genotype proposal: (masks uint64[n], log_prob[n])."""
        masks, lp = np.zeros(n, np.uint64), np.zeros(n)
        f = self.lib.orc_stp_sample_latent_many
        f.restype = None
        f.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_double,
                      C.c_void_p, C.c_void_p]
        f(seed, n, int(obs_mask), A, coi, eps_pos, eps_neg, _p(masks), _p(lp))
        return masks, lp

    def latent_log_prob(self, prec, A, obs_pos, fn, fp, coi, eps_pos, eps_neg):
        return self.lib.orc_latent_log_prob(prec, A, obs_pos, fn, fp, coi, eps_pos, eps_neg)

    def constrained(self, prec, z, curr, lo, hi):
        a, b = C.c_double(), C.c_double()
        self.lib.orc_constrained(prec, z, curr, lo, hi, C.byref(a), C.byref(b))
        return a.value, b.value

    def salt(self, prec, p, idx, logit_prop):
        p = _f64(p)
        out = np.zeros(len(p))
        adj = C.c_double()
        self.lib.orc_salt(prec, _p(p), len(p), idx, logit_prop, _p(out), C.byref(adj))
        return out, adj.value

    def logit(self, prec, x):
        return self.lib.orc_logit(prec, x)

    def log_pq(self, prec, x):
        a, b = C.c_double(), C.c_double()
        self.lib.orc_log_pq(prec, x, C.byref(a), C.byref(b))
        return a.value, b.value

    def logit_sum(self, prec, x):
        x = _f64(x)
        return self.lib.orc_logit_sum(prec, _p(x), len(x))

    def logit_scale(self, prec, x, scale):
        x = _f64(x)
        out = np.zeros(len(x))
        self.lib.orc_logit_scale(prec, _p(x), len(x), scale, _p(out))
        return out

    def expit(self, prec, x):
        return self.lib.orc_expit(prec, x)

    def log_sum_exp(self, prec, x):
        x = _f64(x)
        return self.lib.orc_log_sum_exp(prec, _p(x), len(x))

    def dztpois(self, prec, x, lam):
        return self.lib.orc_dztpois(prec, x, lam)

    def dbinom(self, prec, x, size, prob):
        return self.lib.orc_dbinom(prec, x, size, prob)

    def beta_log_prior(self, prec, x, a, b):
        return self.lib.orc_beta_log_prior(prec, x, a, b)

    def gamma_log_prior(self, prec, x, shape, scale):
        return self.lib.orc_gamma_log_prior(prec, x, shape, scale)

    def eval_cells(self, prec, num_alleles, obs, is_missing, latent, m, r, eps_neg, eps_pos, p,
                   allow_relatedness, max_eps_neg=2.0, max_eps_pos=2.0):
        """obs/latent/is_missing: [N][L] (masks of W words: [N][L][W]); p: [L][a_pad].  Returns (T, O)
        each [N][L]."""
        obs = np.asarray(obs)
        N, L = obs.shape[:2]
        na = _i32(num_alleles)
        obs = np.ascontiguousarray(obs, dtype=np.uint64)
        lat = np.ascontiguousarray(latent, dtype=np.uint64)
        mis = np.ascontiguousarray(is_missing, dtype=np.uint8)
        m = _i32(m)
        r, en, ep, p = _f64(r), _f64(eps_neg), _f64(eps_pos), _f64(p)
        t = np.zeros((N, L))
        o = np.zeros((N, L))
        if obs.ndim == 3:
            assert lat.shape == obs.shape, (lat.shape, obs.shape)
            self.lib.orc_eval_cells_wide(prec, N, L, obs.shape[2], _p(na), _p(obs), _p(mis), _p(lat),
                                         _p(m), _p(r), _p(en), _p(ep), _p(p), p.shape[1],
                                         int(allow_relatedness), max_eps_neg, max_eps_pos, _p(t), _p(o))
            return t, o
        self.lib.orc_eval_cells(prec, N, L, _p(na), _p(obs), _p(mis), _p(lat), _p(m), _p(r),
                                _p(en), _p(ep), _p(p), p.shape[1], int(allow_relatedness),
                                max_eps_neg, max_eps_pos, _p(t), _p(o))
        return t, o


PARAM_ORDER = ["allow_relatedness", "thin", "burnin", "samples", "pt_num_threads", "adapt_temp",
               "pre_adapt_steps", "temp_adapt_steps", "max_initialization_tries",
               "record_latent_genotypes", "max_coi", "mean_coi_shape", "mean_coi_scale",
               "max_eps_pos", "eps_pos_alpha", "eps_pos_beta", "max_eps_neg", "eps_neg_alpha",
               "eps_neg_beta", "r_alpha", "r_beta"]

# defaults of moire::run_mcmc (R/mcmc.R:76-106)
DEFAULT_PARAMS = dict(allow_relatedness=1, thin=1, burnin=10000, samples=1000, pt_num_threads=1,
                      adapt_temp=1, pre_adapt_steps=25, temp_adapt_steps=25,
                      max_initialization_tries=10000, record_latent_genotypes=0, max_coi=40,
                      mean_coi_shape=0.1, mean_coi_scale=10.0, max_eps_pos=2.0, eps_pos_alpha=1.0,
                      eps_pos_beta=1.0, max_eps_neg=2.0, eps_neg_alpha=1.0, eps_neg_beta=1.0,
                      r_alpha=1.0, r_beta=1.0)

UPDATE_KINDS = dict(eps_neg=0, eps_pos=1, p=2, m=3, r=4, eff_coi=5, samples=6, mean_coi=7)


class Ref:
    """The compiled reference (float build: prec=32; mechanical double variant: prec=64).

    GenotypingData lives in static members of the reference, so one process holds ONE dataset
    per loaded library at a time (set_data replaces it)."""

    def __init__(self, prec=64):
        path = os.path.join(HERE, "_ref", f"libmoire_ref{prec}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with `make -C oracle ref` where "
                                    "/root/reference exists)")
        L = self.lib = C.CDLL(path)
        self.prec = prec
        assert L.ref_real_bytes() * 8 == prec
        d, i, vp, u64, lg = C.c_double, C.c_int, C.c_void_p, C.c_uint64, C.c_long
        L.ref_seed_runif.argtypes = [u64]
        L.ref_draw_log_runif.restype = d
        L.ref_combinations.restype = lg
        L.ref_combinations.argtypes = [i, i, vp, lg, vp, vp]
        L.ref_pam_scalar.restype = d
        L.ref_pam_scalar.argtypes = [vp, i, i]
        L.ref_pam_vectorized.argtypes = [vp, i, i, vp]
        L.ref_set_data.argtypes = [i, i, vp, vp, vp]
        L.ref_set_params.argtypes = [vp, vp, i]
        L.ref_chain_new.restype = vp
        L.ref_chain_new.argtypes = [d, u64]
        L.ref_chain_free.argtypes = [vp]
        L.ref_chain_reseed.argtypes = [vp, u64]
        L.ref_chain_set_temp.argtypes = [vp, d]
        L.ref_chain_set_state.argtypes = [vp, vp, vp, vp, vp, vp, i, vp, vp, d]
        L.ref_chain_get_state.argtypes = [vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp]
        L.ref_chain_get_tuning.argtypes = [vp, vp, vp]
        L.ref_chain_update.argtypes = [vp, i, i]
        L.ref_chain_transmission.restype = d
        L.ref_chain_transmission.argtypes = [vp, vp, i, vp, i, i, d]
        L.ref_chain_observation.restype = d
        L.ref_chain_observation.argtypes = [vp, vp, i, vp, i, d, d]
        L.ref_sample_latent_genotype.restype = i
        L.ref_sample_latent_genotype.argtypes = [u64, vp, i, i, d, d, vp, vp]
        L.ref_sample_constrained.argtypes = [u64, d, d, d, d, vp, vp, vp]
        L.ref_dztpois.restype = d
        L.ref_dztpois.argtypes = [i, d]
        L.ref_dbinom.restype = d
        L.ref_dbinom.argtypes = [i, i, d]
        L.ref_beta_log_prior.restype = d
        L.ref_beta_log_prior.argtypes = [d, d, d]
        L.ref_gamma_log_prior.restype = d
        L.ref_gamma_log_prior.argtypes = [d, d, d]
        L.ref_min_sampled.restype = d
        L.ref_logit.restype = d
        L.ref_logit.argtypes = [d]
        L.ref_log_pq.argtypes = [d, vp, vp]
        L.ref_logit_sum.restype = d
        L.ref_logit_sum.argtypes = [vp, i]
        L.ref_logit_scale.argtypes = [vp, i, d, vp]
        L.ref_expit.restype = d
        L.ref_expit.argtypes = [d]
        L.ref_salt_proposal.argtypes = [vp, i, i, d, vp, vp]
        L.ref_log_sum_exp.restype = d
        L.ref_log_sum_exp.argtypes = [vp, i]
        L.ref_mcmc_new.restype = vp
        L.ref_mcmc_free.argtypes = [vp]
        L.ref_mcmc_init_failed.argtypes = [vp]
        L.ref_mcmc_reseed.argtypes = [vp, u64]
        L.ref_mcmc_burnin.argtypes = [vp, i]
        L.ref_mcmc_sample.argtypes = [vp, i]
        L.ref_mcmc_finalize.argtypes = [vp]
        L.ref_mcmc_llik.restype = d
        L.ref_mcmc_llik.argtypes = [vp]
        L.ref_mcmc_prior.restype = d
        L.ref_mcmc_prior.argtypes = [vp]
        L.ref_mcmc_time_burnin.restype = d
        L.ref_mcmc_time_burnin.argtypes = [vp, i, i, i]
        L.ref_eval_count.restype = C.c_ulonglong
        L.ref_eval_count.argtypes = []
        L.ref_eval_count_reset.restype = None
        L.ref_eval_count_reset.argtypes = []
        L.ref_mcmc_summary.restype = i
        L.ref_mcmc_summary.argtypes = [vp, vp, vp, vp, vp, vp, i, vp]
        L.ref_mcmc_coi_draws.restype = i
        L.ref_mcmc_coi_draws.argtypes = [vp, vp, i]
        L.ref_mcmc_trace.restype = i
        L.ref_mcmc_trace.argtypes = [vp, vp, i, vp, i]
        L.ref_mcmc_swap_pass.argtypes = [vp, vp, vp, vp, i, i, i, vp, vp, vp, vp, vp, vp]
        L.ref_mcmc_adapt_temp.argtypes = [vp, vp, vp, lg, vp]
        L.ref_mcmc_get_ladder.argtypes = [vp, vp, vp, vp, vp]
        self.N = self.L = 0
        self.num_alleles = None

    # ---- pure pieces ----
    def combinations(self, n, r, max_rows=5000):
        out = np.zeros((max_rows, max(r, 1)), dtype=np.int8)
        nc, gen = C.c_ulong(0), C.c_ulong(0)
        rows = self.lib.ref_combinations(n, r, _p(out), max_rows, C.byref(nc), C.byref(gen))
        return out[:rows, :r].copy(), int(nc.value), int(gen.value)

    def pam_scalar(self, q, n):
        q = _f64(q)
        return self.lib.ref_pam_scalar(_p(q), len(q), n)

    def pam_vectorized(self, q, n_max):
        q = _f64(q)
        out = np.zeros(n_max)
        self.lib.ref_pam_vectorized(_p(q), len(q), n_max, _p(out))
        return out

    def sample_latent_genotype(self, seed, obs, coi, eps_pos, eps_neg):
        obs = _i32(obs)
        out = np.zeros(len(obs) + 1, dtype=np.int32)
        lp = C.c_double()
        k = self.lib.ref_sample_latent_genotype(seed, _p(obs), len(obs), coi, eps_pos, eps_neg,
                                                _p(out), C.byref(lp))
        return (out[:max(k, 0)].copy(), lp.value)

    def sample_coi_delta(self, seed, mean, n):
        out = np.zeros(n, np.int32)
        f = self.lib.ref_sample_coi_delta_many
        f.restype = None
        f.argtypes = [C.c_uint64, C.c_double, C.c_int, C.c_void_p]
        f(seed, mean, n, _p(out))
        return out

    def sample_constrained(self, seed, curr, sd, lo, hi):
        z, prop, adj = C.c_double(), C.c_double(), C.c_double()
        self.lib.ref_sample_constrained(seed, curr, sd, lo, hi, C.byref(z), C.byref(prop),
                                        C.byref(adj))
        return z.value, prop.value, adj.value

    def dztpois(self, x, lam):
        return self.lib.ref_dztpois(x, lam)

    def dbinom(self, x, size, prob):
        return self.lib.ref_dbinom(x, size, prob)

    def beta_log_prior(self, x, a, b):
        return self.lib.ref_beta_log_prior(x, a, b)

    def gamma_log_prior(self, x, shape, scale):
        return self.lib.ref_gamma_log_prior(x, shape, scale)

    def min_sampled(self):
        return self.lib.ref_min_sampled()

    def logit(self, x):
        return self.lib.ref_logit(x)

    def log_pq(self, x):
        a, b = C.c_double(), C.c_double()
        self.lib.ref_log_pq(x, C.byref(a), C.byref(b))
        return a.value, b.value

    def logit_sum(self, x):
        x = _f64(x)
        return self.lib.ref_logit_sum(_p(x), len(x))

    def logit_scale(self, x, scale):
        x = _f64(x)
        out = np.zeros(len(x))
        self.lib.ref_logit_scale(_p(x), len(x), scale, _p(out))
        return out

    def expit(self, x):
        return self.lib.ref_expit(x)

    def salt(self, p, idx, logit_prop):
        p = _f64(p)
        out = np.zeros(len(p))
        adj = C.c_double()
        self.lib.ref_salt_proposal(_p(p), len(p), idx, logit_prop, _p(out), C.byref(adj))
        return out, adj.value

    def log_sum_exp(self, x):
        x = _f64(x)
        return self.lib.ref_log_sum_exp(_p(x), len(x))

    # ---- data / params ----
    def set_data(self, num_alleles, obs_mask, is_missing):
        """obs_mask: uint64 [N][L]; is_missing: [N][L] (sample-major, like the engine)."""
        N, L = obs_mask.shape
        na = _i32(num_alleles)
        parts = []
        for j in range(L):
            bits = (obs_mask[:, j][:, None] >> np.arange(na[j], dtype=np.uint64)[None, :]) & 1
            parts.append(bits.astype(np.uint8).ravel())
        flat = np.ascontiguousarray(np.concatenate(parts))
        mis = np.ascontiguousarray(np.asarray(is_missing, dtype=np.uint8).T)  # -> [L][N]
        self.lib.ref_set_data(L, N, _p(na), _p(flat), _p(mis))
        self.N, self.L, self.num_alleles = N, L, na
        self.a_pad = int(na.max())

    def set_params(self, pt_chains=(1.0,), **kw):
        prm = dict(DEFAULT_PARAMS)
        prm.update(kw)
        v = _f64([prm[k] for k in PARAM_ORDER])
        t = _f64(pt_chains)
        self.lib.ref_set_params(_p(v), _p(t), len(t))
        self.params = prm

    # ---- one chain ----
    def chain_new(self, temp=1.0, seed=1):
        return self.lib.ref_chain_new(temp, seed)

    def chain_free(self, h):
        self.lib.ref_chain_free(h)

    def chain_reseed(self, h, seed):
        self.lib.ref_chain_reseed(h, seed)

    def chain_set_temp(self, h, t):
        self.lib.ref_chain_set_temp(h, t)

    def chain_set_state(self, h, st):
        m = _i32(st["m"])
        r, ep, en = _f64(st["r"]), _f64(st["eps_pos"]), _f64(st["eps_neg"])
        p = _f64(st["p"])
        lat = np.ascontiguousarray(st["latent"], dtype=np.uint64)
        adj = _f64(st["lg_adj"])
        self.lib.ref_chain_set_state(h, _p(m), _p(r), _p(ep), _p(en), _p(p), p.shape[1], _p(lat),
                                     _p(adj), float(st["mean_coi"]))

    def chain_get_state(self, h):
        N, L, A = self.N, self.L, self.a_pad
        st = dict(m=np.zeros(N, np.int32), r=np.zeros(N), eps_pos=np.zeros(N), eps_neg=np.zeros(N),
                  p=np.zeros((L, A)), latent=np.zeros((N, L), np.uint64), lg_adj=np.zeros((N, L)),
                  cell_llik=np.zeros((N, L)), priors=np.zeros((4, N)), scalars=np.zeros(7))
        self.lib.ref_chain_get_state(h, _p(st["m"]), _p(st["r"]), _p(st["eps_pos"]),
                                     _p(st["eps_neg"]), _p(st["p"]), A, _p(st["latent"]),
                                     _p(st["lg_adj"]), _p(st["cell_llik"]), _p(st["priors"]),
                                     _p(st["scalars"]))
        s = st.pop("scalars")
        st.update(llik=s[0], prior=s[1], temp=s[2], mean_coi=s[3], mean_coi_hyper_prior=s[4],
                  mean_coi_var=s[5], mean_coi_accept=s[6])
        return st

    def chain_get_tuning(self, h):
        acc = np.zeros((6, self.N), np.int32)
        sds = np.zeros((4, self.N))
        self.lib.ref_chain_get_tuning(h, _p(acc), _p(sds))
        return acc, sds

    def chain_get_p_tuning(self, h):
        """(p_accept, p_attempt, p_prop_var), each [L][A_max], of the SALT sampler."""
        A = int(self.num_alleles.max())
        acc, att = np.zeros((self.L, A), np.int32), np.zeros((self.L, A), np.int32)
        sd = np.zeros((self.L, A))
        self.lib.ref_chain_get_p_tuning(h, A, _p(acc), _p(att), _p(sd))
        return acc, att, sd

    def chain_update(self, h, kind, iteration):
        self.lib.ref_chain_update(h, UPDATE_KINDS[kind] if isinstance(kind, str) else kind,
                                  iteration)

    def chain_step(self, h, iteration):
        """All updates in the order MCMC::burnin runs them (src/mcmc.cpp:162-173)."""
        order = ["eps_neg", "eps_pos", "p", "m"]
        if self.params["allow_relatedness"]:
            order += ["r", "eff_coi"]
        order += ["samples", "mean_coi"]
        for k in order:
            self.chain_update(h, k, iteration)

    def chain_transmission(self, h, idx, p, coi, r):
        idx, p = _i32(idx), _f64(p)
        return self.lib.ref_chain_transmission(h, _p(idx), len(idx), _p(p), len(p), coi, r)

    def chain_observation(self, h, idx, obs, eps_neg, eps_pos):
        idx, obs = _i32(idx), _i32(obs)
        return self.lib.ref_chain_observation(h, _p(idx), len(idx), _p(obs), len(obs), eps_neg,
                                              eps_pos)

    # ---- MCMC driver ----
    def mcmc_new(self, seed=None):
        h = self.lib.ref_mcmc_new()
        if seed is not None:
            self.lib.ref_mcmc_reseed(h, seed)
            self.lib.ref_seed_runif(seed)
        return h

    def mcmc_free(self, h):
        self.lib.ref_mcmc_free(h)

    def mcmc_time_burnin(self, h, first_step, warmup, steps):
        """Wall seconds of `steps` MCMC::burnin steps after `warmup` untimed ones, driven the way
        src/main.cpp:54-70 drives them."""
        return float(self.lib.ref_mcmc_time_burnin(h, int(first_step), int(warmup), int(steps)))

    def eval_count(self):
        """Non-missing calculate_genotype_likelihood evaluations since eval_count_reset()."""
        return int(self.lib.ref_eval_count())

    def eval_count_reset(self):
        self.lib.ref_eval_count_reset()

    def omp_max_threads(self):
        return int(self.lib.ref_omp_max_threads())

    def seed_runif(self, seed):
        self.lib.ref_seed_runif(seed)

    def draw_log_runif(self):
        return float(self.lib.ref_draw_log_runif())

    def mcmc_swap_pass(self, h, chain_llik, chain_temp, swap_indices, even_swap, step, burnin,
                       swap_barriers, swap_acceptances, num_swaps):
        """MCMC::swap_chains (src/mcmc.cpp:190-240) on explicit inputs; returns the new ladder."""
        T = len(chain_llik)
        ll, tp = _f64(chain_llik), _f64(chain_temp)
        si = np.ascontiguousarray(swap_indices, dtype=np.int64)
        so, to, hot = np.zeros(T, np.int64), np.zeros(T), np.zeros(T, np.int32)
        sb = _f64(swap_barriers).copy()
        sa = np.ascontiguousarray(swap_acceptances, dtype=np.int64).copy()
        ns = C.c_long(int(num_swaps))
        self.lib.ref_mcmc_swap_pass(h, _p(ll), _p(tp), _p(si), int(even_swap), int(step),
                                    int(burnin), _p(so), _p(to), _p(hot), _p(sb), _p(sa),
                                    C.byref(ns))
        return dict(swap_indices=so, chain_temps=to, hot=hot, swap_barriers=sb,
                    swap_acceptances=sa, num_swaps=ns.value)

    def mcmc_adapt_temp(self, h, temp_gradient, swap_barriers, num_swaps):
        """MCMC::adapt_temp (src/mcmc.cpp:244-304) on explicit inputs; returns the new gradient."""
        out = np.zeros(len(temp_gradient))
        self.lib.ref_mcmc_adapt_temp(h, _p(_f64(temp_gradient)), _p(_f64(swap_barriers)),
                                     int(num_swaps), _p(out))
        return out

    def mcmc_summary(self, h):
        N, L, A = self.N, self.L, self.a_pad
        out = dict(m=np.zeros(N), r=np.zeros(N), eps_neg=np.zeros(N), eps_pos=np.zeros(N),
                   p=np.zeros((L, A)))
        mc = C.c_double()
        D = self.lib.ref_mcmc_summary(h, _p(out["m"]), _p(out["r"]), _p(out["eps_neg"]),
                                      _p(out["eps_pos"]), _p(out["p"]), A, C.byref(mc))
        out["mean_coi"] = mc.value
        out["draws"] = D
        return out

    def mcmc_coi_draws(self, h, max_draws):
        out = np.zeros((self.N, max_draws), np.int32)
        D = self.lib.ref_mcmc_coi_draws(h, _p(out), max_draws)
        return out[:, :D]


class _StpChain(C.Structure):
    _fields_ = ([("N", C.c_int), ("L", C.c_int), ("a_pad", C.c_int),
                 ("num_alleles", C.c_void_p), ("obs", C.c_void_p), ("is_missing", C.c_void_p),
                 ("allow_relatedness", C.c_int), ("burnin", C.c_int), ("max_coi", C.c_int)] +
                [(n, C.c_double) for n in ("mean_coi_shape", "mean_coi_scale", "max_eps_pos",
                                           "eps_pos_alpha", "eps_pos_beta", "max_eps_neg",
                                           "eps_neg_alpha", "eps_neg_beta", "r_alpha", "r_beta")] +
                [("seed", C.c_uint64), ("chain", C.c_int), ("temp", C.c_double)] +
                [(n, C.c_void_p) for n in ("m", "r", "eps_pos", "eps_neg", "p", "latent", "lg_adj",
                                           "cell_t", "cell_o", "prior_eps_neg", "prior_eps_pos",
                                           "prior_coi", "prior_r", "mean_coi", "hyper",
                                           "sd_eps_neg", "sd_eps_pos", "sd_r", "sd_mr",
                                           "acc_eps_neg", "acc_eps_pos", "acc_m", "acc_r",
                                           "acc_mr", "acc_samples", "p_sd", "p_acc", "p_att",
                                           "mean_coi_sd", "mean_coi_acc")] +
                [("evals", C.c_ulonglong), ("W", C.c_int)])


STATE_FIELDS = ("m", "r", "eps_pos", "eps_neg", "p", "latent", "lg_adj", "cell_t", "cell_o",
                "prior_eps_neg", "prior_eps_pos", "prior_coi", "prior_r")
TUNING_FIELDS = ("sd_eps_neg", "sd_eps_pos", "sd_r", "sd_mr", "acc_eps_neg", "acc_eps_pos",
                 "acc_m", "acc_r", "acc_mr", "acc_samples", "p_sd", "p_acc", "p_att")


class Stepper:
    """One chain advanced on the CPU with the engine's stream addressing (fp64).

    ``state``: arrays as returned by ``Engine.dump_state`` (p padded to the engine's a_pad);
    ``diag``: as returned by ``Engine.diagnostics`` (or None for a fresh chain)."""

    def __init__(self, oracle, packed, model, seed, chain, temp, state, diag=None):
        self.lib = oracle.lib
        self.lib.orc_step_update.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self.lib.orc_eval_priors.argtypes = [C.c_void_p]
        N, L = packed.num_samples, packed.num_loci
        A = state["p"].shape[1]
        a = self.a = {}
        a["num_alleles"] = _i32(packed.num_alleles)
        a["obs"] = np.ascontiguousarray(packed.obs_mask, dtype=np.uint64)
        a["is_missing"] = np.ascontiguousarray(packed.is_missing, dtype=np.uint8)
        a["m"] = _i32(state["m"]).copy()
        for k in ("r", "eps_pos", "eps_neg", "p", "lg_adj", "cell_t", "cell_o", "prior_eps_neg",
                  "prior_eps_pos", "prior_coi", "prior_r"):
            a[k] = _f64(state[k]).copy()
        a["latent"] = np.ascontiguousarray(state["latent"], dtype=np.uint64).copy()
        a["mean_coi"] = np.array([state["mean_coi"]], dtype=np.float64)
        a["hyper"] = np.array([state["mean_coi_hyper_prior"]], dtype=np.float64)
        if diag is None:
            diag = dict(eps_neg_var=np.ones(N), eps_pos_var=np.ones(N), r_var=np.ones(N),
                        m_r_var=np.ones(N), p_prop_var=np.ones((L, A)), mean_coi_var=1.0,
                        mean_coi_accept=0.0)
        z = lambda k, shape: _i32(diag[k]).copy() if k in diag else np.zeros(shape, np.int32)
        a["sd_eps_neg"] = _f64(diag["eps_neg_var"]).copy()
        a["sd_eps_pos"] = _f64(diag["eps_pos_var"]).copy()
        a["sd_r"] = _f64(diag["r_var"]).copy()
        a["sd_mr"] = _f64(diag["m_r_var"]).copy()
        a["acc_eps_neg"] = z("eps_neg_accept", N)
        a["acc_eps_pos"] = z("eps_pos_accept", N)
        a["acc_m"] = z("m_accept", N)
        a["acc_r"] = z("r_accept", N)
        a["acc_mr"] = z("m_r_accept", N)
        a["acc_samples"] = z("sample_accept", N)
        a["p_sd"] = _f64(diag["p_prop_var"]).copy()
        a["p_acc"] = z("p_accept", (L, A))
        a["p_att"] = z("p_attempt", (L, A))
        a["mean_coi_sd"] = np.array([diag["mean_coi_var"]], dtype=np.float64)
        a["mean_coi_acc"] = np.array([diag["mean_coi_accept"]], dtype=np.float64)
        s = self.s = _StpChain()
        s.N, s.L, s.a_pad = N, L, A
        s.allow_relatedness = int(bool(model["allow_relatedness"]))
        s.burnin, s.max_coi = int(model["burnin"]), int(model["max_coi"])
        for k in ("mean_coi_shape", "mean_coi_scale", "max_eps_pos", "eps_pos_alpha",
                  "eps_pos_beta", "max_eps_neg", "eps_neg_alpha", "eps_neg_beta", "r_alpha",
                  "r_beta"):
            setattr(s, k, float(model[k]))
        s.seed, s.chain, s.temp = seed, chain, float(temp)
        for k, v in a.items():
            setattr(s, k, v.ctypes.data)
        s.evals = 0
        s.W = 1 if a["obs"].ndim == 2 else int(a["obs"].shape[2])   # 64-bit words per allele set
        assert a["latent"].shape == a["obs"].shape, (a["latent"].shape, a["obs"].shape)

    def update(self, kind, iteration):
        k = UPDATE_KINDS[kind] if isinstance(kind, str) else int(kind)
        self.lib.orc_step_update(C.byref(self.s), k, iteration)

    def set_temp(self, t):
        self.s.temp = float(t)

    @property
    def evals(self):
        return int(self.s.evals)

    def llik(self):
        return float((self.a["cell_t"] + self.a["cell_o"]).sum())

    def prior(self):
        a = self.a
        return float((a["prior_eps_neg"] + a["prior_eps_pos"] + a["prior_r"]
                      + a["prior_coi"]).sum() + a["hyper"][0])
