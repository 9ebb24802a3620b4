// TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product; never linked by it.
//
// C-ABI harness around the UNMODIFIED reference sources, compiled in place from
// $(MOIRE_REF_SRC) (default /root/reference/src) against oracle/shim/ (see oracle/Makefile).
// It exists to (1) mint golden vectors for tests/golden/ (tests/golden/make_golden.py),
// (2) pin oracle/moire_oracle.c against the reference itself, and (3) serve as the timed
// CPU baseline (`bench.py --impl reference`, cpu_baseline.kind = "reference").
//
// Built twice: libmoire_ref32.so from the sources as they are (float arithmetic) and
// libmoire_ref64.so from a mechanically generated double variant (sed float->double, see
// Makefile).  Every floating-point value crosses this ABI as double (a float is exactly
// representable), so one ctypes binding serves both builds; ref_real_bytes() tells which.
//
// Private members of the reference's Chain are reached with the usual
// `#define private public` trick, applied after every standard header has been included.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <execution>
#include <iostream>
#include <limits>
#include <map>
#include <numeric>
#include <random>
#include <string>
#include <tuple>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include <Rcpp.h>
#include <boost/math/special_functions/binomial.hpp>
#include <boost/random.hpp>
#include <boost/range/algorithm.hpp>
#include <progress.hpp>

#define private public
#include "chain.h"
#include "mcmc.h"
#undef private
#include "mcmc_utils.h"

namespace
{
using real = decltype(Chain::llik);

Parameters g_params;

std::vector<int> mask_to_indices(unsigned long long mask)
{
    std::vector<int> out;
    for (int b = 0; b < 64; ++b)
    {
        if (mask >> b & 1ULL) out.push_back(b);
    }
    return out;
}

unsigned long long indices_to_mask(const std::vector<int> &idx)
{
    unsigned long long m = 0;
    for (int i : idx) m |= 1ULL << i;
    return m;
}
}  // namespace

extern "C"
{
int ref_real_bytes() { return static_cast<int>(sizeof(real)); }

void ref_seed_runif(std::uint64_t seed) { moire_shim::seed_runif(seed); }

// log(U) exactly as src/mcmc.cpp:212 draws it, from the shim's (seedable) stream.
double ref_draw_log_runif()
{
    real u = log(R::runif(0, 1));
    return u;
}

// ---- a6: CombinationIndicesGenerator (src/combination_indices_generator.cpp) -------------
long ref_combinations(int n, int r, signed char *out, long max_rows,
                      unsigned long *num_combinations, unsigned long *generated)
{
    CombinationIndicesGenerator g;
    g.reset(n, r);
    long rows = 0;
    while (!g.completed && rows < max_rows)
    {
        for (int i = 0; i < r; ++i) out[rows * r + i] = g.curr[i];
        ++rows;
        g.next();
    }
    *num_combinations = g.numCombinations;
    *generated = g.generated;
    return rows;
}

// ---- a3 / a4: probAnyMissingFunctor (src/prob_any_missing.cpp) -------------------------------
double ref_pam_scalar(const double *q, int k, int n)
{
    probAnyMissingFunctor f;
    std::vector<real> v(q, q + k);
    return f(v, n);
}

void ref_pam_vectorized(const double *q, int k, int n_max, double *out)
{
    probAnyMissingFunctor f;
    std::vector<real> v(q, q + k);
    const auto res = f.vectorized(v, static_cast<unsigned>(n_max));
    for (int i = 0; i < n_max; ++i) out[i] = res[i];
}

// ---- data + parameters ------------------------------------------------------------------------
// obs: concatenated 0/1 bytes, locus-major, then sample, then allele (ragged by locus).
// is_missing: L*N bytes, locus-major.
void ref_set_data(int L, int N, const int *num_alleles, const unsigned char *obs,
                  const unsigned char *is_missing)
{
    GenotypingData::observed_alleles.assign(L, {});
    GenotypingData::is_missing_.assign(L, std::vector<bool>(N, false));
    GenotypingData::num_alleles.assign(num_alleles, num_alleles + L);
    GenotypingData::observed_coi.assign(N, 0);
    GenotypingData::num_loci = L;
    GenotypingData::num_samples = N;
    GenotypingData::max_alleles = 0;
    std::size_t off = 0;
    for (int j = 0; j < L; ++j)
    {
        const int A = num_alleles[j];
        GenotypingData::max_alleles = std::max(GenotypingData::max_alleles, A);
        GenotypingData::observed_alleles[j].assign(N, std::vector<int>(A, 0));
        for (int i = 0; i < N; ++i)
        {
            int tot = 0;
            for (int a = 0; a < A; ++a)
            {
                const int v = obs[off++];
                GenotypingData::observed_alleles[j][i][a] = v;
                tot += v;
            }
            GenotypingData::observed_coi[i] = std::max(GenotypingData::observed_coi[i], tot);
            GenotypingData::is_missing_[j][i] = is_missing[static_cast<std::size_t>(j) * N + i] != 0;
        }
    }
}

// vals order (doubles): 0 allow_relatedness, 1 thin, 2 burnin, 3 samples, 4 pt_num_threads,
// 5 adapt_temp, 6 pre_adapt_steps, 7 temp_adapt_steps, 8 max_initialization_tries,
// 9 record_latent_genotypes, 10 max_coi, 11 mean_coi_shape, 12 mean_coi_scale,
// 13 max_eps_pos, 14 eps_pos_alpha, 15 eps_pos_beta, 16 max_eps_neg, 17 eps_neg_alpha,
// 18 eps_neg_beta, 19 r_alpha, 20 r_beta          (src/parameters.cpp:9-37)
void ref_set_params(const double *v, const double *pt_chains, int T)
{
    Parameters p;
    p.verbose = false;
    p.simple_verbose = false;
    p.use_message = false;
    p.chain_number = 1;
    p.max_runtime = std::numeric_limits<real>::infinity();
    p.allow_relatedness = v[0] != 0;
    p.thin = static_cast<int>(v[1]);
    p.burnin = static_cast<int>(v[2]);
    p.samples = static_cast<int>(v[3]);
    p.pt_num_threads = static_cast<int>(v[4]);
    p.adapt_temp = v[5] != 0;
    p.pre_adapt_steps = static_cast<int>(v[6]);
    p.temp_adapt_steps = static_cast<int>(v[7]);
    p.max_initialization_tries = static_cast<int>(v[8]);
    p.record_latent_genotypes = v[9] != 0;
    p.max_coi = static_cast<int>(v[10]);
    p.mean_coi_shape = v[11];
    p.mean_coi_scale = v[12];
    p.max_eps_pos = v[13];
    p.eps_pos_alpha = v[14];
    p.eps_pos_beta = v[15];
    p.max_eps_neg = v[16];
    p.eps_neg_alpha = v[17];
    p.eps_neg_beta = v[18];
    p.r_alpha = v[19];
    p.r_beta = v[20];
    p.pt_chains.assign(pt_chains, pt_chains + T);
    g_params = p;
}

// ---- one Chain ---------------------------------------------------------------------------------
void *ref_chain_new(double temp, std::uint64_t seed)
{
    GenotypingData gd;
    Chain *c = new Chain(gd, g_params, static_cast<real>(temp));
    // The constructor seeds from std::random_device (src/sampler.cpp:14-19); re-seed and
    // re-initialise so the starting state is a deterministic function of `seed`.
    // A fresh Sampler also drops any value cached inside the std:: distributions.
    c->sampler = Sampler();
    c->sampler.eng.seed(static_cast<std::uint_fast32_t>(seed));
    c->mean_coi =
        c->sampler.sample_mean_coi(g_params.mean_coi_shape, g_params.mean_coi_scale) + 1;
    c->initialize_parameters();
    return c;
}

void ref_chain_free(void *h) { delete static_cast<Chain *>(h); }

void ref_chain_reseed(void *h, std::uint64_t seed)
{
    static_cast<Chain *>(h)->sampler.eng.seed(static_cast<std::uint_fast32_t>(seed));
}

void ref_chain_set_temp(void *h, double t) { static_cast<Chain *>(h)->set_temp(t); }

// Load a full parameter state, then rebuild every likelihood/prior cache exactly as
// Chain::initialize_likelihood does (src/chain.cpp:1166-1223).
// latent / lg_adj: [N][L] sample-major; p: [L][a_pad].
void ref_chain_set_state(void *h, const int *m, const double *r, const double *eps_pos,
                         const double *eps_neg, const double *p, int a_pad,
                         const unsigned long long *latent, const double *lg_adj,
                         double mean_coi)
{
    Chain &c = *static_cast<Chain *>(h);
    const int N = static_cast<int>(GenotypingData::num_samples);
    const int L = static_cast<int>(GenotypingData::num_loci);
    for (int i = 0; i < N; ++i)
    {
        c.m[i] = m[i];
        c.r[i] = r[i];
        c.eps_pos[i] = eps_pos[i];
        c.eps_neg[i] = eps_neg[i];
    }
    for (int j = 0; j < L; ++j)
    {
        const int A = GenotypingData::num_alleles[j];
        for (int a = 0; a < A; ++a) c.p[j][a] = p[static_cast<std::size_t>(j) * a_pad + a];
        for (int i = 0; i < N; ++i)
        {
            const std::size_t cell = static_cast<std::size_t>(i) * L + j;
            c.latent_genotypes_new[j][i] = mask_to_indices(latent[cell]);
            c.latent_genotypes_old[j][i] = c.latent_genotypes_new[j][i];
            c.lg_adj_new[j][i] = lg_adj[cell];
            c.lg_adj_old[j][i] = lg_adj[cell];
        }
    }
    c.mean_coi = mean_coi;
    c.initialize_likelihood();
}

// Everything a parity test needs.  Arrays may be null to skip.
// cell_llik: [N][L] (genotyping_llik_new, src/chain.h:85); priors: [4][N] in the order
// eps_neg, eps_pos, coi, relatedness; scalars: llik, prior, temp, mean_coi,
// mean_coi_hyper_prior_new, mean_coi_var, mean_coi_accept.
void ref_chain_get_state(void *h, int *m, double *r, double *eps_pos, double *eps_neg,
                         double *p, int a_pad, unsigned long long *latent, double *lg_adj,
                         double *cell_llik, double *priors, double *scalars)
{
    Chain &c = *static_cast<Chain *>(h);
    const int N = static_cast<int>(GenotypingData::num_samples);
    const int L = static_cast<int>(GenotypingData::num_loci);
    for (int i = 0; i < N; ++i)
    {
        if (m) m[i] = c.m[i];
        if (r) r[i] = c.r[i];
        if (eps_pos) eps_pos[i] = c.eps_pos[i];
        if (eps_neg) eps_neg[i] = c.eps_neg[i];
        if (priors)
        {
            priors[0 * N + i] = c.eps_neg_prior_new[i];
            priors[1 * N + i] = c.eps_pos_prior_new[i];
            priors[2 * N + i] = c.coi_prior_new[i];
            priors[3 * N + i] = c.relatedness_prior_new[i];
        }
    }
    for (int j = 0; j < L; ++j)
    {
        const int A = GenotypingData::num_alleles[j];
        if (p)
        {
            for (int a = 0; a < a_pad; ++a)
                p[static_cast<std::size_t>(j) * a_pad + a] = a < A ? c.p[j][a] : 0.0;
        }
        for (int i = 0; i < N; ++i)
        {
            const std::size_t cell = static_cast<std::size_t>(i) * L + j;
            if (latent) latent[cell] = indices_to_mask(c.latent_genotypes_new[j][i]);
            if (lg_adj) lg_adj[cell] = c.lg_adj_new[j][i];
            if (cell_llik) cell_llik[cell] = c.genotyping_llik_new[cell];
        }
    }
    if (scalars)
    {
        scalars[0] = c.llik;
        scalars[1] = c.prior;
        scalars[2] = c.temp;
        scalars[3] = c.mean_coi;
        scalars[4] = c.mean_coi_hyper_prior_new;
        scalars[5] = c.mean_coi_var;
        scalars[6] = c.mean_coi_accept;
    }
}

// Proposal scales and acceptance counters (src/chain.h:104-141); [6][N] ints then [4][N] sds.
void ref_chain_get_tuning(void *h, int *accepts, double *sds)
{
    Chain &c = *static_cast<Chain *>(h);
    const int N = static_cast<int>(GenotypingData::num_samples);
    for (int i = 0; i < N; ++i)
    {
        accepts[0 * N + i] = c.eps_neg_accept[i];
        accepts[1 * N + i] = c.eps_pos_accept[i];
        accepts[2 * N + i] = c.m_accept[i];
        accepts[3 * N + i] = c.r_accept[i];
        accepts[4 * N + i] = c.m_r_accept[i];
        accepts[5 * N + i] = c.sample_accept[i];
        sds[0 * N + i] = c.eps_neg_var[i];
        sds[1 * N + i] = c.eps_pos_var[i];
        sds[2 * N + i] = c.r_var[i];
        sds[3 * N + i] = c.m_r_var[i];
    }
}

// The SALT sampler's per-allele tuning state (src/chain.h:126-128, src/chain.cpp:479-562):
// [L][a_pad] each.
void ref_chain_get_p_tuning(void *h, int a_pad, int *p_accept, int *p_attempt, double *p_prop_var)
{
    Chain &c = *static_cast<Chain *>(h);
    for (std::size_t j = 0; j < c.p_prop_var.size(); ++j)
        for (std::size_t a = 0; a < c.p_prop_var[j].size() && a < static_cast<std::size_t>(a_pad); ++a)
        {
            p_accept[j * a_pad + a] = c.p_accept[j][a];
            p_attempt[j * a_pad + a] = c.p_attempt[j][a];
            p_prop_var[j * a_pad + a] = c.p_prop_var[j][a];
        }
}

// kind: 0 eps_neg, 1 eps_pos, 2 p, 3 m, 4 r, 5 eff_coi, 6 samples, 7 mean_coi
// (the order MCMC::burnin runs them, src/mcmc.cpp:162-173).
void ref_chain_update(void *h, int kind, int iteration)
{
    Chain &c = *static_cast<Chain *>(h);
    switch (kind)
    {
        case 0: c.update_eps_neg(iteration); break;
        case 1: c.update_eps_pos(iteration); break;
        case 2: c.update_p(iteration); break;
        case 3: c.update_m(iteration); break;
        case 4: c.update_r(iteration); break;
        case 5: c.update_eff_coi(iteration); break;
        case 6: c.update_samples(iteration); break;
        case 7: c.update_mean_coi(iteration); break;
        default: break;
    }
}

// a2: Chain::calc_transmission_process (src/chain.cpp:858-921).  allow_relatedness comes from
// the parameters the chain was built with.
double ref_chain_transmission(void *h, const int *idx, int k, const double *p, int A, int coi,
                              double relatedness)
{
    Chain &c = *static_cast<Chain *>(h);
    std::vector<int> iv(idx, idx + k);
    std::vector<real> pv(p, p + A);
    return c.calc_transmission_process(iv, pv, coi, static_cast<real>(relatedness));
}

// a7: Chain::calc_observation_process (src/chain.cpp:923-961).
double ref_chain_observation(void *h, const int *idx, int k, const int *obs, int A,
                             double eps_neg, double eps_pos)
{
    Chain &c = *static_cast<Chain *>(h);
    std::vector<int> iv(idx, idx + k);
    std::vector<int> ov(obs, obs + A);
    return c.calc_observation_process(iv, ov, static_cast<real>(eps_neg),
                                      static_cast<real>(eps_pos));
}

// ---- Sampler pieces (src/sampler.cpp) ----------------------------------------------------------
// a8.  Returns k (size of the proposed latent set); -1 if the reference returned the empty set.
int ref_sample_latent_genotype(std::uint64_t seed, const int *obs, int A, int coi,
                               double eps_pos, double eps_neg, int *out_idx, double *log_prob)
{
    Sampler s;
    s.eng.seed(static_cast<std::uint_fast32_t>(seed));
    std::vector<int> ov(obs, obs + A);
    const LatentGenotype lg =
        s.sample_latent_genotype(ov, coi, static_cast<real>(eps_pos), static_cast<real>(eps_neg));
    for (std::size_t i = 0; i < lg.value.size(); ++i) out_idx[i] = lg.value[i];
    *log_prob = lg.log_prob;
    return lg.value.empty() ? -1 : static_cast<int>(lg.value.size());
}

// n draws of Sampler::sample_coi_delta(mean) (src/sampler.cpp:152-156) from one seeded sampler.
void ref_sample_coi_delta_many(std::uint64_t seed, double mean, int n, int *out)
{
    Sampler s;
    s.eng.seed(static_cast<std::uint_fast32_t>(seed));
    for (int i = 0; i < n; ++i) out[i] = s.sample_coi_delta(static_cast<real>(mean));
}

// sample_constrained (src/sampler.cpp:176-190) plus the normal draw it consumed, obtained by
// replaying an identically seeded engine through an identical fresh distribution.
void ref_sample_constrained(std::uint64_t seed, double curr, double sd, double lo, double hi,
                            double *z, double *prop, double *adj)
{
    Sampler s;
    s.eng.seed(static_cast<std::uint_fast32_t>(seed));
    std::ranlux24_base replay(static_cast<std::uint_fast32_t>(seed));
    std::normal_distribution<real> nd(0, static_cast<real>(sd));
    *z = nd(replay);
    const auto pa = s.sample_constrained(static_cast<real>(curr), static_cast<real>(sd),
                                         static_cast<real>(lo), static_cast<real>(hi));
    *prop = std::get<0>(pa);
    *adj = std::get<1>(pa);
}

double ref_dztpois(int x, double lambda)
{
    Sampler s;
    return s.dztpois(x, static_cast<real>(lambda));
}

double ref_dbinom(int x, int size, double prob)
{
    Sampler s;
    return s.dbinom(x, size, static_cast<real>(prob));
}

double ref_beta_log_prior(double x, double a, double b)
{
    Sampler s;
    return s.get_epsilon_log_prior(static_cast<real>(x), static_cast<real>(a),
                                   static_cast<real>(b));
}

double ref_gamma_log_prior(double x, double shape, double scale)
{
    Sampler s;
    return s.get_coi_mean_log_hyper_prior(static_cast<real>(x), static_cast<real>(shape),
                                          static_cast<real>(scale));
}

double ref_min_sampled() { return std::numeric_limits<real>::min(); }

// ---- SALT proposal pieces (src/mcmc_utils.h:150-313, driven as src/chain.cpp:481-501) ----------
double ref_logit(double x) { return UtilFunctions::logit(static_cast<real>(x)); }

void ref_log_pq(double x, double *lp, double *lq)
{
    const auto r = UtilFunctions::log_pq(static_cast<real>(x));
    *lp = r.first;
    *lq = r.second;
}

double ref_logit_sum(const double *x, int n)
{
    std::vector<real> v(x, x + n);
    return UtilFunctions::logitSum(v);
}

void ref_logit_scale(const double *x, int n, double scale, double *out)
{
    std::vector<real> v(x, x + n);
    const auto r = UtilFunctions::logitScale(v, static_cast<real>(scale));
    for (int i = 0; i < n; ++i) out[i] = r[i];
}

double ref_expit(double x) { return UtilFunctions::expit(static_cast<real>(x)); }

// Given the current simplex, the picked allele and the proposed logit, the proposed simplex and
// the log proposal-ratio adjustment (the deterministic part of one SALT proposal).
void ref_salt_proposal(const double *p, int A, int idx, double logit_prop_in, double *prop_p,
                       double *log_adj)
{
    std::vector<real> pv(p, p + A);
    auto lv = UtilFunctions::logitVec(pv);
    const real logit_curr = lv[idx];
    const real logit_prop = static_cast<real>(logit_prop_in);
    const auto cur = UtilFunctions::log_pq(logit_curr);
    const auto pro = UtilFunctions::log_pq(logit_prop);
    lv.erase(lv.begin() + idx);
    const real ls = pro.second - UtilFunctions::logitSum(lv);
    lv = UtilFunctions::logitScale(lv, ls);
    lv.insert(lv.begin() + idx, logit_prop);
    const real adj = (cur.first - pro.first) + (A - 1) * (cur.second - pro.second);
    const auto out = UtilFunctions::expitVec(lv);
    for (int a = 0; a < A; ++a) prop_p[a] = out[a];
    *log_adj = adj;
}

double ref_log_sum_exp(const double *x, int n)
{
    std::vector<real> v(x, x + n);
    return UtilFunctions::logSumExp(v);
}

// ---- MCMC driver (src/mcmc.cpp) ----------------------------------------------------------------
void *ref_mcmc_new()
{
    GenotypingData gd;
    return new MCMC(gd, g_params);
}

void ref_mcmc_free(void *h) { delete static_cast<MCMC *>(h); }

int ref_mcmc_init_failed(void *h) { return static_cast<MCMC *>(h)->initialization_failed; }

void ref_mcmc_reseed(void *h, std::uint64_t seed)
{
    MCMC &mc = *static_cast<MCMC *>(h);
    for (std::size_t i = 0; i < mc.chains.size(); ++i)
    {
        Chain &c = mc.chains[i];
        c.sampler = Sampler();
        c.sampler.eng.seed(static_cast<std::uint_fast32_t>(seed + 7919 * i));
        c.mean_coi =
            c.sampler.sample_mean_coi(g_params.mean_coi_shape, g_params.mean_coi_scale) + 1;
        int tries = g_params.max_initialization_tries;
        do
        {
            c.initialize_parameters();
        } while (!std::isfinite(c.get_llik()) && tries-- > 0);
    }
}

void ref_mcmc_burnin(void *h, int step) { static_cast<MCMC *>(h)->burnin(step); }
void ref_mcmc_sample(void *h, int step) { static_cast<MCMC *>(h)->sample(step); }
void ref_mcmc_finalize(void *h) { static_cast<MCMC *>(h)->finalize(); }
double ref_mcmc_llik(void *h) { return static_cast<MCMC *>(h)->get_llik(); }
double ref_mcmc_prior(void *h) { return static_cast<MCMC *>(h)->get_prior(); }

// Time `steps` burn-in steps (after `warmup` untimed ones) the way src/main.cpp:54-70 drives
// them.  Returns wall seconds for the timed steps.
double ref_mcmc_time_burnin(void *h, int first_step, int warmup, int steps)
{
    MCMC &mc = *static_cast<MCMC *>(h);
    int step = first_step;
    for (int i = 0; i < warmup; ++i) mc.burnin(step++);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; ++i) mc.burnin(step++);
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- evaluation counter ("count them exactly on both sides") ----------------------
// Chain::calculate_genotype_likelihood (src/chain.cpp:1114-1134) asks
// GenotypingData::is_missing(locus, sample) exactly once per call, nowhere else on the sampler's
// path, and evaluates the cell iff the answer is false.  The reference sources stay untouched: the
// linker redirects that one call (-Wl,--wrap, oracle/Makefile) through the function below, which
// forwards to the real member and counts the non-missing answers per OpenMP thread.
namespace
{
struct alignas(64) EvalSlot { unsigned long long n; };
EvalSlot g_eval_slots[1024];
}
bool __real__ZNK14GenotypingData10is_missingEii(const GenotypingData *, int, int);
bool __wrap__ZNK14GenotypingData10is_missingEii(const GenotypingData *self, int locus, int sample)
{
    const bool missing = __real__ZNK14GenotypingData10is_missingEii(self, locus, sample);
#ifdef _OPENMP
    const int slot = omp_get_thread_num() & 1023;
#else
    const int slot = 0;
#endif
    g_eval_slots[slot].n += missing ? 0ull : 1ull;
    return missing;
}

// Non-missing cell likelihood evaluations since the last reset (all threads).
unsigned long long ref_eval_count()
{
    unsigned long long tot = 0;
    for (const auto &sl : g_eval_slots) tot += sl.n;
    return tot;
}

void ref_eval_count_reset()
{
    for (auto &sl : g_eval_slots) sl.n = 0;
}

int ref_omp_max_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Posterior summaries of the stored draws (src/mcmc.cpp:327-347 fills the stores).
// mean_m[N], mean_r[N], mean_eps_neg[N], mean_eps_pos[N], mean_p[L*a_pad]; returns #draws.
int ref_mcmc_summary(void *h, double *mean_m, double *mean_r, double *mean_eps_neg,
                     double *mean_eps_pos, double *mean_p, int a_pad, double *mean_mean_coi)
{
    MCMC &mc = *static_cast<MCMC *>(h);
    const int N = static_cast<int>(GenotypingData::num_samples);
    const int L = static_cast<int>(GenotypingData::num_loci);
    const int D = static_cast<int>(mc.mean_coi_store.size());
    if (D == 0) return 0;
    for (int i = 0; i < N; ++i)
    {
        double sm = 0, sr = 0, sn = 0, sp = 0;
        for (int d = 0; d < D; ++d)
        {
            sm += mc.m_store[i][d];
            sr += mc.r_store[i][d];
            sn += mc.eps_neg_store[i][d];
            sp += mc.eps_pos_store[i][d];
        }
        mean_m[i] = sm / D;
        mean_r[i] = sr / D;
        mean_eps_neg[i] = sn / D;
        mean_eps_pos[i] = sp / D;
    }
    for (int j = 0; j < L; ++j)
    {
        const int A = GenotypingData::num_alleles[j];
        for (int a = 0; a < a_pad; ++a)
        {
            double s = 0;
            if (a < A)
                for (int d = 0; d < D; ++d) s += mc.p_store[j][d][a];
            mean_p[static_cast<std::size_t>(j) * a_pad + a] = s / D;
        }
    }
    double s = 0;
    for (int d = 0; d < D; ++d) s += mc.mean_coi_store[d];
    *mean_mean_coi = s / D;
    return D;
}

// Raw per-draw COI store, [N][D] (for quantiles in the posterior-equivalence study).
int ref_mcmc_coi_draws(void *h, int *out, int max_draws)
{
    MCMC &mc = *static_cast<MCMC *>(h);
    const int N = static_cast<int>(GenotypingData::num_samples);
    const int D = std::min<int>(static_cast<int>(mc.mean_coi_store.size()), max_draws);
    for (int i = 0; i < N; ++i)
        for (int d = 0; d < D; ++d) out[static_cast<std::size_t>(i) * max_draws + d] = mc.m_store[i][d];
    return D;
}

int ref_mcmc_trace(void *h, double *llik_burnin, int nb, double *llik_sample, int ns)
{
    MCMC &mc = *static_cast<MCMC *>(h);
    for (int i = 0; i < nb && i < static_cast<int>(mc.llik_burnin.size()); ++i)
        llik_burnin[i] = mc.llik_burnin[i];
    for (int i = 0; i < ns && i < static_cast<int>(mc.llik_sample.size()); ++i)
        llik_sample[i] = mc.llik_sample[i];
    return static_cast<int>(mc.llik_sample.size());
}

// swap_chains as a function of (per-chain llik, temperatures, the shim's uniform stream):
// set the T chains' llik/temp, run one pass, read back swap_indices / temps / counters.
// (src/mcmc.cpp:190-240).  chain_llik/chain_temp are indexed by chain id.
void ref_mcmc_swap_pass(void *h, const double *chain_llik, const double *chain_temp,
                        const long *swap_indices_in, int even_swap, int step, int burnin,
                        long *swap_indices_out, double *chain_temp_out, int *hot_out,
                        double *swap_barriers_io, long *swap_acceptances_io, long *num_swaps_io)
{
    MCMC &mc = *static_cast<MCMC *>(h);
    const std::size_t T = mc.chains.size();
    for (std::size_t c = 0; c < T; ++c)
    {
        mc.chains[c].set_llik(static_cast<real>(chain_llik[c]));
        mc.chains[c].set_temp(static_cast<real>(chain_temp[c]));
        mc.swap_indices[c] = static_cast<std::size_t>(swap_indices_in[c]);
    }
    for (std::size_t c = 0; c + 1 < T; ++c)
    {
        mc.swap_barriers[c] = static_cast<real>(swap_barriers_io[c]);
        mc.swap_acceptances[c] = static_cast<std::size_t>(swap_acceptances_io[c]);
    }
    mc.even_swap = even_swap != 0;
    mc.num_swaps = static_cast<std::size_t>(*num_swaps_io);
    mc.swap_chains(step, burnin != 0);
    for (std::size_t c = 0; c < T; ++c)
    {
        swap_indices_out[c] = static_cast<long>(mc.swap_indices[c]);
        chain_temp_out[c] = mc.chains[c].get_temp();
        hot_out[c] = mc.chains[c].get_hot();
    }
    for (std::size_t c = 0; c + 1 < T; ++c)
    {
        swap_barriers_io[c] = mc.swap_barriers[c];
        swap_acceptances_io[c] = static_cast<long>(mc.swap_acceptances[c]);
    }
    *num_swaps_io = static_cast<long>(mc.num_swaps);
}

// adapt_temp as a function of (temp_gradient, swap_barriers, num_swaps) (src/mcmc.cpp:244-304).
void ref_mcmc_adapt_temp(void *h, const double *temp_gradient_in, const double *swap_barriers_in,
                         long num_swaps, double *temp_gradient_out)
{
    MCMC &mc = *static_cast<MCMC *>(h);
    const std::size_t T = mc.chains.size();
    for (std::size_t c = 0; c < T; ++c)
    {
        mc.temp_gradient[c] = static_cast<real>(temp_gradient_in[c]);
        mc.swap_indices[c] = c;
    }
    for (std::size_t c = 0; c + 1 < T; ++c)
        mc.swap_barriers[c] = static_cast<real>(swap_barriers_in[c]);
    mc.num_swaps = static_cast<std::size_t>(num_swaps);
    mc.adapt_temp();
    for (std::size_t c = 0; c < T; ++c) temp_gradient_out[c] = mc.temp_gradient[c];
}

void ref_mcmc_get_ladder(void *h, double *temp_gradient, long *swap_indices, double *chain_temps,
                         double *chain_lliks)
{
    MCMC &mc = *static_cast<MCMC *>(h);
    const std::size_t T = mc.chains.size();
    for (std::size_t c = 0; c < T; ++c)
    {
        temp_gradient[c] = mc.temp_gradient[c];
        swap_indices[c] = static_cast<long>(mc.swap_indices[c]);
        chain_temps[c] = mc.chains[c].get_temp();
        chain_lliks[c] = mc.chains[c].get_llik();
    }
}
}  // extern "C"
