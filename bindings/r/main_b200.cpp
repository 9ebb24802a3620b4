// src/main.cpp of moire, re-pointed at libmoire_b200.so -- the ONE file a maintainer replaces.
//
// Same entry point, same arguments, same 23-field result as the reference's run_mcmc
// (src/main.cpp:13-191); `R/mcmc.R`, `R/summary.R` and the exported R API are untouched.  The
// argument parsers are the reference's own (Parameters, GenotypingData: src/parameters.cpp,
// src/genotyping_data.cpp), so are the progress bar and the initialisation diagnostics; what used
// to be `MCMC mcmc(genotyping_data, params)` plus the two loops is the C ABI of
// include/moire_mcmc.h.  Build: add this file as src/main.cpp, `PKG_CPPFLAGS += -I$(MOIRE_B200_HOME)/include`
// and `PKG_LIBS += -L$(MOIRE_B200_HOME)/moire_b200 -lmoire_b200` in src/Makevars.
//
// Compiled and run in this repository's tests against a small functional Rcpp stand-in
// (oracle/shim_r/Rcpp.h, oracle/Makefile target `binding`, tests/test_binding.py): on the CPU in
// the driver's ladder-only mode (MOIRE_B200_DEVICE=-1: every field of the list, nothing evaluated),
// on the GPU against `moire_b200.mcmc.run_mcmc` with the same seed.
//
// Environment (all optional): MOIRE_B200_DEVICE (CUDA device, default 0), MOIRE_B200_SEED (default:
// std::random_device, like the reference's Sampler).
#include <cstdint>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "genotyping_data.h"
#include "initialization_diagnostics.h"
#include "mcmc_progress_bar.h"
#include "mcmc_utils.h"
#include "parameters.h"

#include <progress.hpp>

#include "moire_mcmc.h"

namespace
{
struct ProgressCtx
{
    const Parameters *params;
    MCMCProgressBar *pb;
    Progress *p;
};

// src/main.cpp:56-69, 75-88: user interrupt, the 100-step message, the progress bar
int progress_cb(void *ctx_, int32_t phase, int32_t step, double posterior, int32_t hot_chain)
{
    ProgressCtx *ctx = static_cast<ProgressCtx *>(ctx_);
    try
    {
        Rcpp::checkUserInterrupt();
    }
    catch (...)
    {
        return 1;
    }
    const Parameters &params = *ctx->params;
    if (params.verbose && params.simple_verbose && (step % 100) == 0)
    {
        UtilFunctions::print("Chain", params.chain_number, ":", phase == 0 ? step : params.burnin + step, "/",
                             params.burnin + params.samples, "completed");
    }
    else if (!params.simple_verbose)
    {
        ctx->pb->set_llik((float)posterior);
        ctx->pb->set_hot_chain(hot_chain);
        ctx->p->increment();
    }
    return 0;
}

struct Handle  // moire_mcmc_destroy on every way out
{
    moire_mcmc *h = nullptr;
    ~Handle()
    {
        if (h) moire_mcmc_destroy(h);
    }
};

void check(int rc, moire_mcmc *h, const char *what)
{
    if (rc != MOIRE_OK) Rcpp::stop(std::string(what) + ": " + moire_mcmc_last_error(h));
}
}  // namespace

//----------------------------------------------
// [[Rcpp::export(name='run_mcmc_rcpp')]]
Rcpp::List run_mcmc(Rcpp::List args)
{
    Parameters params(args);
    GenotypingData gd(args);

    if (params.verbose && !params.simple_verbose)
    {
        UtilFunctions::print("-- Starting MCMC --");
        UtilFunctions::print("Total Burnin:", params.burnin);
        UtilFunctions::print("Total Samples:", params.samples);
        UtilFunctions::print("Thinning:", params.thin);
        UtilFunctions::print("Allow Relatedness:", params.allow_relatedness ? "Yes" : "No");
        UtilFunctions::print("Parallel Tempering:", params.pt_chains.size() > 1 ? "Yes" : "No");
        UtilFunctions::print("Adapt Temperature:", params.adapt_temp ? "Yes" : "No");
    }

    // ---- GenotypingData -> SoA bit masks (moire_data, include/moire_engine.h)
    const int N = (int)gd.num_samples, L = (int)gd.num_loci;
    std::vector<int32_t> num_alleles(gd.num_alleles.begin(), gd.num_alleles.end());
    // allele sets are W 64-bit words each: W = 1 in libmoire_b200.so, 2 in libmoire_b200_w128.so (link the
    // latter for data with a locus of 65..128 alleles; include/moire_engine.h, "Mask width")
    const size_t W = (size_t)moire_engine_mask_words();
    for (int a : num_alleles)
        if (a > moire_engine_max_alleles())
            Rcpp::stop("this build of libmoire_b200 holds at most " + std::to_string(moire_engine_max_alleles()) +
                       " alleles per locus (libmoire_b200_w128: 128)");
    std::vector<uint64_t> obs((size_t)N * L * W, 0);
    std::vector<uint8_t> missing((size_t)N * L, 0);
    for (int j = 0; j < L; ++j)
        for (int i = 0; i < N; ++i)
        {
            const std::vector<int> &o = gd.get_observed_alleles(j, i);  // src/genotyping_data.cpp:49
            for (size_t a = 0; a < o.size(); ++a)
                if (o[a]) obs[((size_t)i * L + j) * W + (a >> 6)] |= 1ull << (a & 63);
            missing[(size_t)i * L + j] = gd.is_missing(j, i);           // src/genotyping_data.cpp:55
        }
    std::vector<int32_t> obs_coi(gd.observed_coi.begin(), gd.observed_coi.end());
    const moire_data d{N, L, num_alleles.data(), obs.data(), missing.data(), obs_coi.data()};

    const moire_params m{params.allow_relatedness, params.burnin, params.max_coi, MOIRE_PIPELINE_AUTO,
                         params.mean_coi_shape, params.mean_coi_scale,
                         params.max_eps_pos, params.eps_pos_alpha, params.eps_pos_beta,
                         params.max_eps_neg, params.eps_neg_alpha, params.eps_neg_beta,
                         params.r_alpha, params.r_beta};
    const moire_mcmc_params mp{params.thin, params.burnin, params.samples, params.adapt_temp,
                               params.pre_adapt_steps, params.temp_adapt_steps,
                               params.max_initialization_tries, params.record_latent_genotypes,
                               (double)params.max_runtime};
    const std::vector<double> ladder(params.pt_chains.begin(), params.pt_chains.end());
    const int T = (int)ladder.size();

    const char *dev_env = std::getenv("MOIRE_B200_DEVICE");
    const char *seed_env = std::getenv("MOIRE_B200_SEED");
    const int device = dev_env ? std::atoi(dev_env) : 0;
    const uint64_t seed = seed_env ? std::strtoull(seed_env, nullptr, 10)
                                   : ((uint64_t)std::random_device{}() << 32) | std::random_device{}();

    Handle guard;
    if (moire_mcmc_create(&d, &m, &mp, ladder.data(), T, /*rank*/ 0, /*world*/ 1, device, seed, &guard.h) != MOIRE_OK)
        Rcpp::stop(std::string("moire_mcmc_create: ") + moire_mcmc_last_error(nullptr));
    moire_mcmc *h = guard.h;
    moire_engine *eng = moire_mcmc_engine(h);  // null in ladder-only mode

    int32_t failed = 0;
    check(moire_mcmc_init(h, &failed), h, "moire_mcmc_init");
    if (failed)  // src/main.cpp:39-46 with MCMC::MCMC's bookkeeping (src/mcmc.cpp:106-152)
    {
        InitializationDiagnostics diag;
        diag.max_initialization_tries = params.max_initialization_tries;
        int32_t n = 0;
        moire_engine_read_init_log(eng, nullptr, 0, &n);
        std::vector<int32_t> rec(3 * (size_t)std::max(n, 1));
        moire_engine_read_init_log(eng, rec.data(), n, &n);
        for (int k = 0; k < n; ++k) diag.record_ill_conditioned(rec[3 * k + 1], rec[3 * k + 2]);
        std::vector<int32_t> ill(T, 0), pri(5 * (size_t)T, 0);
        moire_mcmc_get_init_counts(h, ill.data(), pri.data());
        static const char *terms[5] = {"eps_neg", "eps_pos", "relatedness", "coi", "mean_coi_hyper"};
        for (int c = 0; c < T; ++c)
        {
            ++diag.chains_attempted;
            diag.ill_conditioned_per_chain.push_back(ill[c]);
            for (int q = 0; q < 5; ++q)
                for (int r = 0; r < pri[5 * c + q]; ++r) diag.record_prior_nonfinite(terms[q]);
        }
        diag.classify();
        return Rcpp::List::create(Rcpp::Named("initialization_failed") = true,
                                  Rcpp::Named("diagnostics") = diag.to_list());
    }

    MCMCProgressBar pb(params.burnin, params.samples, params.use_message);
    Progress p(params.burnin + params.samples, params.verbose, pb);
    ProgressCtx ctx{&params, &pb, &p};
    int32_t reached = 0;
    double minutes = 0.0;
    check(moire_mcmc_run(h, progress_cb, &ctx, &reached, &minutes), h, "moire_mcmc_run");  // interrupt / CUDA error -> R error

    // ---- the stores of src/mcmc.h:25-52, in the nesting R/summary.R indexes into
    moire_mcmc_sizes sz;
    check(moire_mcmc_get_sizes(h, &sz), h, "get_sizes");
    const int D = sz.num_draws, AP = sz.a_pad;
    std::vector<double> llik_burnin(sz.burnin_steps), prior_burnin(sz.burnin_steps), posterior_burnin(sz.burnin_steps),
        llik_sample(sz.sample_steps), prior_sample(sz.sample_steps), posterior_sample(sz.sample_steps);
    check(moire_mcmc_get_traces(h, llik_burnin.data(), prior_burnin.data(), posterior_burnin.data(),
                                llik_sample.data(), prior_sample.data(), posterior_sample.data()), h, "get_traces");
    std::vector<int32_t> swap_store(sz.num_swap_store), swap_indices(T);
    std::vector<int64_t> swap_acc(T > 1 ? T - 1 : 0);
    std::vector<double> swap_barriers(T > 1 ? T - 1 : 0), temp_gradient(T), chain_temps(T);
    check(moire_mcmc_get_ladder(h, swap_store.data(), swap_acc.data(), swap_barriers.data(), temp_gradient.data(),
                                swap_indices.data(), chain_temps.data()), h, "get_ladder");
    std::vector<int32_t> draw_step(D), coi((size_t)D * N);
    std::vector<double> freqs((size_t)D * L * AP), eps_neg((size_t)D * N), eps_pos((size_t)D * N),
        rel((size_t)D * N), data_llik((size_t)D * N), lam(D);
    std::vector<uint64_t> latent(params.record_latent_genotypes ? (size_t)D * N * L * W : 0);
    check(moire_mcmc_get_draws(h, draw_step.data(), freqs.data(), coi.data(), eps_neg.data(), eps_pos.data(),
                               rel.data(), data_llik.data(), lam.data(),
                               params.record_latent_genotypes ? latent.data() : nullptr), h, "get_draws");

    std::vector<std::vector<std::vector<double>>> p_store(L);          // [locus][draw][allele]
    for (int j = 0; j < L; ++j)
        for (int dr = 0; dr < D; ++dr)
        {
            const double *row = &freqs[((size_t)dr * L + j) * AP];
            p_store[j].emplace_back(row, row + num_alleles[j]);
        }
    auto by_sample = [&](const auto &flat) {                           // [sample][draw]
        using E = typename std::decay<decltype(flat[0])>::type;
        std::vector<std::vector<E>> out(N, std::vector<E>(D));
        for (int dr = 0; dr < D; ++dr)
            for (int i = 0; i < N; ++i) out[i][dr] = flat[(size_t)dr * N + i];
        return out;
    };
    // latent_genotypes[[sample]][[locus]][[draw]] = sorted 0-based allele indices (src/mcmc.cpp:340-345)
    std::vector<std::vector<std::vector<std::vector<int>>>> latent_store(
        N, std::vector<std::vector<std::vector<int>>>(L));
    if (params.record_latent_genotypes)
        for (int dr = 0; dr < D; ++dr)
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < L; ++j)
                {
                    std::vector<int> idx;
                    for (size_t w = 0; w < W; ++w)
                        for (uint64_t msk = latent[(((size_t)dr * N + i) * L + j) * W + w]; msk; msk &= msk - 1)
                            idx.push_back((int)(64 * w) + __builtin_ctzll(msk));
                    latent_store[i][j].push_back(idx);
                }

    // ---- acceptance counters and proposal scales per chain (src/main.cpp:99-136)
    Rcpp::List acceptance_rates;
    Rcpp::List sampling_variances;
    for (int c = 0; c < T && eng; ++c)
    {
        std::vector<int32_t> en_a(N), ep_a(N), m_a(N), r_a(N), mr_a(N), s_a(N), p_a((size_t)L * AP), p_att((size_t)L * AP);
        std::vector<double> en_v(N), ep_v(N), r_v(N), mr_v(N), p_v((size_t)L * AP);
        double mc_a = 0.0, mc_v = 0.0;
        moire_chain_diagnostics dg{en_a.data(), ep_a.data(), m_a.data(), r_a.data(), mr_a.data(), s_a.data(),
                                   en_v.data(), ep_v.data(), r_v.data(), mr_v.data(),
                                   p_a.data(), p_att.data(), p_v.data(), &mc_a, &mc_v};
        if (moire_engine_read_diagnostics(eng, c, &dg) != MOIRE_OK)
            Rcpp::stop(std::string("read_diagnostics: ") + moire_engine_last_error(eng));
        std::vector<std::vector<int>> p_accept(L);
        std::vector<std::vector<double>> p_var(L);
        for (int j = 0; j < L; ++j)
        {
            p_accept[j].assign(&p_a[(size_t)j * AP], &p_a[(size_t)j * AP] + num_alleles[j]);
            p_var[j].assign(&p_v[(size_t)j * AP], &p_v[(size_t)j * AP] + num_alleles[j]);
        }
        Rcpp::List ar;
        ar.push_back(Rcpp::wrap(p_accept));
        ar.push_back(Rcpp::wrap(m_a));
        ar.push_back(Rcpp::wrap(en_a));
        ar.push_back(Rcpp::wrap(ep_a));
        ar.push_back(Rcpp::wrap(r_a));
        ar.push_back(Rcpp::wrap(mr_a));
        ar.push_back(Rcpp::wrap(s_a));
        ar.push_back(Rcpp::wrap(mc_a));
        Rcpp::StringVector ar_names{"allele_freq_accept", "coi_accept", "eps_neg_accept", "eps_pos_accept",
                                    "r_accept", "m_r_accept", "full_sample_accept", "mean_coi_accept"};
        ar.names() = ar_names;
        acceptance_rates.push_back(ar);
        Rcpp::List sv;
        sv.push_back(Rcpp::wrap(p_var));
        sv.push_back(Rcpp::wrap(en_v));
        sv.push_back(Rcpp::wrap(ep_v));
        sv.push_back(Rcpp::wrap(r_v));
        sv.push_back(Rcpp::wrap(mr_v));
        sv.push_back(Rcpp::wrap(mc_v));
        Rcpp::StringVector sv_names{"allele_freq_var", "eps_neg_var", "eps_pos_var", "r_var", "m_r_var", "mean_coi_var"};
        sv.names() = sv_names;
        sampling_variances.push_back(sv);
    }

    Rcpp::List res;
    res.push_back(Rcpp::wrap(llik_burnin));
    res.push_back(Rcpp::wrap(llik_sample));
    res.push_back(Rcpp::wrap(prior_burnin));
    res.push_back(Rcpp::wrap(prior_sample));
    res.push_back(Rcpp::wrap(posterior_burnin));
    res.push_back(Rcpp::wrap(posterior_sample));
    res.push_back(Rcpp::wrap(by_sample(data_llik)));
    res.push_back(Rcpp::wrap(by_sample(coi)));
    res.push_back(Rcpp::wrap(lam));
    res.push_back(Rcpp::wrap(p_store));
    res.push_back(Rcpp::wrap(by_sample(eps_neg)));
    res.push_back(Rcpp::wrap(by_sample(eps_pos)));
    res.push_back(Rcpp::wrap(by_sample(rel)));
    res.push_back(Rcpp::wrap(latent_store));
    res.push_back(Rcpp::wrap(gd.observed_coi));
    res.push_back(Rcpp::wrap(swap_store));
    res.push_back(Rcpp::wrap(std::vector<double>(swap_acc.begin(), swap_acc.end())));
    res.push_back(Rcpp::wrap(swap_barriers));
    res.push_back(Rcpp::wrap(temp_gradient));
    res.push_back(Rcpp::wrap(acceptance_rates));
    res.push_back(Rcpp::wrap(sampling_variances));
    res.push_back(Rcpp::wrap(reached != 0));
    res.push_back(Rcpp::wrap(minutes));
    Rcpp::StringVector res_names{"llik_burnin", "llik_sample", "prior_burnin", "prior_sample", "posterior_burnin",
                                 "posterior_sample", "data_llik", "coi", "lam_coi", "allele_freqs", "eps_neg",
                                 "eps_pos", "relatedness", "latent_genotypes", "observed_coi", "swap_store",
                                 "swap_acceptances", "swap_barriers", "temp_gradient", "acceptance_rates",
                                 "sampling_variances", "max_runtime_reached", "total_runtime"};
    res.names() = res_names;
    return res;
}
