/* moire_mcmc.h -- C ABI of the host-side parallel-tempering driver in libmoire_b200.so.
 *
 * This is the reference's `MCMC` class (src/mcmc.h:16-61, src/mcmc.cpp) and the loop of
 * `run_mcmc(Rcpp::List)` (src/main.cpp:48-93) re-pointed at the CUDA engine of moire_engine.h:
 * the OpenMP loop over `Chain` objects collapses into one `moire_engine_step` per MCMC step, and
 * what stays on the host is exactly what the reference keeps on its master thread -- the
 * even/odd temperature swaps (MCMC::swap_chains, src/mcmc.cpp:190-240), the adaptive ladder
 * (MCMC::adapt_temp, src/mcmc.cpp:244-304), the traces and the stores of the temp = 1 chain
 * (src/mcmc.cpp:175-177, 327-355).
 *
 * Where the ladder lives.  With an engine and a device-side exchange (world_size == 1, or the
 * NCCL exchange of moire_mcmc_init_nccl) the swap pass, its accumulators, the traces and the
 * stored draws live ON THE DEVICE (csrc/ladder.cuh): moire_mcmc_burnin / moire_mcmc_sample only
 * enqueue -- one CUDA-graph launch per step holding the engine's kernels, the all-gather and the
 * one-warp swap pass -- and the host reads results back where MCMC::adapt_temp is due (every
 * temp_adapt_steps steps; the spline stays on the host) or a getter asks.  Otherwise (ladder-only
 * mode, a callback exchange, MOIRE_B200_HOST_LADDER=1) the pass runs on the host after a per-step
 * read-back.  Same arithmetic (one swap_pair() for both), same results.
 *
 * Sharding: the ladder of T chains is cut into `world_size` contiguous blocks, one driver (and
 * one engine, one GPU) per rank.  Chains never move; a swap exchanges TEMPERATURES, as in the
 * reference (src/mcmc.cpp:217-219).  The only exchange per step is an all-gather of every
 * chain's scalar (llik, prior), done by a caller-supplied function (NCCL / gloo through
 * torch.distributed in this repo); every rank then runs the identical swap pass from a shared
 * counter-based uniform stream, so all ranks hold the same ladder at all times.
 *
 * Conventions as in moire_engine.h: 0 on success, negative MOIRE_E* otherwise; caller owns host
 * buffers; one host thread per handle.
 */
#ifndef MOIRE_MCMC_H
#define MOIRE_MCMC_H

#include <stdint.h>

#include "moire_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

/* The MCMC half of Parameters (src/parameters.cpp:9-24). */
typedef struct moire_mcmc_params
{
    int32_t thin;
    int32_t burnin;
    int32_t samples;
    int32_t adapt_temp;
    int32_t pre_adapt_steps;
    int32_t temp_adapt_steps;
    int32_t max_initialization_tries;
    int32_t record_latent_genotypes;
    double max_runtime; /* minutes; negative or non-finite: unlimited (0 stops at once, like the
                           reference's `timer.elapsed() < max_runtime`, src/main.cpp:54) */
} moire_mcmc_params;

typedef struct moire_mcmc moire_mcmc;

/* All-gather of `count` doubles per rank: recv[r * count + i] = send_r[i].  Return 0 on success.
 * With world_size == 1 no function is needed. */
typedef int (*moire_allgather_fn)(void *ctx, const double *send, int32_t count, double *recv);
/* Called once for every step of moire_mcmc_run (Rcpp::checkUserInterrupt + the progress bar,
 * src/main.cpp:56-69): phase 0 = burn-in, 1 = sampling; `posterior` and `hot_chain` are that
 * step's.  Return non-zero to stop the run.  With the ladder on the device (below) steps are
 * queued and read back in batches of up to 32, so the calls for a batch arrive together after its
 * last step, and a stop request (like max_runtime) takes effect at the end of a batch -- on
 * several ranks, at the end of the batch in which every rank has seen the request. */
typedef int (*moire_progress_fn)(void *ctx, int32_t phase, int32_t step, double posterior,
                                 int32_t hot_chain);

/* MCMC::MCMC up to (not including) chain initialisation.  `pt_chains` are the T starting
 * temperatures in ladder order (Parameters::pt_chains); `model->burnin` is overwritten with
 * `mcmc->burnin`.  This rank's chains are [first, first + count) with the block partition
 * first = rank * T / world_size (integer arithmetic, see moire_mcmc_partition).
 * device_id < 0 creates the driver in LADDER-ONLY mode, without an engine: the swap / adaptation /
 * trace logic runs on scalars injected with moire_mcmc_set_local_scalars and nothing is
 * evaluated.  That mode exists for host-logic and multi-process CPU tests; it is not a CPU
 * implementation of the likelihood path. */
int moire_mcmc_create(const moire_data *data, const moire_params *model,
                      const moire_mcmc_params *mcmc, const double *pt_chains, int32_t T,
                      int32_t rank, int32_t world_size, int32_t device_id, uint64_t seed,
                      moire_mcmc **out);
int moire_mcmc_destroy(moire_mcmc *h);
const char *moire_mcmc_last_error(const moire_mcmc *h);
/* The block of chains rank `rank` of `world_size` owns in a ladder of T. */
void moire_mcmc_partition(int32_t T, int32_t rank, int32_t world_size, int32_t *first,
                          int32_t *count);
int moire_mcmc_set_allgather(moire_mcmc *h, moire_allgather_fn fn, void *ctx);
/* The native exchange: an ncclAllGather enqueued on the engine's stream straight from the device
 * sums (no host staging, no callback in the step loop).  Rank 0 makes the 128-byte id with
 * moire_mcmc_nccl_unique_id and the embedding code hands it to every rank (any broadcast it has),
 * then every rank calls moire_mcmc_init_nccl (collective).  `library_path` names libnccl.so.2
 * (NULL: the one already loaded in the process, else the loader's search path).  Replaces
 * moire_mcmc_set_allgather; world_size == 1 needs neither.  The communicator lives as long as
 * the process; id == NULL re-uses it for another driver of the same (rank, world_size). */
int moire_mcmc_nccl_unique_id(char id[128], const char *library_path);
int moire_mcmc_init_nccl(moire_mcmc *h, const char id[128], const char *library_path);
/* The engine of this rank's chains (borrowed; destroyed with the driver). */
moire_engine *moire_mcmc_engine(moire_mcmc *h);

/* Chain construction with the retry loop (src/mcmc.cpp:68-104).  *failed is set if any chain of
 * ANY rank has no finite starting log-likelihood (MCMC::initialization_failed).  The per-attempt
 * failure records of this rank's chains are read with moire_engine_read_init_log. */
int moire_mcmc_init(moire_mcmc *h, int32_t *failed);
/* Per chain of THIS rank after moire_mcmc_init: number of ill-conditioned starts [local_chains]
 * (InitializationDiagnostics::ill_conditioned_per_chain) and non-finite prior-term counts
 * [local_chains][5] in the order eps_neg, eps_pos, relatedness, coi, mean_coi_hyper
 * (collect_nonfinite_prior_terms, src/chain.cpp:1050-1081).  NULL = skip. */
int moire_mcmc_get_init_counts(moire_mcmc *h, int32_t *ill_count, int32_t *prior_nonfinite);

/* MCMC::burnin(step) / MCMC::sample(step) (src/mcmc.cpp:157-188, 306-355): every update on every
 * chain, traces, swap pass, ladder adaptation; sample() also stores the draw of the temp = 1
 * chain on the rank that owns it. */
int moire_mcmc_burnin(moire_mcmc *h, int32_t step);
int moire_mcmc_sample(moire_mcmc *h, int32_t step);
/* MCMC::finalize (src/mcmc.cpp:357-363). */
int moire_mcmc_finalize(moire_mcmc *h);
/* The two loops of run_mcmc (src/main.cpp:52-93) incl. max_runtime, then finalize.
 * *max_runtime_reached and *runtime_minutes may be NULL. */
int moire_mcmc_run(moire_mcmc *h, moire_progress_fn progress, void *ctx,
                   int32_t *max_runtime_reached, double *runtime_minutes);

/* Where the ladder of this driver lives (1 = on the device, 0 = on the host) and whether its steps
 * replay a captured CUDA graph (0 before the first step, or when a library in the step -- the
 * NCCL exchange -- refused capture and the step goes out as plain launches). */
int moire_mcmc_get_mode(moire_mcmc *h, int32_t *device_ladder, int32_t *step_graph);

/* ---- results (src/main.cpp:138-189) ---------------------------------------------------------- */
typedef struct moire_mcmc_sizes
{
    int32_t num_chains, first_chain, local_chains;
    int32_t burnin_steps, sample_steps; /* steps actually run */
    int32_t num_draws;                  /* draws stored on THIS rank */
    int32_t num_swap_store;
    int32_t a_pad;
} moire_mcmc_sizes;
int moire_mcmc_get_sizes(moire_mcmc *h, moire_mcmc_sizes *out);
/* Traces of the chain at ladder position 0: [burnin_steps] / [sample_steps] each; any NULL
 * skipped. */
int moire_mcmc_get_traces(moire_mcmc *h, double *llik_burnin, double *prior_burnin,
                          double *posterior_burnin, double *llik_sample, double *prior_sample,
                          double *posterior_sample);
/* Ladder state: swap_store [num_swap_store], swap_acceptances [T-1], swap_barriers [T-1],
 * temp_gradient [T], swap_indices [T] (ladder position -> chain id), chain_temps [T]. */
int moire_mcmc_get_ladder(moire_mcmc *h, int32_t *swap_store, int64_t *swap_acceptances,
                          double *swap_barriers, double *temp_gradient, int32_t *swap_indices,
                          double *chain_temps);
/* Chain::get_hot of every chain, by chain id [T] (src/chain.h:176-177; see the quirk noted at
 * src/mcmc.cpp:221-225, 294: no chain is hot before the first accepted swap at ladder position 0
 * or the first adapt_temp). */
int moire_mcmc_get_hot(moire_mcmc *h, int32_t *hot);
/* The draws stored on this rank, draw-major: draw_step [D] (sampling step of each draw),
 * allele_freqs [D][L][a_pad], coi [D][N], eps_neg/eps_pos/relatedness/data_llik [D][N],
 * lam_coi [D], latent [D][N][L][W] mask words, W = moire_engine_mask_words() (only with
 * record_latent_genotypes).  NULL = skip. */
int moire_mcmc_get_draws(moire_mcmc *h, int32_t *draw_step, double *allele_freqs, int32_t *coi,
                         double *eps_neg, double *eps_pos, double *relatedness,
                         double *data_llik, double *lam_coi, uint64_t *latent);

/* ---- pieces exposed for parity tests against the reference's MCMC ----------------------------- */
/* One swap pass on explicit inputs (all T chains' llik, by chain id). */
int moire_mcmc_swap_pass(moire_mcmc *h, const double *chain_llik, int32_t step,
                         int32_t is_burnin);
/* Overrides the uniforms of the next swap pass (one per candidate pair, in pair order) so that a
 * test can feed the reference's R::runif draws; n = 0 restores the Philox stream. */
int moire_mcmc_set_swap_uniforms(moire_mcmc *h, const double *u, int32_t n);
/* Overrides ladder bookkeeping (tests): any NULL keeps its value. */
int moire_mcmc_set_ladder(moire_mcmc *h, const double *temp_gradient, const double *swap_barriers,
                          const int32_t *swap_indices, const double *chain_temps,
                          int64_t num_swaps, int32_t even_swap);
int moire_mcmc_adapt_temp(moire_mcmc *h);
/* Ladder-only mode: the (llik, prior) of this rank's chains that the next step will report. */
int moire_mcmc_set_local_scalars(moire_mcmc *h, const double *llik, const double *prior);
/* The spline behind adapt_temp (tk::spline cspline + make_monotonic + solve_one,
 * src/include/spline/src/spline.h:235-514, 676-720): x strictly increasing, n >= 3; returns the
 * first root of spline(x) = y or NaN. */
double moire_monotone_spline_solve(const double *x, const double *y, int32_t n, double y_target);

#ifdef __cplusplus
}
#endif
#endif /* MOIRE_MCMC_H */
