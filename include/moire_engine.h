/* moire_engine.h -- C ABI of the B200-native MCMC likelihood engine (libmoire_b200.so).
 *
 * This is the drop-in boundary for moire's data-parallel hot path: everything the reference's
 * parallel-tempering driver (src/mcmc.cpp) and result packer (src/main.cpp) do with a `Chain`
 * (src/chain.h:17-180) goes through these entry points.  One engine holds the device-resident
 * state of a contiguous block of tempered chains on ONE GPU; the ladder shards over GPUs by
 * creating one engine per rank with different `chain_offset`.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 on success
 * and a negative MOIRE_E* code otherwise (no exceptions cross the ABI); the caller owns every host
 * buffer; the engine owns all device memory behind the opaque handle; one host thread per handle.
 * There is NO CPU fallback: without a CUDA device `moire_engine_create` fails with
 * MOIRE_ECUDA and `moire_last_error()` says why.
 *
 * Layouts (sample-major, like the reference's genotyping_llik_new[sample * L + locus],
 * src/chain.cpp:1116):  per-cell arrays are [N][L]; per-sample arrays [N]; allele frequencies
 * [L][a_pad] with a_pad = moire_engine_a_pad().  Allele sets (observed and latent) are bit
 * masks, bit a = allele a (sorted index list of the reference == ascending bit order), of
 * W = moire_engine_mask_words() 64-bit words each, word 0 = alleles 0..63.
 *
 * Mask width.  The reference keeps allele sets as vectors of any length
 * (src/genotyping_data.cpp:17-48).  This ABI is built twice from the same sources:
 *   libmoire_b200.so       W = 1: up to 64 alleles per locus (every shape BASELINE.json names);
 *   libmoire_b200_w128.so  W = 2: up to 128 alleles per locus, flat pipelines only.
 * Same entry points, same structs; every `uint64_t *` mask array below is [..][W].  A caller picks
 * the library by the widest locus of its data (moire_b200/_lib.py does) and can ask either one with
 * moire_engine_mask_words() / moire_engine_max_alleles().
 */
#ifndef MOIRE_ENGINE_H
#define MOIRE_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOIRE_OK 0
#define MOIRE_EINVAL (-1)   /* bad argument / unsupported shape */
#define MOIRE_ECUDA (-2)    /* CUDA runtime error or no device */
#define MOIRE_ENOMEM (-3)
#define MOIRE_ESTATE (-4)   /* call order (e.g. step before init) */

#define MOIRE_MAX_ALLELES 64 /* alleles per locus PER MASK WORD: the limit is 64 * moire_engine_mask_words() */
#define MOIRE_MAX_COI 127    /* the reference's own bound: its lgamma table (src/sampler.cpp:22-25) */

/* Kernel generation: AUTO picks the flat pipelines (device-wide cost-class sort + persistent
 * evaluator) unless the problem is tiny (< 1e5 cells), where the fused per-tile kernels win on
 * latency.  Both produce the same chain; tests force each on one shape. */
#define MOIRE_PIPELINE_AUTO 0
#define MOIRE_PIPELINE_FLAT 1
#define MOIRE_PIPELINE_FUSED 2
/* OR-ed into moire_params::pipeline: emulate the one place where the shipped (float) reference and
 * an fp64 evaluation differ to leading order -- P(any missing) is rounded to float and clamped at
 * 1 - FLT_EPSILON (src/chain.cpp:888-893, src/prob_any_missing.cpp:99-112) instead of
 * 1 - DBL_EPSILON, which floors a cell's log(1 - PAM) at log(2^-23).  For attributing posterior
 * differences to the reference's precision; the default (off) is the fp64 evaluation whose cells
 * match the reference's double variant to 1e-9.  Process-wide: the last created engine decides. */
#define MOIRE_NUMERICS_REF32 0x100
/* OR-ed into moire_params::pipeline: update_p's PAM cache (flat pipelines only).  SALT
 * (src/mcmc_utils.h:252-300) moves one allele frequency and rescales the others by a common factor,
 * so a cell whose latent genotype does not hold the proposed allele keeps its normalised frequencies
 * and every P(any missing); with the cache ON such a cell re-runs only the relatedness mixture on
 * the new frequency sum, and only the cells that hold the proposed allele (a fraction k / A_j) are
 * evaluated from scratch each round.  Default (neither bit): OFF -- a cached vector was computed
 * from frequencies that have been rescaled since, equal to the last bits, and a cell whose coverage
 * probability is tiny amplifies that difference (observed: 2.6e-9 relative on a cell term, past the
 * 1e-9 the from-scratch evaluation holds).  Costs 16 B x max_coi per cell of device memory. */
#define MOIRE_P_CACHE_ON 0x200
#define MOIRE_P_CACHE_OFF 0x400

/* The 8 Metropolis updates in the order MCMC::burnin runs them (src/mcmc.cpp:162-173). */
enum moire_update_kind
{
    MOIRE_UPDATE_EPS_NEG = 0,  /* Chain::update_eps_neg   src/chain.cpp:632-695 */
    MOIRE_UPDATE_EPS_POS = 1,  /* Chain::update_eps_pos   src/chain.cpp:567-630 */
    MOIRE_UPDATE_P = 2,        /* Chain::update_p         src/chain.cpp:466-565 */
    MOIRE_UPDATE_M = 3,        /* Chain::update_m         src/chain.cpp:165-224 */
    MOIRE_UPDATE_R = 4,        /* Chain::update_r         src/chain.cpp:315-374 */
    MOIRE_UPDATE_EFF_COI = 5,  /* Chain::update_eff_coi   src/chain.cpp:226-313 */
    MOIRE_UPDATE_SAMPLES = 6,  /* Chain::update_samples   src/chain.cpp:697-807 */
    MOIRE_UPDATE_MEAN_COI = 7  /* Chain::update_mean_coi  src/chain.cpp:809-856 */
};

/* Replaces GenotypingData (src/genotyping_data.h:15-24), repacked SoA + bit masks. */
typedef struct moire_data
{
    int32_t num_samples;          /* N */
    int32_t num_loci;             /* L */
    const int32_t *num_alleles;   /* [L], 2..64 W */
    const uint64_t *obs_mask;     /* [N][L][W] observed alleles */
    const uint8_t *is_missing;    /* [N][L] 0/1 */
    const int32_t *observed_coi;  /* [N] or NULL (then derived as max_j popcount(obs_mask)) */
} moire_data;

/* POD mirror of the model half of Parameters (src/parameters.h:7-53, src/parameters.cpp:9-37). */
typedef struct moire_params
{
    int32_t allow_relatedness;
    int32_t burnin;  /* proposal-scale adaptation runs while iteration < burnin */
    int32_t max_coi;
    int32_t pipeline; /* MOIRE_PIPELINE_*: which kernel generation runs the updates (0 = by size) */
    double mean_coi_shape, mean_coi_scale;
    double max_eps_pos, eps_pos_alpha, eps_pos_beta;
    double max_eps_neg, eps_neg_alpha, eps_neg_beta;
    double r_alpha, r_beta;
} moire_params;

/* Full parameter state of one chain (public members of Chain, src/chain.h:84-142).
 * Any pointer may be NULL to skip that field.  On load, fields not given keep their value and
 * every likelihood/prior cache is rebuilt (Chain::initialize_likelihood, src/chain.cpp:1166-1223).
 * Output-only fields are ignored on load. */
typedef struct moire_chain_state
{
    int32_t *m;         /* [N] COI */
    double *r;          /* [N] relatedness */
    double *eps_pos;    /* [N] */
    double *eps_neg;    /* [N] */
    double *p;          /* [L][a_pad] allele frequencies (entries a >= A_j are 0) */
    uint64_t *latent;   /* [N][L][W] latent genotypes */
    double *lg_adj;     /* [N][L] log proposal density of the current latent genotype */
    double *mean_coi;   /* [1] */
    /* output only */
    double *cell_t;     /* [N][L] transmission term (calc_transmission_process) */
    double *cell_o;     /* [N][L] observation term (calc_observation_process); cell = t + o */
    double *prior_eps_neg, *prior_eps_pos, *prior_coi, *prior_r; /* [N] */
    double *mean_coi_hyper_prior;                                  /* [1] */
} moire_chain_state;

/* Acceptance counters and proposal scales returned at exit (src/main.cpp:99-136). */
typedef struct moire_chain_diagnostics
{
    int32_t *eps_neg_accept, *eps_pos_accept, *m_accept, *r_accept, *m_r_accept,
        *sample_accept;                                    /* [N] */
    double *eps_neg_var, *eps_pos_var, *r_var, *m_r_var; /* [N] proposal sds */
    int32_t *p_accept, *p_attempt;                       /* [L][a_pad] */
    double *p_prop_var;                                  /* [L][a_pad] */
    double *mean_coi_accept, *mean_coi_var;              /* [1] */
} moire_chain_diagnostics;

typedef struct moire_engine moire_engine;

/* Last error of a failed call that had no handle (create), or of `e`. */
const char *moire_last_error(void);
const char *moire_engine_last_error(const moire_engine *e);

/* Creates the engine for chains [chain_offset, chain_offset + num_chains) of a ladder, on CUDA
 * device `device_id`, with the given starting temperatures (Chain::Chain(..., temp),
 * src/chain.cpp:1294-1310).  `seed` keys every counter-based Philox stream; streams are keyed by
 * GLOBAL chain id, so results do not depend on how the ladder is sharded. */
int moire_engine_create(const moire_data *data, const moire_params *params, int32_t num_chains,
                        int32_t chain_offset, const double *temps, int32_t device_id,
                        uint64_t seed, moire_engine **out);
int moire_engine_destroy(moire_engine *e);

int32_t moire_engine_a_pad(const moire_engine *e);
/* Mask width of THIS library: 64-bit words per allele set (1 or 2) and the alleles a locus may have
 * (64 or 128).  The reference has no such limit (src/genotyping_data.cpp:17-48): see "Mask width". */
int32_t moire_engine_mask_words(void);
int32_t moire_engine_max_alleles(void);
int32_t moire_engine_num_chains(const moire_engine *e);

/* Chain::initialize_parameters with the retry loop of MCMC::MCMC (src/mcmc.cpp:68-104): draws a
 * starting state per chain and retries chains whose log-likelihood is not finite, up to
 * max_tries.  Per chain (arrays of num_chains, any may be NULL): number of ill-conditioned
 * attempts, whether it is still ill-conditioned, the 1-based (sample, locus) of the first
 * non-finite genotyping term of the LAST failed attempt (0,0 if none), and non-finite prior-term
 * counts [num_chains][5] in the order eps_neg, eps_pos, relatedness, coi, mean_coi_hyper
 * (Chain::first_nonfinite_genotyping_term / collect_nonfinite_prior_terms,
 * src/chain.cpp:1032-1081). */
int moire_engine_init_chains(moire_engine *e, int32_t max_tries, int32_t *ill_count,
                             int32_t *still_ill, int32_t *first_bad_sample,
                             int32_t *first_bad_locus, int32_t *prior_nonfinite);

/* Chain::first_nonfinite_genotyping_term and Chain::collect_nonfinite_prior_terms
 * (src/chain.cpp:1032-1081) of one chain's CURRENT state, whenever asked (init_chains calls the same
 * routine for every failed attempt): the 1-based (sample, locus) of the first cell, in the
 * reference's sample-major order, whose log-likelihood is not finite (0, 0 if none) and the counts
 * of non-finite prior terms [5] (eps_neg, eps_pos, relatedness, coi, mean_coi_hyper).  NULL = skip. */
int moire_engine_diagnose(moire_engine *e, int32_t chain, int32_t *first_bad_sample,
                          int32_t *first_bad_locus, int32_t *prior_nonfinite);

/* Every ill-conditioned start seen by the last moire_engine_init_chains, in attempt order:
 * records of (local chain, 1-based sample, 1-based locus), (chain, 0, 0) when no genotyping term
 * was non-finite (record_ill_conditioned_start, src/mcmc.cpp:25-39).  *count gets the number of
 * records; at most max_records are copied. */
int moire_engine_read_init_log(moire_engine *e, int32_t *records, int32_t max_records,
                               int32_t *count);

/* One sweep of all updates on every chain of the engine, in the reference's order, followed by
 * the per-chain log-likelihood / prior reduction: the body of the OpenMP loops in
 * MCMC::burnin / MCMC::sample (src/mcmc.cpp:159-174, 308-323).  `iteration` is what the
 * reference passes to update_* (step during burn-in, burnin + step while sampling). */
int moire_engine_step(moire_engine *e, int32_t iteration);
/* For a caller that replays steps without a host round trip (the driver captures a step together
 * with its all-gather and device-side swap pass into one CUDA graph): with iteration < 0 the
 * kernels of moire_engine_step read the iteration from the device word below, which the caller
 * advances on the device; a step issued while the engine's stream is being captured goes into the
 * caller's graph; moire_engine_captured_step_launches says how many kernel launches that step
 * was, and moire_engine_add_launches credits the launch counter for every replay (and for the
 * caller's own kernels that travel with a step). */
#define MOIRE_ITERATION_FROM_DEVICE (-1)
int moire_engine_device_iteration(moire_engine *e, int32_t **dev_iteration);
int moire_engine_device_temps(moire_engine *e, double **dev_temps); /* Chain::temp, [num_chains] */
int moire_engine_captured_step_launches(moire_engine *e, int64_t *launches);
int moire_engine_add_launches(moire_engine *e, int64_t n);
/* A single update kind on every chain (parity tests, profiling). */
int moire_engine_update(moire_engine *e, int32_t kind, int32_t iteration);
/* Recompute every cell and prior from the current parameters, then reduce. */
int moire_engine_eval_all(moire_engine *e);
/* Per-chain sums only (Chain::calc_new_likelihood / calc_new_prior, src/chain.cpp:978-1026). */
int moire_engine_reduce(moire_engine *e);

/* Chain::get_llik / get_prior for every chain of the engine (synchronises the stream). */
int moire_engine_get_scalars(moire_engine *e, double *llik, double *prior);
/* Device addresses of the same [num_chains] arrays, for a collective (NCCL all-gather) that
 * wants no host round trip.  Valid until destroy. */
int moire_engine_device_scalars(moire_engine *e, void **dev_llik, void **dev_prior);
/* Engines return their device buffers to a per-process free list when destroyed (the next engine
 * of the same shape re-uses them); this hands the list back to the driver. */
int moire_engine_trim_memory(void);
/* The CUDA stream (cudaStream_t) every call of this engine is ordered on, for work that must follow
 * a step without a host round trip (the driver's NCCL all-gather of the scalars). */
int moire_engine_stream(moire_engine *e, void **stream);
/* Chain::set_temp / get_temp for every chain. */
int moire_engine_set_temps(moire_engine *e, const double *temps);
int moire_engine_get_temps(moire_engine *e, double *temps);

int moire_engine_dump_state(moire_engine *e, int32_t chain, moire_chain_state *out);
int moire_engine_load_state(moire_engine *e, int32_t chain, const moire_chain_state *in);
/* chain < 0: every chain of the engine at once, each field [num_chains] times its per-chain size. */
int moire_engine_read_diagnostics(moire_engine *e, int32_t chain, moire_chain_diagnostics *out);
/* Chain::get_llik(sample) for every sample (src/chain.cpp:1083-1097): [N]. */
int moire_engine_read_sample_llik(moire_engine *e, int32_t chain, double *out);
/* What MCMC::sample stores of the hot chain (src/mcmc.cpp:327-347) in one call: p [L][a_pad],
 * m, eps_neg, eps_pos, r, per-sample llik [N], mean_coi[1], optional latent [N][L]. */
int moire_engine_record_draw(moire_engine *e, int32_t chain, double *p, int32_t *m,
                             double *eps_neg, double *eps_pos, double *r, double *sample_llik,
                             double *mean_coi, uint64_t *latent);

/* The same as an enqueue-only call for a device-resident store of many draws (DEVICE pointers,
 * draw-major like moire_mcmc_get_draws; `latent` may be NULL), gated and addressed on the device:
 * dev_sel[0] = chain of this engine to record or -1 for none, dev_sel[1] = slot < capacity.
 * Nothing is synchronised; the caller reads the store back whenever it likes. */
typedef struct moire_draw_store
{
    double *p;            /* [capacity][L][a_pad] */
    int32_t *m;           /* [capacity][N] */
    double *eps_neg, *eps_pos, *r, *sample_llik; /* [capacity][N] */
    double *mean_coi;     /* [capacity] */
    uint64_t *latent;     /* [capacity][N][L][W] or NULL */
    int32_t capacity;
} moire_draw_store;
int moire_engine_enqueue_record_draw(moire_engine *e, const int32_t *dev_sel,
                                     const moire_draw_store *store);

/* Work counters since create: non-missing cell likelihood evaluations, kernel launches. */
int moire_engine_counters(moire_engine *e, uint64_t *cell_evals, uint64_t *kernel_launches);
/* The same evaluation count split by update kind: [8], indexed by moire_update_kind. */
int moire_engine_kind_evals(moire_engine *e, uint64_t *evals);
/* A CUDA-event stopwatch on the engine's stream: start records an event, stop records a second
 * one, waits for it and returns the device time between the two in ms. */
int moire_engine_timer_start(moire_engine *e);
int moire_engine_timer_stop(moire_engine *e, double *ms);
/* Device time (ms) spent in each update kind (CUDA events) since create; [9]: the 8 kinds + the
 * reduction.  Enabled by moire_engine_set_timing(e, 1) (adds event records around each launch). */
int moire_engine_set_timing(moire_engine *e, int32_t on);
int moire_engine_get_timing(moire_engine *e, int32_t *on);
int moire_engine_read_timing(moire_engine *e, double *ms, uint64_t *launches);
/* Device time and launch count of the transmission evaluator kernel alone (k_flat_eval, the
 * dominant kernel; it runs inside update_p and the four per-sample updates that need
 * transmission terms), accumulated like moire_engine_read_timing. */
int moire_engine_read_evaluator_timing(moire_engine *e, double *ms, uint64_t *launches);
int moire_engine_synchronize(moire_engine *e);

/* The engine's subset enumeration for k latent alleles -- what replaces
 * CombinationIndicesGenerator (src/combination_indices_generator.cpp:5-76) as driven by
 * src/prob_any_missing.cpp:20-52, 186-214: sizes 1..k in turn, each in lexicographic order of the
 * index tuples.  Decoded ON THE DEVICE from one of the three tables the kernels read:
 * which = 0 the depth/tail table of the one-thread walk (k <= 10), 1 the full-mask table of the
 * warp-cooperative evaluator (k <= 10), 2 the compile-time order of the unrolled walks (k <= 6).
 * out[i], i < 2^k - 1: bits 0..k-1 = elements of the i-th subset, bit 15 = its term is subtracted.
 * Must match the reference's sequences bit for bit (tests/test_gpu_parity.py). */
int moire_engine_subset_order(moire_engine *e, int32_t which, int32_t k, uint16_t *out);

/* Synthetic genotyping data generated on the device: the per-cell part of the reference's
 * simulate_data (R/simulator.R:154-212) -- strains per cell, their alleles, the false-negative /
 * false-positive observation process and missingness -- given the per-locus allele frequencies
 * [L][a_pad], per-sample COIs and within-host relatedness (drawn by the caller: a few thousand
 * numbers).  Outputs are HOST arrays in the engine's input layout: obs_mask [N][L], is_missing
 * [N][L], and (optional) true_mask [N][L] = the alleles actually present.  Streams are keyed by
 * (seed, sample, locus): the result does not depend on the launch shape. */
int moire_simulate_data(int32_t num_samples, int32_t num_loci, const int32_t *num_alleles,
                        const double *allele_freqs, int32_t a_pad, const int32_t *sample_cois,
                        const double *sample_relatedness, double epsilon_pos, double epsilon_neg,
                        double missingness, uint64_t seed, int32_t device_id, uint64_t *obs_mask,
                        uint8_t *is_missing, uint64_t *true_mask);

/* Philox4x32-10 block (counter, key) -> 4 words; exported so tests can pin the generator against
 * the Random123 known-answer vectors. */
void moire_philox4x32(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* MOIRE_ENGINE_H */
