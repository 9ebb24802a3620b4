/* TEST INFRASTRUCTURE ONLY (oracle/).  CPU restatement of the reference's algorithm for moire's
 * likelihood hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this; the product (moire_b200/) never does and fails loudly without its CUDA library.
 *
 * Pinning: the reference's own tests hold NO golden vectors for this path, so the
 * restatement is pinned against outputs of the reference itself, compiled unmodified here
 * (oracle/_ref, see Makefile) -- tests/test_oracle_vs_ref.py when oracle/_ref is present, and
 * always against the committed fixtures tests/golden/*.npz minted from it by
 * tests/golden/make_golden.py.
 *
 * ABI: every floating-point value crosses as double; `prec` selects the arithmetic the
 * restatement runs in: 32 = float (what the reference runs), 64 = double (the reference's
 * mechanical float->double variant, which the CUDA engine follows).
 */
#ifndef MOIRE_ORACLE_H
#define MOIRE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_A 128  /* alleles per locus: one 64-bit mask word, or two (orc_eval_cells_wide) */
#define ORC_MAX_N 127  /* COI bound of the reference's lgamma LUT (src/sampler.cpp:22-25) */
#define ORC_K_IE_MAX 10 /* src/prob_any_missing.cpp:11 */

long orc_combinations(int n, int r, signed char *out, long max_rows,
                      unsigned long *num_combinations, unsigned long *generated);

double orc_pam_scalar(int prec, const double *q, int k, int n);
void orc_pam_vectorized(int prec, const double *q, int k, int n_max, double *out);
double orc_transmission(int prec, const int *idx, int k, const double *p, int A, int coi,
                        double relatedness, int allow_relatedness);
double orc_observation(int prec, const int *idx, int k, const int *obs, int A, double eps_neg,
                       double eps_pos, double max_eps_neg, double max_eps_pos);
double orc_latent_log_prob(int prec, int A, int obs_pos, int fn, int fp, int coi, double eps_pos,
                           double eps_neg);
void orc_constrained(int prec, double z, double curr, double lo, double hi, double *prop,
                     double *adj);
void orc_salt(int prec, const double *p, int A, int idx, double logit_prop, double *prop_p,
              double *log_adj);
double orc_logit(int prec, double x);
void orc_log_pq(int prec, double x, double *lp, double *lq);
double orc_logit_sum(int prec, const double *x, int n);
void orc_logit_scale(int prec, const double *x, int n, double scale, double *out);
double orc_expit(int prec, double x);
double orc_log_sum_exp(int prec, const double *x, int n);
double orc_dztpois(int prec, int x, double lambda);
double orc_dbinom(int prec, int x, int size, double prob);
double orc_beta_log_prior(int prec, double x, double a, double b);
double orc_gamma_log_prior(int prec, double x, double shape, double scale);

/* All N*L cells of one chain state on bit masks; layouts as the engine's (sample-major [N][L]).
 * Writes transmission and observation parts separately; cell = t + o (0 for missing cells). */
void orc_eval_cells(int prec, int N, int L, const int *num_alleles, const uint64_t *obs,
                    const uint8_t *is_missing, const uint64_t *latent, const int *m,
                    const double *r, const double *eps_neg, const double *eps_pos, const double *p,
                    int a_pad, int allow_relatedness, double max_eps_neg, double max_eps_pos,
                    double *t_out, double *o_out);

/* The same for allele sets of W 64-bit words, [N][L][W] (word 0 = alleles 0..63): what the two-word
 * build of the engine (libmoire_b200_w128.so, loci with 65..128 alleles) is checked against.  The
 * reference's functions take index vectors of any length (src/chain.cpp:858-961); the masks are
 * unpacked into exactly those. */
void orc_eval_cells_wide(int prec, int N, int L, int W, const int *num_alleles, const uint64_t *obs,
                         const uint8_t *is_missing, const uint64_t *latent, const int *m,
                         const double *r, const double *eps_neg, const double *eps_pos, const double *p,
                         int a_pad, int allow_relatedness, double max_eps_neg, double max_eps_pos,
                         double *t_out, double *o_out);

/* Philox4x32-10 of the stepper (independent of the engine's), for the known-answer test. */
void orc_philox4x32(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

/* Engine-semantics Metropolis stepper, one chain, one update kind per call (see
 * engine_stepper_impl.h).  `chain` is the struct stp_chain defined there (mirrored field by
 * field in oracle/pyoracle.py). */
struct stp_chain_s;
void orc_stp_coi_delta_many(uint64_t seed, int n, int *out);
void orc_stp_sample_latent_many(uint64_t seed, int n, uint64_t obs, int A, int coi, double eps_pos,
                                double eps_neg, uint64_t *masks, double *log_prob);
void orc_step_update(void *chain, int kind, int iteration);
void orc_eval_priors(void *chain);

#ifdef __cplusplus
}
#endif
#endif
