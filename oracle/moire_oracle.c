/* TEST INFRASTRUCTURE ONLY (oracle/).  See moire_oracle.h for the rules and the pinning status. */
#define _GNU_SOURCE
#include "moire_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- a6: CombinationIndicesGenerator (src/combination_indices_generator.{h,cpp}) -------------
 * Lexicographic r-subsets of {0..n-1}; `completed`, `generated`, `num_combinations` as the
 * reference maintains them. */
typedef struct
{
    int completed;
    unsigned long generated;
    unsigned long num_combinations;
    signed char curr[ORC_MAX_A];
    int n, r;
} orc_comb;

static void orc_comb_count(orc_comb *g)
{
    /* calculateNumCombinations, src/combination_indices_generator.cpp:50-76 */
    int tmp = g->r;
    if (tmp > g->n)
    {
        g->num_combinations = 0;
        return;
    }
    if (tmp * 2 > g->n) tmp = g->n - tmp;
    if (tmp == 0)
    {
        g->num_combinations = 1;
        return;
    }
    g->num_combinations = (unsigned long)g->n;
    for (int i = 2; i <= tmp; ++i)
    {
        g->num_combinations *= (unsigned long)(g->n - i + 1);
        g->num_combinations /= (unsigned long)i;
    }
}

static void orc_comb_reset(orc_comb *g, int n, int r)
{
    /* reset, src/combination_indices_generator.cpp:14-25 */
    g->completed = (n < 1) || (r > n) || (r == 0);
    g->generated = 1;
    g->n = n;
    g->r = r;
    for (int i = 0; i < r && i < ORC_MAX_A; ++i) g->curr[i] = (signed char)i;
    orc_comb_count(g);
}

static void orc_comb_next(orc_comb *g)
{
    /* next, src/combination_indices_generator.cpp:27-42 */
    g->completed = 1;
    const int lim = g->n - g->r;
    for (int i = g->r - 1; i >= 0; --i)
    {
        if (g->curr[i] < lim + i)
        {
            signed char j = (signed char)(g->curr[i] + 1);
            while (i < g->r) g->curr[i++] = j++;
            g->completed = 0;
            g->generated++;
            break;
        }
    }
}

long orc_combinations(int n, int r, signed char *out, long max_rows,
                      unsigned long *num_combinations, unsigned long *generated)
{
    orc_comb g;
    orc_comb_reset(&g, n, r);
    long rows = 0;
    while (!g.completed && rows < max_rows)
    {
        for (int i = 0; i < r; ++i) out[rows * r + i] = g.curr[i];
        ++rows;
        orc_comb_next(&g);
    }
    *num_combinations = g.num_combinations;
    *generated = g.generated;
    return rows;
}

/* ---- the two precisions ------------------------------------------------------------------- */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define F(name) CAT(name, _f32)
#define REAL_EPSILON FLT_EPSILON
#define LOG logf
#define LOG1P log1pf
#define EXP expf
#define EXPM1 expm1f
#define FABS fabsf
#include "moire_oracle_impl.h"
#undef REAL
#undef F
#undef REAL_EPSILON
#undef LOG
#undef LOG1P
#undef EXP
#undef EXPM1
#undef FABS

#define REAL double
#define F(name) CAT(name, _f64)
#define REAL_EPSILON DBL_EPSILON
#define LOG log
#define LOG1P log1p
#define EXP exp
#define EXPM1 expm1
#define FABS fabs
#include "moire_oracle_impl.h"
#undef REAL
#undef F
#undef REAL_EPSILON
#undef LOG
#undef LOG1P
#undef EXP
#undef EXPM1
#undef FABS

/* ---- double-ABI exports ---------------------------------------------------------------------- */
#define TO_F32(dst, src, n)                 \
    float dst[ORC_MAX_N + ORC_MAX_A + 1];   \
    for (int i_ = 0; i_ < (n); ++i_) dst[i_] = (float)(src)[i_]

double orc_pam_scalar(int prec, const double *q, int k, int n)
{
    if (prec == 64) return pam_scalar_f64(q, k, n);
    TO_F32(qf, q, k);
    return pam_scalar_f32(qf, k, n);
}

void orc_pam_vectorized(int prec, const double *q, int k, int n_max, double *out)
{
    if (prec == 64)
    {
        pam_vectorized_f64(q, k, n_max, out);
        return;
    }
    TO_F32(qf, q, k);
    float of[ORC_MAX_N + 1];
    pam_vectorized_f32(qf, k, n_max, of);
    for (int i = 0; i < n_max; ++i) out[i] = of[i];
}

double orc_transmission(int prec, const int *idx, int k, const double *p, int A, int coi,
                        double relatedness, int allow_relatedness)
{
    if (prec == 64) return transmission_f64(idx, k, p, coi, relatedness, allow_relatedness);
    TO_F32(pf, p, A);
    return transmission_f32(idx, k, pf, coi, (float)relatedness, allow_relatedness);
}

double orc_observation(int prec, const int *idx, int k, const int *obs, int A, double eps_neg,
                       double eps_pos, double max_eps_neg, double max_eps_pos)
{
    if (prec == 64)
        return observation_f64(idx, k, obs, A, eps_neg, eps_pos, max_eps_neg, max_eps_pos);
    return observation_f32(idx, k, obs, A, (float)eps_neg, (float)eps_pos, (float)max_eps_neg,
                           (float)max_eps_pos);
}

double orc_latent_log_prob(int prec, int A, int obs_pos, int fn, int fp, int coi, double eps_pos,
                           double eps_neg)
{
    if (prec == 64) return latent_log_prob_f64(A, obs_pos, fn, fp, coi, eps_pos, eps_neg);
    return latent_log_prob_f32(A, obs_pos, fn, fp, coi, (float)eps_pos, (float)eps_neg);
}

void orc_constrained(int prec, double z, double curr, double lo, double hi, double *prop,
                     double *adj)
{
    if (prec == 64)
    {
        constrained_f64(z, curr, lo, hi, prop, adj);
        return;
    }
    float pf, af;
    constrained_f32((float)z, (float)curr, (float)lo, (float)hi, &pf, &af);
    *prop = pf;
    *adj = af;
}

void orc_salt(int prec, const double *p, int A, int idx, double logit_prop, double *prop_p,
              double *log_adj)
{
    if (prec == 64)
    {
        salt_f64(p, A, idx, logit_prop, prop_p, log_adj);
        return;
    }
    TO_F32(pf, p, A);
    float of[ORC_MAX_A], adj;
    salt_f32(pf, A, idx, (float)logit_prop, of, &adj);
    for (int a = 0; a < A; ++a) prop_p[a] = of[a];
    *log_adj = adj;
}

double orc_logit(int prec, double x) { return prec == 64 ? logit_f64(x) : logit_f32((float)x); }

void orc_log_pq(int prec, double x, double *lp, double *lq)
{
    if (prec == 64)
    {
        log_pq_f64(x, lp, lq);
        return;
    }
    float a, b;
    log_pq_f32((float)x, &a, &b);
    *lp = a;
    *lq = b;
}

double orc_logit_sum(int prec, const double *x, int n)
{
    if (prec == 64) return logit_sum_f64(x, n);
    TO_F32(xf, x, n);
    return logit_sum_f32(xf, n);
}

void orc_logit_scale(int prec, const double *x, int n, double scale, double *out)
{
    if (prec == 64)
    {
        logit_scale_f64(x, n, scale, out);
        return;
    }
    TO_F32(xf, x, n);
    float of[ORC_MAX_A];
    logit_scale_f32(xf, n, (float)scale, of);
    for (int i = 0; i < n; ++i) out[i] = of[i];
}

double orc_expit(int prec, double x) { return prec == 64 ? expit_f64(x) : expit_f32((float)x); }

double orc_log_sum_exp(int prec, const double *x, int n)
{
    if (prec == 64) return log_sum_exp_f64(x, n);
    TO_F32(xf, x, n);
    return log_sum_exp_f32(xf, n);
}

double orc_dztpois(int prec, int x, double lambda)
{
    return prec == 64 ? dztpois_f64(x, lambda) : dztpois_f32(x, (float)lambda);
}

double orc_dbinom(int prec, int x, int size, double prob)
{
    return prec == 64 ? dbinom_f64(x, size, prob) : dbinom_f32(x, size, (float)prob);
}

double orc_beta_log_prior(int prec, double x, double a, double b)
{
    return prec == 64 ? dbeta_log_f64(x, a, b) : dbeta_log_f32((float)x, (float)a, (float)b);
}

double orc_gamma_log_prior(int prec, double x, double shape, double scale)
{
    return prec == 64 ? dgamma_log_f64(x, shape, scale)
                      : dgamma_log_f32((float)x, (float)shape, (float)scale);
}

void orc_eval_cells(int prec, int N, int L, const int *num_alleles, const uint64_t *obs,
                    const uint8_t *is_missing, const uint64_t *latent, const int *m,
                    const double *r, const double *eps_neg, const double *eps_pos, const double *p,
                    int a_pad, int allow_relatedness, double max_eps_neg, double max_eps_pos,
                    double *t_out, double *o_out)
{
    for (int i = 0; i < N; ++i)
    {
        for (int j = 0; j < L; ++j)
        {
            const size_t cell = (size_t)i * (size_t)L + (size_t)j;
            const int A = num_alleles[j];
            const double *pj = p + (size_t)j * (size_t)a_pad;
            if (prec == 64)
            {
                cell_llik_f64(latent[cell], obs[cell], A, is_missing[cell], pj, m[i], r[i],
                              eps_neg[i], eps_pos[i], allow_relatedness, max_eps_neg, max_eps_pos,
                              &t_out[cell], &o_out[cell]);
            }
            else
            {
                float pf[ORC_MAX_A], t, o;
                for (int a = 0; a < A; ++a) pf[a] = (float)pj[a];
                cell_llik_f32(latent[cell], obs[cell], A, is_missing[cell], pf, m[i], (float)r[i],
                              (float)eps_neg[i], (float)eps_pos[i], allow_relatedness,
                              (float)max_eps_neg, (float)max_eps_pos, &t, &o);
                t_out[cell] = t;
                o_out[cell] = o;
            }
        }
    }
}

void orc_eval_cells_wide(int prec, int N, int L, int W, const int *num_alleles, const uint64_t *obs,
                         const uint8_t *is_missing, const uint64_t *latent, const int *m,
                         const double *r, const double *eps_neg, const double *eps_pos, const double *p,
                         int a_pad, int allow_relatedness, double max_eps_neg, double max_eps_pos,
                         double *t_out, double *o_out)
{
    for (int i = 0; i < N; ++i)
    {
        for (int j = 0; j < L; ++j)
        {
            const size_t cell = (size_t)i * (size_t)L + (size_t)j;
            const int A = num_alleles[j];
            const double *pj = p + (size_t)j * (size_t)a_pad;
            t_out[cell] = o_out[cell] = 0.0;
            if (is_missing[cell]) continue;
            int idx[ORC_MAX_A], ov[ORC_MAX_A], k = 0;
            for (int a = 0; a < A && a < 64 * W; ++a)
            {
                if (latent[cell * (size_t)W + (size_t)(a >> 6)] >> (a & 63) & 1ULL) idx[k++] = a;
                ov[a] = (int)(obs[cell * (size_t)W + (size_t)(a >> 6)] >> (a & 63) & 1ULL);
            }
            if (prec == 64)
            {
                t_out[cell] = transmission_f64(idx, k, pj, m[i], r[i], allow_relatedness);
                o_out[cell] = observation_f64(idx, k, ov, A, eps_neg[i], eps_pos[i], max_eps_neg, max_eps_pos);
            }
            else
            {
                float pf[ORC_MAX_A];
                for (int a = 0; a < A; ++a) pf[a] = (float)pj[a];
                t_out[cell] = transmission_f32(idx, k, pf, m[i], (float)r[i], allow_relatedness);
                o_out[cell] = observation_f32(idx, k, ov, A, (float)eps_neg[i], (float)eps_pos[i],
                                              (float)max_eps_neg, (float)max_eps_pos);
            }
        }
    }
}

/* ---- the engine-semantics Metropolis stepper (fp64) ------------------------------------------ */
#include "engine_stepper_impl.h"
