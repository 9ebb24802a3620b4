/* TEST INFRASTRUCTURE ONLY (oracle/).  Included twice by moire_oracle.c, once with
 * REAL=float (the arithmetic the reference actually runs in) and once with REAL=double
 * (the reference's mechanical double variant, which the CUDA engine follows).
 *
 * Plain-C restatement of the pure functions on moire's likelihood hot path.  Every function
 * cites the reference lines it follows (paths relative to /root/reference).  Operation ORDER
 * is kept identical to the reference wherever rounding depends on it (inclusion-exclusion,
 * EGF recurrence, sequential sums), so that in either precision the results are expected to
 * be bit-identical to the compiled reference up to libm (log/exp/log1p) differences.
 */

#ifndef REAL
#error "include from moire_oracle.c"
#endif

/* ---- a3: inclusion-exclusion, scalar n (src/prob_any_missing.cpp:13-53) -------------------- */
static REAL F(ie_scalar)(const REAL *q, int k, int n)
{
    orc_comb g;
    REAL bases[1024];
    REAL prob = (REAL)0;
    int sign = -1;
    for (int i = 1; i <= k; ++i)
    {
        sign = -sign;
        orc_comb_reset(&g, k, i);
        int nb = 0;
        while (!g.completed)
        {
            REAL base = (REAL)1;
            for (int t = 0; t < i; ++t) base -= q[(int)g.curr[t]];
            bases[nb++] = base;
            orc_comb_next(&g);
        }
        for (int b = 0; b < nb; ++b)
        {
            REAL base = bases[b];
            REAL r = (REAL)sign;
            int mc = n;
            while (mc > 0)
            {
                if (mc & 1) r *= base;
                base = base * base;
                mc >>= 1;
            }
            prob += r;
        }
    }
    return prob;
}

/* ---- a5: EGF coverage recurrence, always double (src/prob_any_missing.cpp:57-97) ---------- */
static void F(egf_coverage)(const REAL *q, int k, int max_n, double *a /* max_n+1 */)
{
    double b[ORC_MAX_N + 1];
    double ppow[ORC_MAX_N + 1];
    for (int m = 0; m <= max_n; ++m) a[m] = 0.0;
    a[0] = 1.0;
    for (int al = 0; al < k; ++al)
    {
        const double p = (double)q[al];
        for (int m = 0; m <= max_n; ++m) b[m] = 0.0;
        ppow[0] = 0.0;
        if (max_n >= 1) ppow[1] = p;
        for (int j = 2; j <= max_n; ++j) ppow[j] = ppow[j - 1] * p;
        for (int m = 1; m <= max_n; ++m)
        {
            double comb = (double)m;
            double sum = 0.0;
            for (int j = 1; j <= m; ++j)
            {
                sum += comb * a[m - j] * ppow[j];
                comb *= (double)(m - j) / (double)(j + 1);
            }
            b[m] = sum;
        }
        for (int m = 0; m <= max_n; ++m) a[m] = b[m];
    }
}

/* src/prob_any_missing.cpp:99-112 */
static REAL F(pam_from_coverage)(double coverage)
{
    const double pam = 1.0 - coverage;
    if (pam <= 0.0) return (REAL)0;
    if (pam >= 1.0) return (REAL)1;
    return (REAL)pam;
}

/* ---- a3: probAnyMissingFunctor::operator() (src/prob_any_missing.cpp:131-154) -------------- */
static REAL F(pam_scalar)(const REAL *q, int k, int n)
{
    if (n < k) return (REAL)1;
    if (k == 0) return (REAL)0;
    if (n == k)
    {
        /* exact_pam_n_equals_k, src/prob_any_missing.cpp:114-127 */
        double cov = 1.0;
        for (int i = 0; i < k; ++i) cov *= q[i];
        for (int i = 2; i <= k; ++i) cov *= i;
        return F(pam_from_coverage)(cov);
    }
    if (k <= ORC_K_IE_MAX) return F(ie_scalar)(q, k, n);
    double a[ORC_MAX_N + 1];
    F(egf_coverage)(q, k, n, a);
    return F(pam_from_coverage)(a[n]);
}

/* ---- a4: probAnyMissingFunctor::vectorized(q, 1, n_max) (src/prob_any_missing.cpp:162-234) - */
static void F(pam_vectorized)(const REAL *q, int k, int n_max, REAL *out /* n_max */)
{
    for (int j = 0; j < n_max; ++j) out[j] = (REAL)0;
    if (n_max < k)
    {
        for (int j = 0; j < n_max; ++j) out[j] = (REAL)1;
        return;
    }
    if (k <= ORC_K_IE_MAX)
    {
        orc_comb g;
        REAL bases[1024];
        for (int j = 0; j < k - 1; ++j) out[j] = (REAL)1;
        int sign = -1;
        for (int i = 1; i <= k; ++i)
        {
            sign = -sign;
            orc_comb_reset(&g, k, i);
            int nb = 0;
            for (unsigned long c = 0; c < g.num_combinations; ++c)
            {
                REAL base = (REAL)1;
                for (int t = 0; t < i; ++t) base -= q[(int)g.curr[t]];
                bases[nb++] = base;
                orc_comb_next(&g);
            }
            for (int b = 0; b < nb; ++b)
            {
                const REAL base = bases[b];
                REAL r = (REAL)sign;
                for (int j = 0; j < k - 1; ++j) r *= base;
                for (int j = k - 1; j < n_max; ++j)
                {
                    r *= base;
                    out[j] += r;
                }
            }
        }
        return;
    }
    double a[ORC_MAX_N + 1];
    F(egf_coverage)(q, k, n_max, a);
    for (int n = 1; n <= n_max; ++n)
        out[n - 1] = (n < k) ? (REAL)1 : F(pam_from_coverage)(a[n]);
}

/* ---- Sampler::dbinom with its lgamma LUT (src/sampler.cpp:22-25,44-50) --------------------- */
static REAL F(lgam)(int i) { return (REAL)lgamma((double)(i + 1)); }

static REAL F(dbinom)(int x, int size, REAL prob)
{
    const REAL log_p = x * LOG(prob) + (size - x) * LOG(1 - prob);
    const REAL log_coef = F(lgam)(size) - F(lgam)(x) - F(lgam)(size - x);
    return log_p + log_coef;
}

/* Sampler::dztpois (src/sampler.cpp:38-42) */
static REAL F(dztpois)(int x, REAL lambda)
{
    return x * LOG(lambda) - LOG(EXP(lambda) - 1) - F(lgam)(x);
}

/* clamp_pam lambda (src/chain.cpp:890-893) */
static REAL F(clamp_pam)(REAL x)
{
    const REAL max_pam = (REAL)1 - REAL_EPSILON;
    const REAL lo = x > (REAL)0 ? x : (REAL)0; /* std::max(0.f, x) */
    return max_pam < lo ? max_pam : lo;         /* std::min(max_pam, .) */
}

/* UtilFunctions::logSumExp(vector) (src/mcmc_utils.h:382-397) */
static REAL F(log_sum_exp)(const REAL *x, int n)
{
    REAL mx = x[0];
    for (int i = 1; i < n; ++i)
        if (mx < x[i]) mx = x[i];
    if (mx == -(REAL)INFINITY) return -(REAL)INFINITY;
    REAL sum = (REAL)0;
    for (int i = 0; i < n; ++i) sum += EXP(x[i] - mx);
    return mx + LOG(sum);
}

/* ---- a2: Chain::calc_transmission_process (src/chain.cpp:858-921) -------------------------- */
static REAL F(transmission)(const int *idx, int k, const REAL *p, int coi, REAL relatedness,
                            int allow_relatedness)
{
    REAL q[ORC_MAX_A];
    REAL total = (REAL)0;
    for (int j = 0; j < k; ++j)
    {
        q[j] = p[idx[j]];
        total += q[j];
    }
    const REAL log_total = LOG(total);
    for (int j = 0; j < k; ++j) q[j] = q[j] / total;

    if (!allow_relatedness)
    {
        const REAL pam = F(clamp_pam)(F(pam_scalar)(q, k, coi));
        return LOG1P(-pam) + log_total * coi;
    }
    REAL pam_vec[ORC_MAX_N + 1];
    REAL res[ORC_MAX_N + 1];
    F(pam_vectorized)(q, k, coi, pam_vec);
    int nres = 0;
    /* the reference's bound `i <= coi - total_alleles` is unsigned; k <= coi is an invariant */
    for (int i = 0; i <= coi - k; ++i)
    {
        const REAL pr = F(dbinom)(i, coi - 1, relatedness);
        const REAL pam = F(clamp_pam)(pam_vec[coi - i - 1]);
        const REAL i_res = LOG1P(-pam) + log_total * (REAL)(coi - i);
        res[nres++] = pr + i_res;
    }
    return F(log_sum_exp)(res, nres);
}

/* ---- a7: Chain::calc_observation_process (src/chain.cpp:923-961) --------------------------- */
static REAL F(observation)(const int *idx, int k, const int *obs, int A, REAL eps_neg,
                           REAL eps_pos, REAL max_eps_neg, REAL max_eps_pos)
{
    unsigned fp = 0, tp = 0, fn = 0, tn = 0;
    int ptr = 0;
    for (int j = 0; j < A; ++j)
    {
        const int next = ptr < k ? idx[ptr] : -1;
        const int in = (j == next);
        const int e = obs[j] & 1;
        fp += e & !in;
        tp += e & in;
        fn += !e & in;
        tn += !e & !in;
        ptr += in;
    }
    const REAL norm = (REAL)1 / (REAL)A;
    REAL res = (REAL)0;
    res += LOG(1 - (eps_neg * max_eps_neg * norm)) * tp;
    res += LOG(eps_neg * max_eps_neg * norm) * fn;
    res += LOG(1 - (eps_pos * max_eps_pos * norm)) * tn;
    res += LOG(eps_pos * max_eps_pos * norm) * fp;
    return res;
}

/* ---- a8: log proposal density of Sampler::sample_latent_genotype as a pure function of the
 * counts it draws (src/sampler.cpp:230-340).  Returns -inf for the infeasible branch. -------- */
static REAL F(binom_coef)(unsigned n, unsigned k)
{
    /* boost::math::binomial_coefficient<REAL>: C(n,k) as an integer-valued REAL */
    if (k > n) return (REAL)0;
    if (k > n - k) k = n - k;
    long double c = 1.0L;
    for (unsigned i = 1; i <= k; ++i)
    {
        c = c * (long double)(n - k + i) / (long double)i;
        c = floorl(c + 0.5L);
    }
    return (REAL)c;
}

static REAL F(latent_log_prob)(int A, int obs_pos, int fn, int fp, int coi, REAL eps_pos,
                               REAL eps_neg)
{
    const int obs_neg = A - obs_pos;
    const int min_fn = (obs_pos == 0);
    const int tn = obs_neg - fn;
    const REAL lp_fn = LOG(F(binom_coef)(obs_neg, fn - min_fn)) +
                       (fn - min_fn) * LOG(eps_neg / (REAL)A) + tn * LOG(1 - (eps_neg / (REAL)A));
    int min_fp = (obs_pos + fn) - coi;
    if (min_fp < 0) min_fp = 0;
    int max_fp = obs_pos + fn - 1;
    if (obs_pos < max_fp) max_fp = obs_pos;
    if (max_fp < min_fp) return -(REAL)INFINITY;
    const int tp = obs_pos - fp;
    const REAL lp_fp = LOG(F(binom_coef)(obs_pos, fp - min_fp)) +
                       (fp - min_fp) * LOG(eps_pos / (REAL)A) + tp * LOG(1 - (eps_pos / (REAL)A));
    const REAL lp_pos_idx = -LOG(F(binom_coef)(obs_pos, tp));
    const REAL lp_neg_idx = -LOG(F(binom_coef)(obs_neg, fn));
    return lp_pos_idx + lp_neg_idx + lp_fp + lp_fn;
}

/* ---- Sampler::sample_constrained given its normal draw z (src/sampler.cpp:176-190) --------- */
static void F(constrained)(REAL z, REAL curr, REAL lo, REAL hi, REAL *prop_out, REAL *adj_out)
{
    const REAL unconstrained = LOG(curr - lo) - LOG(hi - curr);
    const REAL exp_prop = EXP(z + unconstrained);
    REAL prop = (hi * exp_prop + lo) / (exp_prop + 1);
    prop = prop < lo ? lo : (prop > hi ? hi : prop); /* UtilFunctions::clamp */
    *adj_out = LOG(prop - lo) + LOG(hi - prop) - LOG(curr - lo) - LOG(hi - curr);
    *prop_out = prop;
}

/* ---- SALT simplex proposal math (src/mcmc_utils.h:150-313) --------------------------------- */
static REAL F(logit)(REAL x)
{
    if (x < .5) return LOG(x) - LOG1P(-x);
    return LOG(x / (1 - x));
}

static void F(log_pq)(REAL x, REAL *lp, REAL *lq)
{
    const REAL ex = EXP(x);
    if (x < 0)
    {
        *lq = -LOG1P(ex);
        *lp = *lq + x;
    }
    else
    {
        *lp = -LOG1P(1 / ex);
        *lq = *lp - x;
    }
}

static int F(cmp_desc)(const void *a, const void *b)
{
    const REAL x = *(const REAL *)a, y = *(const REAL *)b;
    return (x < y) - (x > y);
}

static REAL F(logit_sum)(const REAL *x, int n)
{
    REAL s[ORC_MAX_A], lp[ORC_MAX_A], lq[ORC_MAX_A];
    for (int i = 0; i < n; ++i) s[i] = x[i];
    qsort(s, (size_t)n, sizeof(REAL), F(cmp_desc));
    for (int i = 0; i < n; ++i) F(log_pq)(s[i], &lp[i], &lq[i]);
    REAL cumsum = (REAL)0;
    if (s[0] < 0)
    {
        const REAL lp1 = lp[0];
        for (int i = 1; i < n; ++i) cumsum += EXP(lp[i] - lp1);
        return lp1 + LOG1P(cumsum);
    }
    const REAL lq1 = lq[0];
    for (int i = 1; i < n; ++i) cumsum += EXP(lp[i]);
    return LOG1P(-EXP(lq1) + cumsum);
}

static void F(logit_scale)(const REAL *x, int n, REAL scale, REAL *out)
{
    for (int i = 0; i < n; ++i)
    {
        const int ok = ((double)scale < log(2.0)) && (FABS(scale) < FABS(x[i] + scale));
        const REAL u = -scale - (!ok) * x[i];
        const REAL v = -scale - ok * x[i];
        const REAL ev = EXP(v);
        const REAL eumo = EXPM1(u);
        REAL l2;
        if (isinf(eumo))
            l2 = (u > v ? u : v) + LOG1P(EXP(-FABS(u - v)));
        else
            l2 = LOG(eumo + ev);
        if (v > LOG(2 * FABS(eumo)))
            out[i] = -(v + LOG1P(eumo / ev));
        else
            out[i] = -l2;
    }
}

static REAL F(expit)(REAL x) { return 1 / (1 + EXP(-x)); }

/* One SALT proposal, deterministic part (src/chain.cpp:481-501). */
static void F(salt)(const REAL *p, int A, int idx, REAL logit_prop, REAL *prop_p, REAL *log_adj)
{
    REAL lv[ORC_MAX_A], rest[ORC_MAX_A], scaled[ORC_MAX_A];
    for (int a = 0; a < A; ++a) lv[a] = F(logit)(p[a]);
    const REAL logit_curr = lv[idx];
    REAL clp, clq, plp, plq;
    F(log_pq)(logit_curr, &clp, &clq);
    F(log_pq)(logit_prop, &plp, &plq);
    int n = 0;
    for (int a = 0; a < A; ++a)
        if (a != idx) rest[n++] = lv[a];
    const REAL ls = plq - F(logit_sum)(rest, n);
    F(logit_scale)(rest, n, ls, scaled);
    *log_adj = (clp - plp) + (A - 1) * (clq - plq);
    n = 0;
    for (int a = 0; a < A; ++a) prop_p[a] = F(expit)(a == idx ? logit_prop : scaled[n++]);
}

/* ---- priors: R::dbeta / R::dgamma as called from src/sampler.cpp:28-36,52-55,139-166.  R's
 * nmath is not under /root/reference; restated from its published definition (see
 * oracle/shim/Rmath.h).  Computed in double and rounded to REAL, as the reference's
 * `float Sampler::dbeta(...) { return R::dbeta(...); }` does. ------------------------------- */
static REAL F(dbeta_log)(REAL x_, REAL a_, REAL b_)
{
    const double x = x_, a = a_, b = b_;
    if (isnan(x) || isnan(a) || isnan(b)) return (REAL)(x + a + b);
    if (a < 0 || b < 0) return (REAL)NAN;
    if (x < 0 || x > 1) return -(REAL)INFINITY;
    if (x == 0)
    {
        if (a > 1) return -(REAL)INFINITY;
        if (a < 1) return (REAL)INFINITY;
        return (REAL)log(b);
    }
    if (x == 1)
    {
        if (b > 1) return -(REAL)INFINITY;
        if (b < 1) return (REAL)INFINITY;
        return (REAL)log(a);
    }
    const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
    return (REAL)((a - 1) * log(x) + (b - 1) * log1p(-x) - lbeta);
}

static REAL F(dgamma_log)(REAL x_, REAL shape_, REAL scale_)
{
    const double x = x_, shape = shape_, scale = scale_;
    if (isnan(x) || isnan(shape) || isnan(scale)) return (REAL)(x + shape + scale);
    if (shape < 0 || scale <= 0) return (REAL)NAN;
    if (x < 0) return -(REAL)INFINITY;
    if (shape == 0) return (x == 0) ? (REAL)INFINITY : -(REAL)INFINITY;
    if (x == 0)
    {
        if (shape < 1) return (REAL)INFINITY;
        if (shape > 1) return -(REAL)INFINITY;
        return (REAL)(-log(scale));
    }
    return (REAL)((shape - 1) * log(x) - x / scale - lgamma(shape) - shape * log(scale));
}

/* ---- a1: Chain::calculate_genotype_likelihood (src/chain.cpp:1114-1134) on bit masks -------- */
static REAL F(cell_llik)(unsigned __int128 latent, unsigned __int128 obs, int A, int is_missing, const REAL *p,
                         int coi, REAL relatedness, REAL eps_neg, REAL eps_pos,
                         int allow_relatedness, REAL max_eps_neg, REAL max_eps_pos,
                         REAL *t_out, REAL *o_out)
{
    if (is_missing)
    {
        if (t_out) *t_out = 0;
        if (o_out) *o_out = 0;
        return (REAL)0;
    }
    int idx[ORC_MAX_A], ov[ORC_MAX_A];
    int k = 0;
    for (int a = 0; a < A; ++a)
    {
        if ((unsigned)(latent >> a) & 1u) idx[k++] = a;
        ov[a] = (int)((unsigned)(obs >> a) & 1u);
    }
    const REAL t = F(transmission)(idx, k, p, coi, relatedness, allow_relatedness);
    const REAL o = F(observation)(idx, k, ov, A, eps_neg, eps_pos, max_eps_neg, max_eps_pos);
    if (t_out) *t_out = t;
    if (o_out) *o_out = o;
    return t + o;
}
