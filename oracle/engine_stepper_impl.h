/* TEST INFRASTRUCTURE ONLY (oracle/).  Included at the end of moire_oracle.c.
 *
 * CPU restatement of the reference's eight Metropolis updates (src/chain.cpp:165-856) in the
 * form the CUDA engine runs them: same proposal rules, validity checks, priors, NaN => reject
 * semantics, acceptance comparisons and Robbins-Monro adaptation as the reference, evaluated in
 * double with the f64 functions above, but
 *   (1) every (chain, update, iteration, sample | locus, cell) draws from its own counter-based
 *       Philox4x32-10 stream (the reference's std::ranlux24_base stream is seeded from
 *       std::random_device and cannot be reproduced), and
 *   (2) the Metropolis ratio uses the change of the affected cells only (the reference re-sums
 *       all N*L cells per proposal, src/chain.cpp:978-1026; same number up to its rounding).
 * Samples (resp. loci) are visited in index order; by the conditional independence argued in
 * DESIGN.md the order does not matter.  One call advances ONE chain by ONE update kind, so a
 * test can interleave engine and stepper and compare states after every kind.
 */

/* ---- Philox4x32-10, written independently of moire_b200/csrc/rng.cuh; pinned against the
 * Random123 known-answer vectors in tests/test_oracle_golden.py ------------------------------ */
static void stp_philox(const uint32_t c_in[4], uint32_t k0, uint32_t k1, uint32_t out[4])
{
    uint32_t c[4] = {c_in[0], c_in[1], c_in[2], c_in[3]};
    for (int round = 0; round < 10; ++round)
    {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0;
        c[1] = n1;
        c[2] = n2;
        c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    for (int i = 0; i < 4; ++i) out[i] = c[i];
}

void orc_philox4x32(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4])
{
    stp_philox(counter, key[0], key[1], out);
}

typedef struct
{
    uint32_t ctr[4], k0, k1, buf[4];
    int have;
} stp_stream;

static void stp_open(stp_stream *s, uint64_t seed, uint32_t iteration, uint32_t kind,
                     uint32_t chain, uint32_t unit, uint32_t sub)
{
    s->k0 = (uint32_t)seed;
    s->k1 = (uint32_t)(seed >> 32);
    s->ctr[0] = iteration;
    s->ctr[1] = (kind << 16) | (chain & 0xFFFFu);
    s->ctr[2] = unit;
    s->ctr[3] = sub << 16;
    s->have = 0;
}

static uint32_t stp_u32(stp_stream *s)
{
    if (s->have == 0)
    {
        stp_philox(s->ctr, s->k0, s->k1, s->buf);
        s->ctr[3] += 1;
        s->have = 4;
    }
    const uint32_t v = s->buf[4 - s->have];
    s->have--;
    return v;
}

static double stp_uniform(stp_stream *s) { return ((double)stp_u32(s) + 0.5) * 0x1p-32; }

static double stp_normal(stp_stream *s)
{
    const double u1 = stp_uniform(s);
    const double u2 = stp_uniform(s);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586477 * u2);
}

/* Sampler::sample_coi_delta(mean) (src/sampler.cpp:152-156) by inversion */
static int stp_coi_delta(stp_stream *s, double log_1mp)
{
    const int sign = (stp_u32(s) & 1u) ? 1 : -1;
    const double u = stp_uniform(s);
    return sign * (int)floor(log(u) / log_1mp);
}

/* Allele sets: the stepper computes on 128-bit sets whatever the engine's build; the arrays it shares
 * with the test are W 64-bit words per set (W = 1: libmoire_b200.so, W = 2: libmoire_b200_w128.so,
 * word 0 = alleles 0..63). */
typedef unsigned __int128 stp_mask;
static int stp_popc(stp_mask x)
{
    return __builtin_popcountll((unsigned long long)x) + __builtin_popcountll((unsigned long long)(x >> 64));
}
static stp_mask stp_load(const uint64_t *a, size_t cell, int W)
{
    stp_mask m = a[cell * (size_t)W];
    if (W > 1) m |= (stp_mask)a[cell * (size_t)W + 1] << 64;
    return m;
}
static void stp_store(uint64_t *a, size_t cell, int W, stp_mask m)
{
    a[cell * (size_t)W] = (uint64_t)m;
    if (W > 1) a[cell * (size_t)W + 1] = (uint64_t)(m >> 64);
}

static stp_mask stp_choose_bits(stp_stream *s, stp_mask set, int n, int need)
{
    if (need <= 0) return 0;
    if (need >= n) return set;
    if (need == 1 || need == n - 1)
    {
        /* keep one / drop one: one uniform picks the allele (the engine's choose_bits) */
        const int idx = (int)(stp_uniform(s) * (double)n);
        stp_mask r = set;
        for (int i = 0; i < idx; ++i) r &= r - 1;
        const stp_mask bit = r & ((stp_mask)0 - r);
        return need == 1 ? bit : set ^ bit;
    }
    stp_mask out = 0;
    int remaining = n;
    while (need > 0)
    {
        if (need == remaining)
        {
            out |= set;
            break;
        }
        const stp_mask low = set & ((stp_mask)0 - set);
        set ^= low;
        const double u = stp_uniform(s);
        if (u * (double)remaining < (double)need)
        {
            out |= low;
            --need;
        }
        --remaining;
    }
    return out;
}

/* the engine's pow_uint / binom_inverse (likelihood.cuh), operation for operation */
static double stp_pow_uint(double x, int n)
{
    double r = (n & 1) ? x : 1.0;
    for (int b = 1; b < 7; ++b)
    {
        x = x * x;
        const double rx = r * x;
        r = ((n >> b) & 1) ? rx : r;
    }
    return r;
}

static int stp_binom_inverse(double u, int n, double q, double odds)
{
    double pmf = stp_pow_uint(q, n), cdf = pmf;
    int k = 0;
    while (u >= cdf && k < n)
    {
        ++k;
        const double ratio = (double)(n - k + 1) / (double)k;
        const double po = pmf * odds;
        pmf = po * ratio;
        cdf = cdf + pmf;
    }
    return k;
}

/* Sampler::sample_latent_genotype (src/sampler.cpp:230-340) on masks */
static stp_mask stp_sample_latent(stp_stream *s, stp_mask obs, int A, int coi, double eps_pos,
                                  double eps_neg, double *log_prob)
{
    const stp_mask all = A >= 128 ? ~(stp_mask)0 : (((stp_mask)1 << A) - 1);
    const int obs_pos = stp_popc(obs);
    const int obs_neg = A - obs_pos;
    const int min_fn = obs_pos == 0;
    int max_fn = coi < obs_neg ? coi : obs_neg;
    if (max_fn < min_fn) max_fn = min_fn;
    /* counts by inversion, one uniform each, always consumed in this order (the engine's stream use) */
    const double u_fn = stp_uniform(s), u_fp = stp_uniform(s);
    const double pr_en = eps_neg / (double)A, pr_ep = eps_pos / (double)A;
    const double q_en = 1.0 - pr_en, q_ep = 1.0 - pr_ep;
    const int fn = min_fn + stp_binom_inverse(u_fn, max_fn - min_fn, q_en, pr_en / q_en);
    int min_fp = obs_pos + fn - coi;
    if (min_fp < 0) min_fp = 0;
    int max_fp = obs_pos + fn - 1;
    if (obs_pos < max_fp) max_fp = obs_pos;
    if (max_fp < min_fp)
    {
        *log_prob = -INFINITY;
        return 0;
    }
    const int fp = min_fp + stp_binom_inverse(u_fp, max_fp - min_fp, q_ep, pr_ep / q_ep);
    const stp_mask pos = stp_choose_bits(s, obs, obs_pos, obs_pos - fp);
    const stp_mask neg = stp_choose_bits(s, ~obs & all, obs_neg, fn);
    *log_prob = latent_log_prob_f64(A, obs_pos, fn, fp, coi, eps_pos, eps_neg);
    return pos | neg;
}

/* n draws of the engine's +-Geometric COI step (mean 2, as every update uses it). */
void orc_stp_coi_delta_many(uint64_t seed, int n, int *out)
{
    const double log_geom = -0.40546510810816438; /* log(1 - 1/3) */
    for (int i = 0; i < n; ++i)
    {
        stp_stream s;
        stp_open(&s, seed, (uint32_t)i, 3u, 0u, (uint32_t)i >> 8, 0u);
        out[i] = stp_coi_delta(&s, log_geom);
    }
}

/* n latent-genotype draws from the engine's streams (seed, draw number), for distribution tests
 * against the reference's own sampler: masks[i], log_prob[i]. */
void orc_stp_sample_latent_many(uint64_t seed, int n, uint64_t obs, int A, int coi, double eps_pos,
                                double eps_neg, uint64_t *masks, double *log_prob)
{
    for (int i = 0; i < n; ++i)
    {
        stp_stream s;
        stp_open(&s, seed, (uint32_t)i, 6u, 0u, (uint32_t)i >> 8, 1u);
        masks[i] = (uint64_t)stp_sample_latent(&s, obs, A, coi, eps_pos, eps_neg, &log_prob[i]);
    }
}

typedef struct
{
    /* shapes, data, model */
    int N, L, a_pad;
    const int *num_alleles;
    const uint64_t *obs;       /* [N][L][W] */
    const uint8_t *is_missing; /* [N][L] */
    int allow_relatedness, burnin, max_coi;
    double mean_coi_shape, mean_coi_scale;
    double max_eps_pos, eps_pos_alpha, eps_pos_beta;
    double max_eps_neg, eps_neg_alpha, eps_neg_beta;
    double r_alpha, r_beta;
    uint64_t seed;
    int chain; /* global chain id */
    double temp;
    /* state (in/out) */
    int *m;
    double *r, *eps_pos, *eps_neg; /* [N] */
    double *p;                     /* [L][a_pad] */
    uint64_t *latent;              /* [N][L][W] */
    double *lg_adj, *cell_t, *cell_o;
    double *prior_eps_neg, *prior_eps_pos, *prior_coi, *prior_r; /* [N] */
    double *mean_coi, *hyper;                                    /* [1] */
    /* tuning (in/out) */
    double *sd_eps_neg, *sd_eps_pos, *sd_r, *sd_mr; /* [N] */
    int *acc_eps_neg, *acc_eps_pos, *acc_m, *acc_r, *acc_mr, *acc_samples;
    double *p_sd;        /* [L][a_pad] */
    int *p_acc, *p_att;  /* [L][a_pad] */
    double *mean_coi_sd, *mean_coi_acc; /* [1] */
    unsigned long long evals;           /* out: non-missing cells evaluated */
    int W;                              /* 64-bit words per allele set (0 reads as 1) */
} stp_chain;
static int stp_words(const stp_chain *c) { return c->W > 1 ? c->W : 1; }

static double stp_adapt(double sd, double accepts, double denom, double root_arg, double floor_)
{
    const double rate = accepts / denom;
    const double upd = (rate - .23) / sqrt(root_arg);
    const double nv = sd + upd;
    return nv > floor_ ? nv : floor_;
}

static void stp_cell(const stp_chain *c, int i, int j, stp_mask latent, int m, double r,
                     double eps_neg, double eps_pos, double *t, double *o)
{
    const size_t cell = (size_t)i * c->L + j;
    cell_llik_f64(latent, stp_load(c->obs, cell, stp_words(c)), c->num_alleles[j], 0, c->p + (size_t)j * c->a_pad, m, r,
                  eps_neg, eps_pos, c->allow_relatedness, c->max_eps_neg, c->max_eps_pos, t, o);
}

enum
{
    STP_EPS_NEG = 0,
    STP_EPS_POS = 1,
    STP_P = 2,
    STP_M = 3,
    STP_R = 4,
    STP_EFF_COI = 5,
    STP_SAMPLES = 6,
    STP_MEAN_COI = 7
};

/* The six per-sample updates: update_eps_neg :632-695, update_eps_pos :567-630, update_m
 * :165-224, update_r :315-374, update_eff_coi :226-313, update_samples :697-807. */
static void stp_update_sample_kind(stp_chain *c, int kind, int iteration)
{
    const int L = c->L;
    const int resample = (kind == STP_M || kind == STP_EFF_COI || kind == STP_SAMPLES);
    const int need_t = resample || kind == STP_R;
    const int need_o = resample || kind == STP_EPS_NEG || kind == STP_EPS_POS;
    const double min_sampled = DBL_MIN;
    const double log_geom = -0.40546510810816438; /* log(1 - 1/(1+2)) */
    stp_mask *nlat = (stp_mask *)malloc(sizeof(stp_mask) * L);
    const int W = stp_words(c);
    double *nt = (double *)malloc(sizeof(double) * L);
    double *no = (double *)malloc(sizeof(double) * L);
    double *nadj = (double *)malloc(sizeof(double) * L);
    for (int i = 0; i < c->N; ++i)
    {
        stp_stream rng;
        stp_open(&rng, c->seed, (uint32_t)iteration, (uint32_t)kind, (uint32_t)c->chain,
                 (uint32_t)i, 0);
        int valid = 1, adapt = 0;
        int m_new = c->m[i];
        double r_new = c->r[i], en_new = c->eps_neg[i], ep_new = c->eps_pos[i];
        double adj = 0.0, dprior = 0.0;
        double pr_en = 0, pr_ep = 0, pr_coi = 0, pr_r = 0;
        if (kind == STP_EPS_NEG || kind == STP_SAMPLES)
        {
            const double z = c->sd_eps_neg[i] * stp_normal(&rng);
            double prop, a;
            constrained_f64(z, en_new, min_sampled, 1.0, &prop, &a);
            valid &= (prop < 1.0 && prop > 1e-32);
            en_new = prop;
            adj += a;
            pr_en = dbeta_log_f64(prop, c->eps_neg_alpha, c->eps_neg_beta);
            dprior += pr_en - c->prior_eps_neg[i];
        }
        if (kind == STP_EPS_POS || kind == STP_SAMPLES)
        {
            const double z = c->sd_eps_pos[i] * stp_normal(&rng);
            double prop, a;
            constrained_f64(z, ep_new, min_sampled, 1.0, &prop, &a);
            valid &= (prop < 1.0 && prop > 1e-32);
            ep_new = prop;
            adj += a;
            pr_ep = dbeta_log_f64(prop, c->eps_pos_alpha, c->eps_pos_beta);
            dprior += pr_ep - c->prior_eps_pos[i];
        }
        if (kind == STP_R || (kind == STP_SAMPLES && c->allow_relatedness))
        {
            const double z = c->sd_r[i] * stp_normal(&rng);
            double prop, a;
            constrained_f64(z, r_new, min_sampled, .99, &prop, &a);
            if (kind == STP_SAMPLES) valid &= (prop < 1.0 && prop > 1e-32);
            r_new = prop;
            adj += a;
        }
        if (kind == STP_SAMPLES && !c->allow_relatedness) r_new = 0.0;
        if (kind == STP_EFF_COI)
        {
            const double curr_eff = (double)(m_new - 1) * (1.0 - r_new) + 1.0;
            const double z = c->sd_mr[i] * stp_normal(&rng);
            double prop_eff, a;
            constrained_f64(z, curr_eff, 1.0, (double)c->max_coi, &prop_eff, &a);
            const int prop_m = m_new + stp_coi_delta(&rng, log_geom);
            const double prop_r = 1.0 - (prop_eff - 1.0) / ((double)prop_m - 1.0);
            if (prop_m <= 0 || prop_m > c->max_coi || prop_r > 1.0 || prop_r < 1e-32) valid = 0;
            m_new = prop_m;
            r_new = prop_r;
            adj += a;
        }
        if (kind == STP_M || kind == STP_SAMPLES)
        {
            const int prop_m = m_new + stp_coi_delta(&rng, log_geom);
            valid &= (prop_m > 0 && prop_m <= c->max_coi);
            m_new = prop_m;
        }
        if (kind == STP_R || kind == STP_EFF_COI || kind == STP_SAMPLES)
        {
            pr_r = dbeta_log_f64(r_new, c->r_alpha, c->r_beta);
            dprior += pr_r - c->prior_r[i];
        }
        if (kind == STP_M || kind == STP_EFF_COI || kind == STP_SAMPLES)
        {
            pr_coi = dztpois_f64(valid ? m_new : 1, *c->mean_coi);
            dprior += pr_coi - c->prior_coi[i];
        }
        if (kind == STP_EPS_NEG || kind == STP_EPS_POS || kind == STP_EFF_COI) adapt = valid;
        if (kind == STP_R) adapt = 1;
        const double logu = log(stp_uniform(&rng));

        int accept = 0;
        if (valid)
        {
            double d = 0.0, a = 0.0;
            for (int j = 0; j < L; ++j)
            {
                const size_t cell = (size_t)i * L + j;
                stp_mask lat = stp_load(c->latent, cell, W);
                if (resample)
                {
                    stp_stream crng;
                    stp_open(&crng, c->seed, (uint32_t)iteration, (uint32_t)kind,
                             (uint32_t)c->chain, (uint32_t)i, (uint32_t)(j + 1));
                    double lp;
                    lat = stp_sample_latent(&crng, stp_load(c->obs, cell, W), c->num_alleles[j], m_new, ep_new,
                                            en_new, &lp);
                    nlat[j] = lat;
                    nadj[j] = lp;
                    a += lp - c->lg_adj[cell];
                }
                nt[j] = c->cell_t[cell];
                no[j] = c->cell_o[cell];
                if (!c->is_missing[cell])
                {
                    double t, o;
                    stp_cell(c, i, j, lat, m_new, r_new, en_new, ep_new, &t, &o);
                    if (need_t) nt[j] = t;
                    if (need_o) no[j] = o;
                    d += (nt[j] + no[j]) - (c->cell_t[cell] + c->cell_o[cell]);
                    c->evals++;
                }
            }
            const double dpost = c->temp * d + dprior;
            const double ratio = dpost + (adj + a);
            if (kind == STP_M || kind == STP_SAMPLES)
                accept = !isnan(dpost) && logu <= ratio;
            else if (kind == STP_EFF_COI)
                accept = !(isnan(dpost) || isnan(ratio) || logu > ratio);
            else
                accept = !(isnan(dpost) || logu > ratio);
        }
        if (accept)
        {
            if (kind == STP_EPS_NEG || kind == STP_SAMPLES)
            {
                c->eps_neg[i] = en_new;
                c->prior_eps_neg[i] = pr_en;
            }
            if (kind == STP_EPS_POS || kind == STP_SAMPLES)
            {
                c->eps_pos[i] = ep_new;
                c->prior_eps_pos[i] = pr_ep;
            }
            if (kind == STP_R || kind == STP_EFF_COI || kind == STP_SAMPLES)
            {
                c->r[i] = r_new;
                c->prior_r[i] = pr_r;
            }
            if (kind == STP_M || kind == STP_EFF_COI || kind == STP_SAMPLES)
            {
                c->m[i] = m_new;
                c->prior_coi[i] = pr_coi;
            }
            for (int j = 0; j < L; ++j)
            {
                const size_t cell = (size_t)i * L + j;
                if (resample)
                {
                    stp_store(c->latent, cell, W, nlat[j]);
                    c->lg_adj[cell] = nadj[j];
                }
                if (!c->is_missing[cell])
                {
                    c->cell_t[cell] = nt[j];
                    c->cell_o[cell] = no[j];
                }
            }
        }
        int *accp = kind == STP_EPS_NEG   ? c->acc_eps_neg
                    : kind == STP_EPS_POS ? c->acc_eps_pos
                    : kind == STP_M       ? c->acc_m
                    : kind == STP_R       ? c->acc_r
                    : kind == STP_EFF_COI ? c->acc_mr
                                          : c->acc_samples;
        if (accept) accp[i] += 1;
        if (adapt && iteration < c->burnin && iteration > 15)
        {
            double *sdp = kind == STP_EPS_NEG   ? c->sd_eps_neg
                          : kind == STP_EPS_POS ? c->sd_eps_pos
                          : kind == STP_R       ? c->sd_r
                                                : c->sd_mr;
            sdp[i] = stp_adapt(sdp[i], (double)accp[i], (double)iteration, (double)iteration + 1.0,
                               .01);
        }
    }
    free(nlat);
    free(nt);
    free(no);
    free(nadj);
}

/* update_p, the SALT sampler (src/chain.cpp:466-565) */
static void stp_update_p(stp_chain *c, int iteration)
{
    const int N = c->N, L = c->L;
    double *tn = (double *)malloc(sizeof(double) * N);
    for (int j = 0; j < L; ++j)
    {
        const int A = c->num_alleles[j];
        double *pj = c->p + (size_t)j * c->a_pad;
        double *sd = c->p_sd + (size_t)j * c->a_pad;
        int *acc = c->p_acc + (size_t)j * c->a_pad;
        int *att = c->p_att + (size_t)j * c->a_pad;
        for (int rep = 0; rep < A; ++rep)
        {
            stp_stream rng;
            stp_open(&rng, c->seed, (uint32_t)iteration, STP_P, (uint32_t)c->chain, (uint32_t)j,
                     (uint32_t)rep);
            const int idx = (int)(((uint64_t)stp_u32(&rng) * (uint64_t)A) >> 32);
            const double z = stp_normal(&rng);
            const double logu = log(stp_uniform(&rng));
            att[idx] += 1;
            const double logit_prop = logit_f64(pj[idx]) + sd[idx] * z;
            double prop[ORC_MAX_A], log_adj;
            salt_f64(pj, A, idx, logit_prop, prop, &log_adj);
            int below = 0;
            for (int a = 0; a < A; ++a) below |= prop[a] < 1e-12;
            if (below) break;
            double keep[ORC_MAX_A];
            for (int a = 0; a < A; ++a)
            {
                keep[a] = pj[a];
                pj[a] = prop[a];
            }
            double d = 0.0;
            for (int i = 0; i < N; ++i)
            {
                const size_t cell = (size_t)i * L + j;
                tn[i] = c->cell_t[cell];
                if (!c->is_missing[cell])
                {
                    double t, o;
                    stp_cell(c, i, j, stp_load(c->latent, cell, stp_words(c)), c->m[i], c->r[i], c->eps_neg[i],
                             c->eps_pos[i], &t, &o);
                    tn[i] = t;
                    d += t - c->cell_t[cell];
                    c->evals++;
                }
            }
            const double dpost = c->temp * d;
            const double ratio = dpost + log_adj;
            const int accept = !(isnan(dpost) || logu > ratio);
            if (accept)
            {
                for (int i = 0; i < N; ++i)
                    if (!c->is_missing[(size_t)i * L + j]) c->cell_t[(size_t)i * L + j] = tn[i];
                acc[idx] += 1;
            }
            else
            {
                for (int a = 0; a < A; ++a) pj[a] = keep[a];
            }
            if (iteration < c->burnin && att[idx] > 15)
            {
                const double rate = ((double)acc[idx] + 1.0) / ((double)att[idx] + 1.0);
                const double nv = sd[idx] + (rate - .23) / sqrt((double)att[idx] + 1.0);
                sd[idx] = nv > .01 ? nv : .01;
            }
        }
    }
    free(tn);
}

/* update_mean_coi (src/chain.cpp:809-856) */
static void stp_update_mean_coi(stp_chain *c, int iteration)
{
    stp_stream rng;
    stp_open(&rng, c->seed, (uint32_t)iteration, STP_MEAN_COI, (uint32_t)c->chain, 0, 0);
    const double z = *c->mean_coi_sd * stp_normal(&rng);
    double prop, adj;
    constrained_f64(z, *c->mean_coi, 1e-5, (double)c->max_coi, &prop, &adj);
    const double logu = log(stp_uniform(&rng));
    double d = 0.0;
    for (int i = 0; i < c->N; ++i) d += dztpois_f64(c->m[i], prop) - c->prior_coi[i];
    const double hyper_new = dgamma_log_f64(prop, c->mean_coi_shape, c->mean_coi_scale);
    const double dpost = d + (hyper_new - *c->hyper);
    const double ratio = dpost + adj;
    const int accept = !isnan(dpost) && logu <= ratio;
    if (accept)
    {
        *c->mean_coi = prop;
        *c->hyper = hyper_new;
        *c->mean_coi_acc += 1.0;
        for (int i = 0; i < c->N; ++i) c->prior_coi[i] = dztpois_f64(c->m[i], prop);
    }
    if (iteration < c->burnin && iteration > 15)
        *c->mean_coi_sd = stp_adapt(*c->mean_coi_sd, *c->mean_coi_acc, (double)(iteration + 1),
                                    (double)iteration + 1.0, .1);
}

void orc_step_update(void *chain, int kind, int iteration)
{
    stp_chain *c = (stp_chain *)chain;
    if (kind == STP_P)
        stp_update_p(c, iteration);
    else if (kind == STP_MEAN_COI)
        stp_update_mean_coi(c, iteration);
    else
        stp_update_sample_kind(c, kind, iteration);
}

/* Priors of a state (Chain::initialize_likelihood, src/chain.cpp:1194-1219) */
void orc_eval_priors(void *chain)
{
    stp_chain *c = (stp_chain *)chain;
    for (int i = 0; i < c->N; ++i)
    {
        c->prior_eps_neg[i] = dbeta_log_f64(c->eps_neg[i], c->eps_neg_alpha, c->eps_neg_beta);
        c->prior_eps_pos[i] = dbeta_log_f64(c->eps_pos[i], c->eps_pos_alpha, c->eps_pos_beta);
        c->prior_r[i] = dbeta_log_f64(c->r[i], c->r_alpha, c->r_beta);
        c->prior_coi[i] = dztpois_f64(c->m[i], *c->mean_coi);
    }
    *c->hyper = dgamma_log_f64(*c->mean_coi, c->mean_coi_shape, c->mean_coi_scale);
}
