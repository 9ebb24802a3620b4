// Device math of the hot path: per-cell transmission / observation log-likelihood,
// prob_any_missing, the latent-genotype proposal, priors and the constrained proposal.
//
// Arithmetic is fp64 and follows the double variant of the reference (every `float` read as
// `double`), operation for operation wherever the order decides the rounding: the
// inclusion-exclusion sums, the EGF recurrence and the sequential sums use explicit
// __d{add,sub,mul,div}_rn so that nvcc cannot contract them into FMAs.  What is left to differ
// from the CPU is libm (log / log1p / exp), i.e. a couple of ulps.
//
// Reference (paths relative to /root/reference):
//   calc_transmission_process   src/chain.cpp:858-921
//   calc_observation_process    src/chain.cpp:923-961
//   probAnyMissingFunctor       src/prob_any_missing.cpp:13-234
//   CombinationIndicesGenerator src/combination_indices_generator.cpp:5-76  (as a bit trick)
//   sample_latent_genotype      src/sampler.cpp:230-340
//   sample_constrained          src/sampler.cpp:176-190
//   dbinom / dztpois            src/sampler.cpp:38-50
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#include "mask.cuh"
#include "rng.cuh"

namespace moire
{
// The fused kernels were instruction-fetch bound (ncu: GPC instruction cache at ~85 % of its
// request rate, SM instruction-cache hit rate 66 %, profiles/README.md), and CUDA's
// double-precision log / log1p / exp / expm1 are 100-250 instructions each when inlined.  Only
// the per-cell hot path (the one-thread evaluator's log of the frequency sum and the linear
// mixture) inlines them; every other use -- proposals, priors, SALT, the cooperative and EGF
// paths, the mixture's fallback -- shares one out-of-line copy per function.
//
// log1p(-pam) of a clamped PAM is evaluated as log(1 - pam) on every path: pam is clamped to
// [0, 1 - eps]; 1 - pam is exact for pam >= 1/2 and carries a relative rounding error <= 2^-53
// below, i.e. the term differs from log1p(-pam) by at most ~2e-16 in absolute value -- ulp-level,
// like the difference between CUDA's and glibc's log1p -- while the 250-instruction log1p body
// cost 13 % of the step when it sat in the hot loop (profiles/variants_o.log).
// (Tried, round 2: a leaner logarithm for the one log of a cell -- atanh series on the hardware's
// reciprocal seed, ~35 instructions where CUDA's log() showed as ~20 % of the evaluator's executed
// instructions.  It measured 0.4 % of a step: the evaluator is bound by stalls, not by issue slots.
// Not kept; profiles/README.md.)
__device__ __forceinline__ double hot_log(double x) { return log(x); }
__device__ __noinline__ double cold_log(double x) { return log(x); }
__device__ __noinline__ double cold_log1p(double x) { return log1p(x); }
__device__ __noinline__ double cold_exp(double x) { return exp(x); }
__device__ __noinline__ double cold_expm1(double x) { return expm1(x); }
__device__ __forceinline__ double cold_log1m(double x) { return cold_log(1.0 - x); }

#define MOIRE_COOP_INLINE __noinline__
constexpr int kIeMax = 10;      // src/prob_any_missing.cpp:11
constexpr int kMaxCoi = 127;    // MOIRE_MAX_COI: the reference's lgamma LUT bound (src/sampler.cpp:22-25)
constexpr int kLogCDim = kMaskBits + 1;  // logC[n][k], 0 <= k <= n <= 64 or 128 (alleles of a locus)

// x^n, n >= 0, by square and multiply.
__device__ __forceinline__ double pow_uint(double x, int n)
{
    // n < 128 (a COI): at most seven rounds, as many as n has bits (the lanes of a class-uniform
    // chunk share n up to the COI's spread, so the exit is all but uniform; a COI below 8 takes two
    // rounds where the branch-free version always ran six)
    double r = (n & 1) ? x : 1.0;
#pragma unroll 1
    for (n >>= 1; n; n >>= 1)
    {
        x = __dmul_rn(x, x);
        if (n & 1) r = __dmul_rn(r, x);
    }
    return r;
}

// Binomial(n, pr) by inversion of ONE uniform: pmf(0) = (1 - pr)^n, pmf(k) = pmf(k-1) odds (n-k+1)/k.
// The latent-genotype proposal counts its false negatives / false positives this way (one Philox
// word each instead of one per candidate allele); every operation is an IEEE multiply / divide /
// add in a fixed order, so the CPU stepper reproduces the draw bit for bit.  q = 1 - pr,
// odds = pr / (1 - pr).
__device__ __forceinline__ int binom_inverse(double u, int n, double q, double odds)
{
    double pmf = pow_uint(q, n), cdf = pmf;
    int k = 0;
    while (u >= cdf && k < n)
    {
        ++k;
        pmf = __dmul_rn(__dmul_rn(pmf, odds), __ddiv_rn((double)(n - k + 1), (double)k));
        cdf = __dadd_rn(cdf, pmf);
    }
    return k;
}

// Per (sample, distinct allele count) logs, built once per tile.
struct EpsLogs
{
    // calc_observation_process: log(1 - en*Mn/A), log(en*Mn/A), log(1 - ep*Mp/A), log(ep*Mp/A)
    double o_tp, o_fn, o_tn, o_fp;
    // sample_latent_genotype: log(en/A), log(1 - en/A), log(ep/A), log(1 - ep/A), en/A, ep/A
    double l_en, l_1men, l_ep, l_1mep, pr_en, pr_ep;
    // binom_inverse: 1 - en/A, (en/A) / (1 - en/A), and the same for ep
    double q_en, odds_en, q_ep, odds_ep;
};

__device__ __noinline__ void build_eps_logs(EpsLogs &e, double eps_neg, double eps_pos,
                                               double max_eps_neg, double max_eps_pos, int A)
{
    const double norm = __ddiv_rn(1.0, (double)A);
    const double xn = __dmul_rn(__dmul_rn(eps_neg, max_eps_neg), norm);
    const double xp = __dmul_rn(__dmul_rn(eps_pos, max_eps_pos), norm);
    e.o_tp = cold_log(__dsub_rn(1.0, xn));
    e.o_fn = cold_log(xn);
    e.o_tn = cold_log(__dsub_rn(1.0, xp));
    e.o_fp = cold_log(xp);
    e.pr_en = __ddiv_rn(eps_neg, (double)A);
    e.pr_ep = __ddiv_rn(eps_pos, (double)A);
    e.l_en = cold_log(e.pr_en);
    e.l_1men = cold_log(__dsub_rn(1.0, e.pr_en));
    e.l_ep = cold_log(e.pr_ep);
    e.l_1mep = cold_log(__dsub_rn(1.0, e.pr_ep));
    e.q_en = __dsub_rn(1.0, e.pr_en);
    e.odds_en = __ddiv_rn(e.pr_en, e.q_en);
    e.q_ep = __dsub_rn(1.0, e.pr_ep);
    e.odds_ep = __ddiv_rn(e.pr_ep, e.q_ep);
}

// ---- a7: observation process on masks --------------------------------------------------------
__device__ __forceinline__ double observation_ll(mask_t latent, mask_t obs, int A,
                                                 const EpsLogs &e)
{
    const int tp = mpopc(latent & obs);
    const int fn = mpopc(latent & ~obs);
    const int fp = mpopc(obs & ~latent);
    const int tn = A - tp - fn - fp;
    double res = 0.0;
    res = __dadd_rn(res, __dmul_rn(e.o_tp, (double)tp));
    res = __dadd_rn(res, __dmul_rn(e.o_fn, (double)fn));
    res = __dadd_rn(res, __dmul_rn(e.o_tn, (double)tn));
    res = __dadd_rn(res, __dmul_rn(e.o_fp, (double)fp));
    return res;
}

// Numerics switch (moire_params::pipeline, MOIRE_NUMERICS_REF32): the shipped reference computes
// this path in float, and what that does to a posterior is, to leading order, ONE thing -- its clamp
// max_pam = 1 - FLT_EPSILON (src/chain.cpp:888-893) floors log(1 - PAM) at log(2^-23) = -15.9, so
// every cell whose coverage probability is below ~1e-7 gets the same likelihood.  With
// the switch on, PAM is rounded to float and clamped at 1 - FLT_EPSILON exactly there; everything
// else stays fp64.  It exists to attribute posterior differences between the fp32 package and an
// fp64 engine (scripts/posterior_study.py); per process, set at engine creation.
__constant__ double c_max_pam = 1.0 - DBL_EPSILON;
__constant__ int c_pam_float = 0;

__device__ __forceinline__ double clamp_pam(double x)
{
    // src/chain.cpp:890-893 in double: min(1 - eps, max(0, x)).  fmax / fmin are one instruction each
    // and return the other operand for a NaN, which is std::max(0., x)'s NaN -> 0 (the compare-and-select
    // spelling was 11 % of the evaluator's instructions); a -0 may come out as -0: only 1 - pam is used
    double lo = fmax(x, 0.0);
    if (c_pam_float) lo = (double)(float)lo;
    return fmin(lo, c_max_pam);
}

__device__ __forceinline__ double pam_from_coverage(double cov)
{
    const double pam = __dsub_rn(1.0, cov);
    if (pam <= 0.0) return 0.0;
    if (pam >= 1.0) return 1.0;
    return pam;
}

// ---- a6: subset enumeration as a table ----------------------------------------------------------
// The reference walks the non-empty subsets of the k latent alleles by size, then in
// lexicographic order of their index tuples (CombinationIndicesGenerator,
// src/combination_indices_generator.cpp:5-76; src/prob_any_missing.cpp:20-52, 186-214).  The
// order depends on k only, so it is tabulated once on the host (engine.cu: build_subset_tables)
// for k <= 10: entry idx of block k (offset 2^k - k - 1) describes the idx-th subset as
//   bits 0..9   "tail": the elements that differ from the previous subset of the same size
//   bits 10..13 "depth": how many leading elements it shares with that previous subset
//   bit  14     set when the subset has even size (its term enters with a minus sign)
// The running differences 1 - q[c0] - q[c1] - ... of the shared leading elements are kept on a
// small per-thread stack, so a subset costs ~1 subtraction instead of |S|; the value of every
// base is still produced by exactly the reference's chain of subtractions.  A second table holds
// the full element masks for the warp-cooperative evaluator.
constexpr int kSubsetEntries = 2048;  // 2^11 - 11 - 1 = 2036 used
__constant__ unsigned short c_sub_tail[kSubsetEntries];
__device__ unsigned short g_sub_mask[kSubsetEntries];

__host__ __device__ inline int subset_block(int k) { return (1 << k) - k - 1; }

// Per-thread scratch column: slot s of the thread lives at scr[s * ST].  Slots 0..9 hold q (the
// normalised frequencies of the latent alleles, ascending allele order), slots 10..19 the stack
// of partial differences; slots 0..15 are reused for the mixture terms once the sums are done.
constexpr int kScratchSlots = 20;
constexpr int kPartBase = 10;

// ---- a4: vectorised inclusion-exclusion, k <= 10 (src/prob_any_missing.cpp:178-216) --------
// acc[j] accumulates the PAM of n = k + t0 + j draws, j < NACC.  Every accumulator sees the
// subsets in the reference's order and every power is built by the same chain of
// multiplications (sign, then base k-1 times, then once per n), so the bits do not depend on
// NACC or t0.
template <int NACC>
__device__ __forceinline__ void ie_vectorized(double *scr, int ST, int k, int t0,
                                              double (&acc)[NACC])
{
#pragma unroll
    for (int t = 0; t < NACC; ++t) acc[t] = 0.0;
    const unsigned short *tab = c_sub_tail + subset_block(k);
    const int nsub = (1 << k) - 1;
    // sign * base is exact, so the chain starts at +-base and one leading multiplication is saved
    // (k >= 2 here: at least one is due)
    const int lead = k - 2 + t0;
#pragma unroll 1
    for (int idx = 0; idx < nsub; ++idx)
    {
        const unsigned int e = tab[idx];
        unsigned int tail = e & 1023u;
        int pos = (e >> 10) & 15;
        double base = pos ? scr[(kPartBase + pos - 1) * ST] : 1.0;
#pragma unroll 1
        for (;;)
        {
            const int el = __ffs(tail) - 1;
            tail &= tail - 1u;
            base = __dsub_rn(base, scr[el * ST]);
            if (!tail) break;
            scr[(kPartBase + pos) * ST] = base;
            ++pos;
        }
        double r = (e & 0x4000u) ? -base : base;
        // r *= base, `lead` times: a jump into a straight run of multiplications (lead <= 9
        // unless this is a later pass of the wide path)
        switch (lead)
        {
            default:
#pragma unroll 1
                for (int j = 9; j < lead; ++j) r = __dmul_rn(r, base);
            case 9: r = __dmul_rn(r, base);
            case 8: r = __dmul_rn(r, base);
            case 7: r = __dmul_rn(r, base);
            case 6: r = __dmul_rn(r, base);
            case 5: r = __dmul_rn(r, base);
            case 4: r = __dmul_rn(r, base);
            case 3: r = __dmul_rn(r, base);
            case 2: r = __dmul_rn(r, base);
            case 1: r = __dmul_rn(r, base);
            case 0: break;
        }
#pragma unroll
        for (int t = 0; t < NACC; ++t)
        {
            r = __dmul_rn(r, base);
            acc[t] = __dadd_rn(acc[t], r);
        }
    }
}

// ---- a3: scalar inclusion-exclusion, k <= 10 (src/prob_any_missing.cpp:13-53) ---------------
__device__ MOIRE_COOP_INLINE double ie_scalar(double *scr, int ST, int k, int n)
{
    const unsigned short *tab = c_sub_tail + subset_block(k);
    const int nsub = (1 << k) - 1;
    double prob = 0.0;
    for (int idx = 0; idx < nsub; ++idx)
    {
        const unsigned int e = tab[idx];
        unsigned int tail = e & 1023u;
        int pos = (e >> 10) & 15;
        double base = pos ? scr[(kPartBase + pos - 1) * ST] : 1.0;
        for (;;)
        {
            const int el = __ffs(tail) - 1;
            tail &= tail - 1u;
            base = __dsub_rn(base, scr[el * ST]);
            if (!tail) break;
            scr[(kPartBase + pos) * ST] = base;
            ++pos;
        }
        double r = (e & 0x4000u) ? -1.0 : 1.0;
        int mc = n;
        while (mc > 0)
        {
            if (mc & 1) r = __dmul_rn(r, base);
            base = __dmul_rn(base, base);
            mc >>= 1;
        }
        prob = __dadd_rn(prob, r);
    }
    return prob;
}

// ---- a5: EGF coverage recurrence, k > 10 (src/prob_any_missing.cpp:57-97) -------------------
// The reference builds the binomials of row m on the fly, comb(m, 1) = m,
// comb(m, j + 1) = comb(m, j) * ((m - j) / (j + 1)) in double (src/prob_any_missing.cpp:81-88).
// They depend on (m, j) only, so the host tabulates them with the very same operations
// (engine.cu: build_egf_binomials) and the device never divides: g_egf_comb[j * 65 + m].
__device__ double g_egf_comb[(kMaxCoi + 1) * (kMaxCoi + 1)];
// C(n, i) as the nearest double, [n][i]: the weights of the linear-domain relatedness mixture.
__device__ double g_binom[(kMaxCoi + 1) * (kMaxCoi + 1)];


// ---- the recurrence as a band ---------------------------------------------------------------
// After e of the k alleles have been folded in, a_e[n] is exactly 0 for n < e (n draws cannot
// cover e alleles), and the result rows n = k .. max_n depend on a_e[n'] only for
// n' <= max_n - (k - e).  So stage e needs the D + 1 rows n = e + d, d = 0 .. D, D = max_n - k, and
//   band_e[d] = sum_{j = 1 .. d + 1} C(e + d, j) * band_{e-1}[d + 1 - j] * q^j
// (the reference's terms j > d + 1 multiply an exact zero and come last in its sum, so leaving
// them out changes no bit; rows outside the band never reach a result).  The update runs in
// place, d descending.  Work per cell: k (D + 1)(D + 2) / 2 multiply-adds instead of
// k max_n (max_n + 1) / 2 -- 3.6 x less at k = 13, COI 25, and almost nothing where k is close to
// the COI.  Each sum is still the reference's: terms in j order, (comb * a) * q^j, q^j by the
// running product q^(j-1) * q (src/prob_any_missing.cpp:68-91).
// band[d * ST], d <= D; on return band[d * ST] = coverage(k + d).
__device__ __forceinline__ void egf_band(mask_t latent, const double *prow, double total, int D,
                                         double *band, int ST)
{
    constexpr int CS = kMaxCoi + 1;
    for (int d = 0; d <= D; ++d) band[d * ST] = d == 0 ? 1.0 : 0.0;
    mask_t rest = latent;
    int e = 0;
    while (rest)
    {
        const int al = mffs(rest) - 1;
        rest &= rest - 1;
        ++e;
        const double q = __ddiv_rn(prow[al], total);
        int d = D;
        // two rows at a time: their sums are independent chains of additions
        for (; d >= 1; d -= 2)
        {
            const double *c0 = g_egf_comb + (e + d), *c1 = g_egf_comb + (e + d - 1);
            double s0 = 0.0, s1 = 0.0, pw = q;
#pragma unroll 2
            for (int j = 1; j <= d; ++j)
            {
                const double x0 = band[(d + 1 - j) * ST], x1 = band[(d - j) * ST];
                s0 = __dadd_rn(s0, __dmul_rn(__dmul_rn(c0[j * CS], x0), pw));
                s1 = __dadd_rn(s1, __dmul_rn(__dmul_rn(c1[j * CS], x1), pw));
                pw = __dmul_rn(pw, q);
            }
            s0 = __dadd_rn(s0, __dmul_rn(__dmul_rn(c0[(d + 1) * CS], band[0]), pw));  // j = d + 1
            band[d * ST] = s0;
            band[(d - 1) * ST] = s1;
        }
        if (d == 0) band[0] = __dadd_rn(0.0, __dmul_rn(__dmul_rn(g_egf_comb[CS + e], band[0]), q));
    }
}

// What a cell needs of its sample's relatedness r: log r, log(1 - r) (the reference's terms,
// src/chain.cpp:903-914) and the odds r / (1 - r) that the linear-domain mixture runs on.
struct Rel
{
    double log_r, log_1mr, odds;
};

// m log(total) + log(s), s > 0: ONE logarithm, of total^m * s, while total^m is comfortably a normal
// number (it nearly always is: total is a sum of allele frequencies, m a COI), else the two
// logarithms.  Which form runs depends on (total, m) only, so every evaluator takes the same one for
// a given cell.  The two differ by rounding (~1e-16 relative).
__device__ __forceinline__ double log_total_pow_times(double total, int m, double s)
{
    const double tm = pow_uint(total, m);
    if (tm >= 1e-250) return hot_log(__dmul_rn(tm, s));
    return __dadd_rn(__dmul_rn((double)m, cold_log(total)), cold_log(s));
}

// Sampler::dbinom (src/sampler.cpp:44-50); lgam[i] = lgamma(i + 1).
__device__ __forceinline__ double dbinom_log(int x, int size, double log_p, double log_1mp,
                                             const double *lgam)
{
    const double lp = __dadd_rn(__dmul_rn((double)x, log_p), __dmul_rn((double)(size - x), log_1mp));
    const double lc = __dsub_rn(__dsub_rn(lgam[size], lgam[x]), lgam[size - x]);
    return __dadd_rn(lp, lc);
}

// Relatedness mixture (src/chain.cpp:903-920, src/mcmc_utils.h:382-397).
// On entry val[t * ST] = PAM(n = k + t), t < cnt (val == nullptr: every PAM is exactly 0, the
// k = 1 case); term i = 0..cnt-1 uses n = m - i, i.e. t = cnt - 1 - i.
//
// The same sum with the common factor of term i = 0 taken out instead of the largest term:
//   sum_i C(m-1,i) r^i (1-r)^(m-1-i) total^(m-i) (1 - pam_i)
//     = (1-r)^(m-1) total^m * S,   S = sum_i C(m-1,i) q^i (1 - pam_i),  q = odds / total,
// so log(sum) = (m-1) log(1-r) + log(total^m S): one division and one logarithm per CELL in place
// of one log and one exp per TERM plus log(total) (round 1: 37 % of the evaluator's instructions
// went there; round 2 dropped the exp and the second log: -60 instructions per cell).  S has only
// positive terms (no cancellation) and starts at 1 - pam_0 >= eps; the weights stay finite while
// q^(cnt-1) <= 2^850 (C(m-1,i) < 2^124), otherwise the reference's log-sum-exp form below runs.
// Difference from the reference's order of operations: rounding level (<= ~1e-14 relative,
// tolerance 1e-9).  binom_s: rows [m - 1][i < kBinomCols] of g_binom staged in
// shared memory by the caller (or nullptr).
constexpr int kBinomCols = 16;
__device__ __noinline__ double relatedness_mixture(const double *val, int ST, int m, int cnt,
                                                   double total, const Rel rel, const double *lgam,
                                                   const double *binom_s)
{
    const double lead = __dmul_rn((double)(m - 1), rel.log_1mr);
    if (cnt == 1)
    {
        // m == k: the single term i = 0 (its binomial log-pmf is (m-1) log(1-r))
        const double pam = val ? clamp_pam(val[0]) : 0.0;
        return __dadd_rn(lead, log_total_pow_times(total, m, __dsub_rn(1.0, pam)));
    }
    const double q = __ddiv_rn(rel.odds, total);
    const int e2 = ((__double2hiint(q) >> 20) & 0x7ff) - 1022;  // q < 2^e2 (NaN / inf: 1025)
    if ((cnt - 1) * max(e2, 0) <= 850)
    {
        const double *cb = g_binom + (m - 1) * (kMaxCoi + 1);
        const double *cs = binom_s ? binom_s + (m - 1) * kBinomCols : cb;
        double qi = 1.0, S = 0.0;
        for (int i = 0; i < cnt; ++i)  // i ascending, as the reference sums; t = cnt - 1 - i
        {
            const double pam = val ? clamp_pam(val[(cnt - 1 - i) * ST]) : 0.0;
            const double c = i < kBinomCols ? cs[i] : cb[i];
            S = __dadd_rn(S, __dmul_rn(__dmul_rn(c, qi), __dsub_rn(1.0, pam)));
            qi = __dmul_rn(qi, q);
        }
        return __dadd_rn(lead, log_total_pow_times(total, m, S));
    }
    // the reference's log-sum-exp form (the terms are recomputed for the sum)
    const double log_total = cold_log(total);
    auto term = [&](int t) -> double {
        const int i = cnt - 1 - t;
        const double pr = dbinom_log(i, m - 1, rel.log_r, rel.log_1mr, lgam);
        const double pam = val ? clamp_pam(val[t * ST]) : 0.0;
        const double i_res = __dadd_rn(cold_log1m(pam), __dmul_rn(log_total, (double)(m - i)));
        return __dadd_rn(pr, i_res);
    };
    double mx = -INFINITY;
    for (int t = 0; t < cnt; ++t)
    {
        const double v = term(t);
        mx = (mx < v) ? v : mx;  // std::max_element semantics
    }
    if (mx == -INFINITY) return -INFINITY;
    double sum = 0.0;
    for (int t = cnt - 1; t >= 0; --t)  // i ascending, as the reference sums
    {
        const double dv = __dsub_rn(term(t), mx);
        sum = __dadd_rn(sum, dv == 0.0 ? 1.0 : cold_exp(dv));
    }
    return __dadd_rn(mx, cold_log(sum));
}

// Relatedness off: log1p(-pam) + m log(total) (src/chain.cpp:895-899), as one logarithm.
__device__ __forceinline__ double norel_term(double pam, double total, int m)
{
    return log_total_pow_times(total, m, __dsub_rn(1.0, clamp_pam(pam)));
}

constexpr int kMaxAcc = 16;  // accumulators kept in registers

// update_p's PAM cache (p_flat.cuh): an evaluator compiled with CACHE leaves the raw PAM vector of
// the cell -- exactly what it hands to relatedness_mixture / norel_term -- at w[t * stride],
// t < m - k + 1 (relatedness off: w[0]), so that a later proposal that does not touch the cell's
// latent alleles (q unchanged, only their sum moves) re-runs the mixture alone.
struct PamOut
{
    double *w;
    size_t stride;
};
template <bool CACHE>
__device__ __forceinline__ void pam_store(const PamOut po, const double *val, int ST, int cnt)
{
    if (CACHE)
        for (int t = 0; t < cnt; ++t) po.w[(size_t)t * po.stride] = val[t * ST];
}
template <bool CACHE>
__device__ __forceinline__ double norel_term_c(double pam, double total, int m, const PamOut po)
{
    if (CACHE) po.w[0] = pam;
    return norel_term(pam, total, m);
}

// The subset sums of one cell, accumulators dumped to scr[t * ST], then the mixture.
// All lanes that arrive together run the accumulator tier of the largest count among them (the
// extra accumulators of a lane with fewer mixture terms are simply never read), so lanes with
// different counts -- and different k: the subset loop just ends earlier for them -- share one
// instruction stream.  Six tiers are compiled (2, 4, 6, 8, 12, 16).  The fused kernels, which were
// instruction-fetch bound, did best with two ({4, 16}: more distinct hot loops cost more than idle
// accumulators, profiles/variants_g.log); the flat evaluator's code fits the instruction cache
// and gains from the tighter fit.
// ---- small k, unrolled ---------------------------------------------------------------------------
// For k <= MOIRE_SMALL_K every lane of a class-uniform warp walks the same handful of subsets, so
// the walk is unrolled at compile time: q in registers, no table decode, no stack in shared
// memory, the leading multiplications a fixed run.  Subset i of the reference's order (by size,
// then lexicographic index tuples) is SmallSubsets<K>::mask[i]; each base is the reference's chain of
// subtractions 1 - q[e1] - q[e2] ..., each power the same chain of multiplications as in
// ie_vectorized, so the sums are bit-identical to the table-driven walk (a cell may take either
// path depending on its neighbours in the sorted list).
#ifndef MOIRE_SMALL_K
#define MOIRE_SMALL_K 6
#endif
template <int K>
struct SmallSubsets
{
    int mask[(1 << K) - 1];
    constexpr SmallSubsets() : mask{}
    {
        int n = 0;
        for (int size = 1; size <= K; ++size)
        {
            int c[K] = {};
            for (int i = 0; i < size; ++i) c[i] = i;
            for (;;)
            {
                int m = 0;
                for (int i = 0; i < size; ++i) m |= 1 << c[i];
                mask[n++] = m;
                int i = size - 1;
                while (i >= 0 && c[i] == K - size + i) --i;
                if (i < 0) break;
                ++c[i];
                for (int j = i + 1; j < size; ++j) c[j] = c[j - 1] + 1;
            }
        }
    }
};

template <int K, int NACC>
__device__ __forceinline__ void ie_small_regs(const double (&q)[K], double (&acc)[NACC]);

template <int K, int NACC>
__device__ __forceinline__ void ie_small(const double *scr, int ST, double (&acc)[NACC])
{
    double q[K];
#pragma unroll
    for (int e = 0; e < K; ++e) q[e] = scr[e * ST];
    ie_small_regs<K, NACC>(q, acc);
}

template <int K, int NACC>
__device__ __forceinline__ void ie_small_regs(const double (&q)[K], double (&acc)[NACC])
{
    static_assert(K >= 2, "one latent allele has no subset sums (every PAM is exactly 0)");
    constexpr SmallSubsets<K> ord{};
#pragma unroll
    for (int t = 0; t < NACC; ++t) acc[t] = 0.0;
#pragma unroll
    for (int i = 0; i < (1 << K) - 1; ++i)
    {
        const int msk = ord.mask[i];
        double base = 1.0;
        int size = 0;
#pragma unroll
        for (int e = 0; e < K; ++e)
            if ((msk >> e) & 1)
            {
                base = __dsub_rn(base, q[e]);
                ++size;
            }
        // sign * base is exact (no rounding), so the chain of K - 1 leading multiplications starts at
        // +-base: one multiplication and one constant fewer per subset, the same bits
        double r = (size & 1) ? base : -base;
#pragma unroll
        for (int j = 0; j < K - 2; ++j) r = __dmul_rn(r, base);
#pragma unroll
        for (int t = 0; t < NACC; ++t)
        {
            r = __dmul_rn(r, base);
            acc[t] = __dadd_rn(acc[t], r);
        }
    }
}

// PAM(n = k + t), t < cnt <= kMaxAcc, into scr[t * ST] (q in slots 0..k-1 is consumed).
// `lanes`: the lanes of the warp that call together (a ballot taken by transmission_ll before
// its branches diverge), so the tier and the class-uniform fast path depend on the data of
// exactly those lanes, never on how the scheduler happened to converge them.
__device__ __noinline__ void ie_pams(double *scr, int ST, int k, int cnt, unsigned lanes)
{
    const int cnt_w = __reduce_max_sync(lanes, cnt);
#define MOIRE_IE_STORE(NACC) _Pragma("unroll") for (int t = 0; t < NACC; ++t) scr[t * ST] = acc[t];
#define MOIRE_IE_TIER(NACC)                      \
    {                                            \
        double acc[NACC];                        \
        ie_vectorized<NACC>(scr, ST, k, 0, acc); \
        MOIRE_IE_STORE(NACC)                     \
    }
#define MOIRE_IE_SMALL(K, NACC)          \
    {                                    \
        double acc[NACC];                \
        ie_small<K, NACC>(scr, ST, acc); \
        MOIRE_IE_STORE(NACC)             \
    }
#define MOIRE_IE_SMALL_TIERS(K)                    \
    {                                              \
        if (cnt_w <= 2) MOIRE_IE_SMALL(K, 2)       \
        else if (cnt_w <= 4) MOIRE_IE_SMALL(K, 4)  \
        else if (cnt_w <= 6) MOIRE_IE_SMALL(K, 6)  \
        else if (cnt_w <= 8) MOIRE_IE_SMALL(K, 8)  \
        else if (cnt_w <= 12) MOIRE_IE_SMALL(K, 12) \
        else MOIRE_IE_SMALL(K, 16)                 \
        return;                                    \
    }
#if MOIRE_SMALL_K >= 2
    {
        const int k_hi = __reduce_max_sync(lanes, k);
        if (k_hi <= MOIRE_SMALL_K && k_hi >= 2 && __reduce_min_sync(lanes, k) == k_hi)
        {
            if (k_hi == 2) MOIRE_IE_SMALL_TIERS(2)
#if MOIRE_SMALL_K >= 3
            if (k_hi == 3) MOIRE_IE_SMALL_TIERS(3)
#endif
#if MOIRE_SMALL_K >= 4
            if (k_hi == 4) MOIRE_IE_SMALL_TIERS(4)
#endif
#if MOIRE_SMALL_K >= 5
            if (k_hi == 5) MOIRE_IE_SMALL_TIERS(5)
#endif
#if MOIRE_SMALL_K >= 6
            if (k_hi == 6) MOIRE_IE_SMALL_TIERS(6)
#endif
        }
    }
#endif
    if (cnt_w <= 2) MOIRE_IE_TIER(2)
    else if (cnt_w <= 4) MOIRE_IE_TIER(4)
    else if (cnt_w <= 6) MOIRE_IE_TIER(6)
    else if (cnt_w <= 8) MOIRE_IE_TIER(8)
    else if (cnt_w <= 12) MOIRE_IE_TIER(12)
    else MOIRE_IE_TIER(16)
#undef MOIRE_IE_TIER
#undef MOIRE_IE_SMALL
#undef MOIRE_IE_SMALL_TIERS
#undef MOIRE_IE_STORE
}

template <bool CACHE>
__device__ __noinline__ double transmission_rel_ie(double *scr, int ST, int k, int m, double total,
                                                   const Rel rel, const double *lgam,
                                                   const double *binom_s, unsigned lanes, const PamOut po)
{
    const int cnt = m - k + 1;
    ie_pams(scr, ST, k, cnt, lanes);
    pam_store<CACHE>(po, scr, ST, cnt);
    return relatedness_mixture(scr, ST, m, cnt, total, rel, lgam, binom_s);
}

// m - k >= kMaxAcc (rare: a COI far above the number of distinct alleles): the accumulators are
// produced kMaxAcc at a time and parked in local memory.
template <bool CACHE>
__device__ __noinline__ double transmission_rel_ie_wide(double *scr, int ST, int k, int m,
                                                        double total, const Rel rel,
                                                        const double *lgam, const double *binom_s,
                                                        const PamOut po)
{
    double v[kMaxCoi];
    const int cnt = m - k + 1;
    for (int t0 = 0; t0 < cnt; t0 += kMaxAcc)
    {
        double acc[kMaxAcc];
        ie_vectorized<kMaxAcc>(scr, ST, k, t0, acc);
#pragma unroll
        for (int j = 0; j < kMaxAcc; ++j)
            if (t0 + j < cnt) v[t0 + j] = acc[j];
    }
    pam_store<CACHE>(po, v, 1, cnt);
    return relatedness_mixture(v, 1, m, cnt, total, rel, lgam, binom_s);
}

// EGF branch (k > 10) of one cell on one thread: the band in the thread's scratch column while it
// fits (m - k < kScratchSlots: nearly always), else in local memory; then the same mixture as every
// other path.
template <bool CACHE>
__device__ __noinline__ double transmission_egf(mask_t latent, const double *prow, double total,
                                                int k, int m, const Rel rel, bool allow_relatedness,
                                                double *scr, int ST, const double *lgam,
                                                const double *binom_s, const PamOut po)
{
    const int D = m - k;
    double local_band[kMaxCoi + 1];
    double *band = D < kScratchSlots ? scr : local_band;
    const int BS = D < kScratchSlots ? ST : 1;
    egf_band(latent, prow, total, D, band, BS);
    // src/chain.cpp:895-899 with probAnyMissingFunctor::operator() (prob_any_missing.cpp:149-153)
    if (!allow_relatedness) return norel_term_c<CACHE>(pam_from_coverage(band[D * BS]), total, m, po);
    for (int d = 0; d <= D; ++d) band[d * BS] = pam_from_coverage(band[d * BS]);
    pam_store<CACHE>(po, band, BS, D + 1);
    return relatedness_mixture(band, BS, m, D + 1, total, rel, lgam, binom_s);
}

// Relatedness off and n == k: 1 - k! * prod(q) in double (src/prob_any_missing.cpp:139-148).
template <bool CACHE>
__device__ __forceinline__ double transmission_norel_exact_cover(mask_t latent, const double *prow,
                                                                 double total, int k, const PamOut po)
{
    double cov = 1.0;
    mask_t rest = latent;
    while (rest)
    {
        const int al = mffs(rest) - 1;
        rest &= rest - 1;
        cov = __dmul_rn(cov, __ddiv_rn(prow[al], total));
    }
    for (int i = 2; i <= k; ++i) cov = __dmul_rn(cov, (double)i);
    return norel_term_c<CACHE>(pam_from_coverage(cov), total, k, po);
}

// The first kGather allele frequencies of a latent genotype, ascending allele order.  The loads do
// not depend on one another (an exhausted genotype re-reads entry 0), so they are in flight
// together: one trip to L2 instead of k dependent ones in front of every cell (the evaluator's
// top stall was the long scoreboard).
constexpr int kGather = 6;
__device__ __forceinline__ void gather_freqs(mask_t latent, const double *prow, double (&pv)[kGather])
{
    mask_t rest = latent;
#pragma unroll
    for (int e = 0; e < kGather; ++e)
    {
        const int al = rest ? mffs(rest) - 1 : 0;
        pv[e] = prow[al];
        rest &= rest - 1;
    }
}

// ---- a2: transmission process of one cell, one thread -------------------------------------------
// prow: allele frequencies of the locus; scr: the thread's scratch column (kScratchSlots slots,
// stride ST); lgam: lgamma(i + 1), i < 128; binom_s: see relatedness_mixture.  `callers`: the lanes
// of the warp that make this call together (a __ballot_sync the caller takes BEFORE branching
// into the call).
template <bool CACHE>
__device__ __forceinline__ double transmission_ll(mask_t latent, const double *prow,
                                                  const double (&pv)[kGather], int m, const Rel rel,
                                                  bool allow_relatedness, double *scr, int ST,
                                                  const double *lgam, const double *binom_s,
                                                  unsigned callers, const PamOut po)
{
    const int k = mpopc(latent);
    // the lanes that will reach the register-tier inclusion-exclusion sums, agreed on while the
    // callers are still together (the branches below diverge by k and m)
    const unsigned ie_lanes = __ballot_sync(
        callers, allow_relatedness && k >= 2 && k <= m && k <= kIeMax && m - k + 1 <= kMaxAcc);
    if (k == 0 || k > m) return NAN;  // unreachable for states the sampler produces
    if (k == 1)
    {
        // One latent allele: q = p / p = 1 exactly, every base 1 - q is exactly 0 and so is every
        // PAM (in each of the reference's branches) -- the subset sums have nothing to add.
        const double total = __dadd_rn(0.0, pv[0]);
        if (!allow_relatedness) return norel_term_c<CACHE>(0.0, total, m, po);
        return relatedness_mixture(nullptr, 0, m, m, total, rel, lgam, binom_s);
    }
    // the frequency sum in ascending allele order: the gathered ones, then any further ones
    double total = 0.0;
    mask_t tail = latent;
#pragma unroll
    for (int e = 0; e < kGather; ++e)
    {
        if (e < k) total = __dadd_rn(total, pv[e]);
        tail &= tail - 1;
    }
    if (k > kGather)
    {
        mask_t rest = tail;
        while (rest)
        {
            const int al = mffs(rest) - 1;
            rest &= rest - 1;
            total = __dadd_rn(total, prow[al]);
        }
    }
    if (k <= kIeMax)
    {
#pragma unroll
        for (int e = 0; e < kGather; ++e)
            if (e < k) scr[e * ST] = __ddiv_rn(pv[e], total);
        if (k > kGather)
        {
            mask_t rest = tail;
            int e = kGather;
            while (rest)
            {
                const int al = mffs(rest) - 1;
                rest &= rest - 1;
                scr[e * ST] = __ddiv_rn(prow[al], total);
                ++e;
            }
        }
    }
    if (!allow_relatedness)
    {
        // src/chain.cpp:895-899 with probAnyMissingFunctor::operator() (prob_any_missing.cpp:131-154)
        if (m == k) return transmission_norel_exact_cover<CACHE>(latent, prow, total, k, po);
        if (k > kIeMax) return transmission_egf<CACHE>(latent, prow, total, k, m, rel, false, scr, ST, lgam, binom_s, po);
        return norel_term_c<CACHE>(ie_scalar(scr, ST, k, m), total, m, po);
    }
    if (k > kIeMax) return transmission_egf<CACHE>(latent, prow, total, k, m, rel, true, scr, ST, lgam, binom_s, po);
    if (m - k + 1 <= kMaxAcc) return transmission_rel_ie<CACHE>(scr, ST, k, m, total, rel, lgam, binom_s, ie_lanes, po);
    return transmission_rel_ie_wide<CACHE>(scr, ST, k, m, total, rel, lgam, binom_s, po);
}

// ---- a2, class-uniform small cells, fully in registers ---------------------------------------------
// ncu, round 2 (profiles/r02g): at the benchmark shapes the subset sums themselves are ~13 % of the
// evaluator's instructions; the rest is everything around them -- 64-bit bit scans, predicated
// gathers, q and the sums travelling through shared memory between out-of-line functions, the
// generic mixture loop.  A chunk whose 32 cells share k <= kSmallK (nearly every chunk: the list is
// class-sorted) takes this path instead: K known at compile time, the K frequencies loaded
// together, q, the accumulators and the mixture terms in registers, one call-free stretch of
// straight code per (K, tier).  Same operations in the same order as the generic path (the sums
// are ie_small's, the mixture is relatedness_mixture's fast branch; its rare fallback is called
// with the sums parked in the scratch column), so a cell gets the same bits whichever path its
// neighbours send it down.
constexpr int kSmallK = 6;      // <= MOIRE_SMALL_K
constexpr int kSmallCnt = 8;    // mixture terms handled in registers

template <int K, int NACC, bool CACHE>
__device__ __forceinline__ double small_cell_tail(const double (&q)[K], double total, int m, int cnt,
                                                  const Rel rel, double *scr, int ST, const double *lgam,
                                                  const double *binom_s, const PamOut po)
{
    double acc[NACC];
    if constexpr (K >= 2)
    {
        ie_small_regs<K, NACC>(q, acc);
        if (CACHE)
        {
#pragma unroll
            for (int t = 0; t < NACC; ++t)
                if (t < cnt) po.w[(size_t)t * po.stride] = acc[t];
        }
    }
    else
    {
#pragma unroll
        for (int t = 0; t < NACC; ++t) acc[t] = 0.0;  // one latent allele: every PAM is exactly 0
    }
    const double lead = __dmul_rn((double)(m - 1), rel.log_1mr);
    if (cnt == 1)
        return __dadd_rn(lead, log_total_pow_times(total, m, __dsub_rn(1.0, clamp_pam(acc[0]))));
    const double qr = __ddiv_rn(rel.odds, total);
    const int e2 = ((__double2hiint(qr) >> 20) & 0x7ff) - 1022;
    if ((cnt - 1) * max(e2, 0) > 850)
    {
        // weights could overflow: the generic function's log-sum-exp branch, sums via the scratch column
#pragma unroll
        for (int t = 0; t < NACC; ++t) scr[t * ST] = acc[t];
        return relatedness_mixture(scr, ST, m, cnt, total, rel, lgam, binom_s);
    }
    const double *cs = binom_s + (m - 1) * kBinomCols;  // cnt <= kSmallCnt <= kBinomCols
    double qi = 1.0, S = 0.0;
    // term i = cnt - 1 - t uses acc[t]: t descending is i ascending, the order of the generic loop
#pragma unroll
    for (int t = NACC - 1; t >= 0; --t)
        if (t < cnt)
        {
            S = __dadd_rn(S, __dmul_rn(__dmul_rn(cs[cnt - 1 - t], qi), __dsub_rn(1.0, clamp_pam(acc[t]))));
            qi = __dmul_rn(qi, qr);
        }
    return __dadd_rn(lead, log_total_pow_times(total, m, S));
}

// lat32: the latent genotype (all alleles below 32); cnt_hi: the largest m - k + 1 among the lanes
// that call together (warp-uniform).  All of them have exactly K latent alleles.
template <int K, bool CACHE>
__device__ __forceinline__ double small_cell(unsigned lat32, const double *prow, int m, int cnt_hi,
                                             const Rel rel, double *scr, int ST, const double *lgam,
                                             const double *binom_s, const PamOut po)
{
    double q[K];
    unsigned rest = lat32;
#pragma unroll
    for (int e = 0; e < K; ++e)
    {
        q[e] = prow[__ffs(rest) - 1];
        rest &= rest - 1u;
    }
    double total = 0.0;
#pragma unroll
    for (int e = 0; e < K; ++e) total = __dadd_rn(total, q[e]);
    if (K >= 2)
    {
#pragma unroll
        for (int e = 0; e < K; ++e) q[e] = __ddiv_rn(q[e], total);
    }
    const int cnt = m - K + 1;
    if (cnt_hi <= 2) return small_cell_tail<K, 2, CACHE>(q, total, m, cnt, rel, scr, ST, lgam, binom_s, po);
    if (cnt_hi <= 4) return small_cell_tail<K, 4, CACHE>(q, total, m, cnt, rel, scr, ST, lgam, binom_s, po);
    return small_cell_tail<K, kSmallCnt, CACHE>(q, total, m, cnt, rel, scr, ST, lgam, binom_s, po);
}

// ---- a2, warp-cooperative: one cell, 32 lanes ---------------------------------------------------
// For a heavy cell (2^k - 1 subsets, k >= kCoopMinK) the subsets are dealt 32 at a time to the
// lanes of a warp.  Each lane builds its subset's terms base^(k+t), t < cnt, with the same chain
// of operations as the one-thread path and parks them in shared memory; lane t then adds the 32
// terms of accumulator t in subset order.  Floating-point addition is not associative, so this
// ordered second phase (not a tree reduction) is what keeps the result bit-identical to the
// one-thread path and to the reference's summation order.  The sums are then gathered to every
// lane and go through relatedness_mixture, the one-thread path's own mixture.
// All lanes must call with identical arguments; every lane returns the value.
// wscr: the warp's 32 scratch columns (slot s, column c at wscr[s * ST + c]); slot 16 holds q.
#ifndef MOIRE_COOP_MIN_K
#define MOIRE_COOP_MIN_K 8
#endif
#ifdef MOIRE_EVAL_TRACE
__device__ unsigned long long *g_coop_trace = nullptr;  // tuning aid: phase clocks of cooperative cells
#endif
constexpr int kCoopMinK = MOIRE_COOP_MIN_K;
template <bool CACHE>
__device__ MOIRE_COOP_INLINE double transmission_ll_coop(mask_t latent, const double *prow, int m,
                                                       const Rel rel, bool allow_relatedness,
                                                       double *wscr, int ST, const double *lgam,
                                                       const double *binom_s, const PamOut po)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int k = mpopc(latent);  // 1 <= k <= kIeMax, k <= m guaranteed by the caller
#ifdef MOIRE_EVAL_TRACE
    long long tc0 = clock64(), tc1 = 0, tc2 = 0, tc3 = 0;
#endif
    // lane e < k fetches the frequency of element e (the k loads are in flight together: the trace
    // showed ~3000 cycles for k dependent trips to L2 in front of every cooperative cell); the sum is
    // then taken in ascending allele order, as everywhere
    double total = 0.0, p_mine = 0.0;
    {
        mask_t rest = latent;
        for (int e = 0; e < lane && rest; ++e) rest &= rest - 1;
        if (lane < k) p_mine = prow[mffs(rest) - 1];
        for (int e = 0; e < k; ++e) total = __dadd_rn(total, __shfl_sync(FULL, p_mine, e));
    }
    if (!allow_relatedness && m == k) return transmission_norel_exact_cover<CACHE>(latent, prow, total, k, po);
    double *q = wscr + 16 * ST;
    if (lane < k) q[lane] = __ddiv_rn(p_mine, total);  // lane e normalises element e
    __syncwarp();
#ifdef MOIRE_EVAL_TRACE
    tc1 = clock64();
#endif
    const unsigned short *masks = g_sub_mask + subset_block(k);
    const int nsub = (1 << k) - 1;
    const int cnt = allow_relatedness ? m - k + 1 : 1;
    // accumulators of up to 8 chunks of 16 mixture terms (cnt <= 127); chunk c, lane t: n = k + 16c + t
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
    {
        const int t0 = 16 * ch;
        if (t0 < cnt)  // uniform
        {
            const int cc = min(16, cnt - t0);
            for (int b0 = 0; b0 < nsub; b0 += 32)
            {
                const int idx = b0 + lane;
                if (idx < nsub)
                {
                    unsigned int msk = masks[idx];
                    const bool neg = (__popc(msk) & 1) == 0;
                    double base = 1.0;
                    while (msk)
                    {
                        const int el = __ffs(msk) - 1;
                        msk &= msk - 1u;
                        base = __dsub_rn(base, q[el]);
                    }
                    double r = neg ? -1.0 : 1.0;
                    if (allow_relatedness)
                    {
                        const int lead = k - 1 + t0;
                        if (lead > 0) r = neg ? -base : base;  // sign * base, exact
                        for (int j = 1; j < lead; ++j) r = __dmul_rn(r, base);
                        for (int t = 0; t < cc; ++t)
                        {
                            r = __dmul_rn(r, base);
                            wscr[t * ST + lane] = r;
                        }
                    }
                    else
                    {
                        int mc = m;  // probAnyMissingFunctor::operator(): square and multiply
                        while (mc > 0)
                        {
                            if (mc & 1) r = __dmul_rn(r, base);
                            base = __dmul_rn(base, base);
                            mc >>= 1;
                        }
                        wscr[lane] = r;
                    }
                }
                __syncwarp();
                if (lane < cc)
                {
                    const int nv = min(32, nsub - b0);
                    const double *row = wscr + lane * ST;
                    double a = acc[ch];
                    if (nv == 32)
                    {
                        double term[32];
#pragma unroll
                        for (int l = 0; l < 32; ++l) term[l] = row[l];  // loads first, then the chain
#pragma unroll
                        for (int l = 0; l < 32; ++l) a = __dadd_rn(a, term[l]);
                    }
                    else if (nv >= 16)
                    {
                        // a short last batch (2^k - 1 is never a multiple of 32) the same way, padded with
                        // zeros: x + 0 is x -- at most the sign of a zero sum changes, and a PAM of +-0
                        // clamps to 0
                        double term[32];
#pragma unroll
                        for (int l = 0; l < 32; ++l) term[l] = row[l];
#pragma unroll
                        for (int l = 0; l < 32; ++l) a = __dadd_rn(a, l < nv ? term[l] : 0.0);
                    }
                    else
                    {
                        for (int l = 0; l < nv; ++l) a = __dadd_rn(a, row[l]);
                    }
                    acc[ch] = a;
                }
                __syncwarp();
            }
        }
    }
#ifdef MOIRE_EVAL_TRACE
    tc2 = clock64();
#endif
    if (!allow_relatedness) return norel_term_c<CACHE>(__shfl_sync(FULL, acc[0], 0), total, m, po);
    // Up to 32 mixture terms: lane L builds term i = cnt - 1 - L (its own q^i by the same chain of
    // multiplications, its own clamp) and the terms are then added in the order i = 0, 1, ... through
    // shuffles -- the operations and the order of relatedness_mixture's linear branch, bit for bit,
    // without every lane walking the whole vector through local memory (trace: gather + mixture were
    // ~45 % of a cooperative cell's 23 000 cycles).  Its other branches (one term, weights that could
    // overflow) go through the function itself, below.
    if (cnt > 1 && cnt <= 32)
    {
        const double qr = __ddiv_rn(rel.odds, total);
        const int e2 = ((__double2hiint(qr) >> 20) & 0x7ff) - 1022;
        if ((cnt - 1) * max(e2, 0) <= 850)
        {
            const double v0 = __shfl_sync(FULL, acc[0], lane & 15), v1 = __shfl_sync(FULL, acc[1], lane & 15);
            const double pam = (lane >> 4) ? v1 : v0;  // PAM(n = k + lane)
            if (CACHE && lane < cnt) po.w[(size_t)lane * po.stride] = pam;
            const int i = cnt - 1 - lane;  // this lane's term (negative: none)
            double qi = 1.0;
            for (int j = 0; j < cnt - 1; ++j)
                if (j < i) qi = __dmul_rn(qi, qr);
            double term = 0.0;
            if (i >= 0)
            {
                const double c = g_binom[(m - 1) * (kMaxCoi + 1) + i];
                term = __dmul_rn(__dmul_rn(c, qi), __dsub_rn(1.0, clamp_pam(pam)));
            }
            double S = 0.0;
            for (int j = 0; j < cnt; ++j) S = __dadd_rn(S, __shfl_sync(FULL, term, cnt - 1 - j));
            const double lead = __dmul_rn((double)(m - 1), rel.log_1mr);
#ifdef MOIRE_EVAL_TRACE
            if (g_coop_trace && lane == 0)
            {
                const unsigned slot = atomicAdd((unsigned int *)g_coop_trace, 1u);
                if (slot < 8000)
                {
                    unsigned long long *r = g_coop_trace + 8 + (size_t)slot * 8;
                    r[0] = k; r[1] = cnt; r[2] = tc1 - tc0; r[3] = tc2 - tc1; r[4] = 0; r[5] = clock64() - tc2;
                }
            }
#endif
            return __dadd_rn(lead, log_total_pow_times(total, m, S));
        }
    }
    // The mixture through the same function, in the same order of operations, as the one-thread
    // path: which evaluator a cell lands in depends on the launch (the cooperative threshold
    // follows the work per thread, so it changes with the number of chains on the device) and must
    // not change a bit of the result.  Every lane gathers the sums and computes the same value.
    double pams[kMaxCoi];
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
        if (16 * ch < cnt)  // uniform
        {
            const int top = min(16, cnt - 16 * ch);
            for (int t = 0; t < top; ++t) pams[16 * ch + t] = __shfl_sync(FULL, acc[ch], t);
        }
    if (CACHE)  // every lane holds the vector: lane t stores entries t, t + 32, ...
        for (int t = lane; t < cnt; t += 32) po.w[(size_t)t * po.stride] = pams[t];
#ifdef MOIRE_EVAL_TRACE
    tc3 = clock64();
    const double res_traced = relatedness_mixture(pams, 1, m, cnt, total, rel, lgam, binom_s);
    if (g_coop_trace && lane == 0)
    {
        const unsigned slot = atomicAdd((unsigned int *)g_coop_trace, 1u);
        if (slot < 8000)
        {
            unsigned long long *r = g_coop_trace + 8 + (size_t)slot * 8;
            r[0] = k; r[1] = cnt; r[2] = tc1 - tc0; r[3] = tc2 - tc1; r[4] = tc3 - tc2; r[5] = clock64() - tc3;
        }
    }
    return res_traced;
#else
    return relatedness_mixture(pams, 1, m, cnt, total, rel, lgam, binom_s);
#endif
}

// ---- a2 for k > 10, warp-cooperative: the band with its rows across lanes ---------------------
// One warp per cell.  The rows of a stage are independent, so lane l computes row d = l (and
// l + 32, ... for bands wider than a warp), each as the sequential sum of egf_band; the band lives
// in the warp's scratch rows.  Every lane then gathers the band and runs the one mixture function.
// All lanes must call with identical arguments; every lane returns the value.
template <bool CACHE>
__device__ MOIRE_COOP_INLINE double transmission_ll_egf_coop(mask_t latent, const double *prow, int m,
                                                           const Rel rel, bool allow_relatedness,
                                                           double *wscr, int ST, const double *lgam,
                                                           const double *binom_s, const PamOut po)
{
    constexpr int CS = kMaxCoi + 1;
    const int lane = threadIdx.x & 31;
    const int k = mpopc(latent);
    const int D = m - k;
    double total = 0.0;
    {
        mask_t rest = latent;
        while (rest)
        {
            const int al = mffs(rest) - 1;
            rest &= rest - 1;
            total = __dadd_rn(total, prow[al]);
        }
    }
    // band[d] at wscr[(d >> 5) * ST + (d & 31)], d <= D < 128: rows 0..3
    auto B = [&](int d) -> double & { return wscr[(d >> 5) * ST + (d & 31)]; };
    for (int d = lane; d <= D; d += 32) B(d) = d == 0 ? 1.0 : 0.0;
    __syncwarp();
    mask_t rest = latent;
    int e = 0;
    while (rest)
    {
        const int al = mffs(rest) - 1;
        rest &= rest - 1;
        ++e;
        const double q = __ddiv_rn(prow[al], total);
        double nb[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int part = 0; part < 4; ++part)
        {
            const int d = lane + 32 * part;
            if (d <= D)
            {
                const double *c = g_egf_comb + (e + d);
                double sum = 0.0, pw = q;
                for (int j = 1; j <= d + 1; ++j)
                {
                    if (j > 1) pw = __dmul_rn(pw, q);
                    sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(c[j * CS], B(d + 1 - j)), pw));
                }
                nb[part] = sum;
            }
        }
        __syncwarp();
#pragma unroll
        for (int part = 0; part < 4; ++part)
        {
            const int d = lane + 32 * part;
            if (d <= D) B(d) = nb[part];
        }
        __syncwarp();
    }
    if (!allow_relatedness) return norel_term_c<CACHE>(pam_from_coverage(B(D)), total, m, po);
    double pams[kMaxCoi];
    for (int d = 0; d <= D; ++d) pams[d] = pam_from_coverage(B(d));
    __syncwarp();
    if (CACHE)
        for (int d = lane; d <= D; d += 32) po.w[(size_t)d * po.stride] = pams[d];
    return relatedness_mixture(pams, 1, m, D + 1, total, rel, lgam, binom_s);
}

// ---- a8: latent-genotype proposal ------------------------------------------------------------
// Uniform choice of `need` of the `n` set bits of `set` (selection sampling; consumes one
// uniform per visited bit while the outcome is still open).  Distributionally what the
// reference's shuffle + prefix does (src/sampler.cpp:312-322).
__device__ __forceinline__ mask_t choose_bits(Stream &rng, mask_t set, int n, int need)
{
    if (need <= 0) return 0;
    if (need >= n) return set;
    if (need == 1 || need == n - 1)
    {
        // keep one / drop one (what a proposal with a single false positive or false negative
        // needs -- nearly every proposal that needs a choice at all): one uniform picks the allele
        const int idx = (int)(rng.next_uniform() * (double)n);
        mask_t s = set;
        for (int i = 0; i < idx; ++i) s &= s - 1;
        const mask_t bit = mlowest(s);
        return need == 1 ? bit : set ^ bit;
    }
    mask_t out = 0;
    int remaining = n;
    while (need > 0)
    {
        if (need == remaining)
        {
            out |= set;
            break;
        }
        const mask_t low = mlowest(set);
        set ^= low;
        const double u = rng.next_uniform();
        if (u * (double)remaining < (double)need)
        {
            out |= low;
            --need;
        }
        --remaining;
    }
    return out;
}

// Draws the proposal and returns its mask; *log_prob gets the log proposal density.
// logC[n * kLogCDim + k] = log(C(n, k)).  Stream use: word 0 -> the number of false negatives, word 1 -> the
// number of false positives (both by inversion, always consumed), then one word per visited allele
// while the two uniform choices are still open (none when every observed allele is kept and no
// false negative was drawn -- the usual cell): one Philox block per cell, where the round-1 sampler
// (a Bernoulli word per candidate allele, each phase on a fresh block) used three.
__device__ __forceinline__ mask_t sample_latent_genotype(Stream &rng, mask_t obs, int A,
                                                           int coi, const EpsLogs &e,
                                                           const double *__restrict__ logC,
                                                           double *log_prob)
{
    const mask_t all = mall(A);
    const int obs_pos = mpopc(obs);
    const int obs_neg = A - obs_pos;
    const int min_fn = obs_pos == 0;
    const int max_fn = max(min_fn, min(coi, obs_neg));
    const double u_fn = rng.next_uniform();
    const double u_fp = rng.next_uniform();
    const int fn = min_fn + binom_inverse(u_fn, max_fn - min_fn, e.q_en, e.odds_en);
    const int tn = obs_neg - fn;
    const double lp_fn =
        __dadd_rn(__dadd_rn(logC[obs_neg * kLogCDim + (fn - min_fn)],
                            __dmul_rn((double)(fn - min_fn), e.l_en)),
                  __dmul_rn((double)tn, e.l_1men));
    const int min_fp = max(0, obs_pos + fn - coi);
    const int max_fp = min(obs_pos, obs_pos + fn - 1);
    if (max_fp < min_fp)
    {
        *log_prob = -INFINITY;
        return 0;
    }
    const int fp = min_fp + binom_inverse(u_fp, max_fp - min_fp, e.q_ep, e.odds_ep);
    const int tp = obs_pos - fp;
    const double lp_fp =
        __dadd_rn(__dadd_rn(logC[obs_pos * kLogCDim + (fp - min_fp)],
                            __dmul_rn((double)(fp - min_fp), e.l_ep)),
                  __dmul_rn((double)tp, e.l_1mep));
    const mask_t pos = choose_bits(rng, obs, obs_pos, tp);
    const mask_t neg = choose_bits(rng, ~obs & all, obs_neg, fn);
    const double lp_pos_idx = -logC[obs_pos * kLogCDim + tp];
    const double lp_neg_idx = -logC[obs_neg * kLogCDim + fn];
    *log_prob = __dadd_rn(__dadd_rn(__dadd_rn(lp_pos_idx, lp_neg_idx), lp_fp), lp_fn);
    return pos | neg;
}

// ---- proposals and priors --------------------------------------------------------------------
// Standard normal from two words (Box-Muller, cosine branch).
__device__ __noinline__ double next_normal(Stream &rng)
{
    const double u1 = rng.next_uniform();
    const double u2 = rng.next_uniform();
    return sqrt(-2.0 * cold_log(u1)) * cos(6.283185307179586477 * u2);
}

// Sampler::sample_coi_delta(mean) (src/sampler.cpp:152-156): +-Geometric(1/(1+mean)) on {0,1,..}.
__device__ __noinline__ int next_coi_delta(Stream &rng, double log_1mp)
{
    const int sign = (rng.next_u32() & 1u) ? 1 : -1;
    const double u = rng.next_uniform();
    const int g = (int)floor(cold_log(u) / log_1mp);
    return sign * g;
}

// Sampler::sample_constrained given its N(0, sd) draw eps (src/sampler.cpp:176-190).
__device__ __noinline__ void constrained_proposal(double eps, double curr, double lo, double hi,
                                                     double *prop_out, double *adj_out)
{
    const double l_cl = cold_log(__dsub_rn(curr, lo));
    const double l_hc = cold_log(__dsub_rn(hi, curr));
    const double unconstrained = __dsub_rn(l_cl, l_hc);
    const double exp_prop = cold_exp(__dadd_rn(eps, unconstrained));
    double prop = __ddiv_rn(__dadd_rn(__dmul_rn(hi, exp_prop), lo), __dadd_rn(exp_prop, 1.0));
    prop = prop < lo ? lo : (prop > hi ? hi : prop);
    const double adj = __dsub_rn(
        __dsub_rn(__dadd_rn(cold_log(__dsub_rn(prop, lo)), cold_log(__dsub_rn(hi, prop))), l_cl), l_hc);
    *prop_out = prop;
    *adj_out = adj;
}

// R::dbeta(x, a, b, log = TRUE) incl. R's edge cases at 0 and 1; lbeta = log B(a, b) from host.
__device__ __noinline__ double dbeta_log(double x, double a, double b, double lbeta)
{
    if (isnan(x)) return x;
    if (x < 0.0 || x > 1.0) return -INFINITY;
    if (x == 0.0)
    {
        if (a > 1.0) return -INFINITY;
        if (a < 1.0) return INFINITY;
        return cold_log(b);
    }
    if (x == 1.0)
    {
        if (b > 1.0) return -INFINITY;
        if (b < 1.0) return INFINITY;
        return cold_log(a);
    }
    return (a - 1.0) * cold_log(x) + (b - 1.0) * cold_log1p(-x) - lbeta;
}

// R::dgamma(x, shape, scale, log = TRUE); lg_shape = lgamma(shape), l_scale = cold_log(scale).
__device__ __noinline__ double dgamma_log(double x, double shape, double scale, double lg_shape,
                                             double l_scale)
{
    if (isnan(x)) return x;
    if (x < 0.0) return -INFINITY;
    if (x == 0.0)
    {
        if (shape < 1.0) return INFINITY;
        if (shape > 1.0) return -INFINITY;
        return -l_scale;
    }
    return (shape - 1.0) * cold_log(x) - x / scale - lg_shape - shape * l_scale;
}

// Sampler::dztpois (src/sampler.cpp:38-42).
__device__ __forceinline__ double dztpois_log(int x, double log_lambda, double log_expm1_lambda,
                                              const double *lgam)
{
    return __dsub_rn(__dsub_rn(__dmul_rn((double)x, log_lambda), log_expm1_lambda), lgam[x]);
}
}  // namespace moire
